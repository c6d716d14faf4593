"""Wire / disk format of collected pseudo-labels.

The reference exchanges what the teacher collected through the filesystem: every rank pickles
Python objects (`exemplars_{rank}_{type}.pkl`, modeling/detector/st_generalized_rcnn.py:134-162) and
`predictions.pth` (engine/inference.py:158-163), and every reader unpickles all of them.  Here the
unit is the fixed-size record the ranks already all-gather over NCCL (`parallel.all_gather_records`,
`modeling.pseudo_label.pack_records`):

    records [n_images, w_max, 8] fp32   (image_id, label_id, x1, y1, x2, y2, score, region_idx)
    counts  [n_images]           int32  valid rows per image

and the file is a flat, versioned, little-endian container of exactly that, so teacher inference
can run once offline and be replayed by every student epoch (np.memmap-able, no pickle):

    offset  0  magic  b"B2PL"            4 bytes
            4  version                   uint32   (1)
            8  n_images                  uint64
           16  w_max                     uint32
           20  n_fields                  uint32   (8)
           24  n_rows = sum(counts)      uint64
           32  meta_len                  uint32   UTF-8 JSON (class list, config, rank, ...)
           36  reserved                  28 bytes (zero)
           64  meta                      meta_len bytes, zero-padded to a multiple of 64
            .  counts                    n_images x int32, zero-padded to a multiple of 64
            .  rows                      n_rows x n_fields x fp32 -- only the valid rows, image-major

Host-side code only (numpy / torch CPU); the device never sees this module.
"""
import json
import os
import struct

import numpy as np
import torch

from .structures import BoxList

MAGIC = b"B2PL"
VERSION = 1
FIELDS = ("image_id", "label_id", "x1", "y1", "x2", "y2", "score", "region_idx")
_HEADER = struct.Struct("<4sIQIIQI28x")
assert _HEADER.size == 64


def _pad64(n):
    return (n + 63) // 64 * 64


def _to_numpy(records, counts):
    if isinstance(records, torch.Tensor):
        records = records.detach().cpu().numpy()
    if isinstance(counts, torch.Tensor):
        counts = counts.detach().cpu().numpy()
    records = np.ascontiguousarray(records, dtype="<f4")
    counts = np.ascontiguousarray(counts, dtype="<i4")
    if records.ndim != 3 or records.shape[2] != len(FIELDS):
        raise ValueError("records must be [n_images, w_max, %d]" % len(FIELDS))
    if counts.shape != (records.shape[0],):
        raise ValueError("counts must be [n_images]")
    if counts.size and (counts.min() < 0 or counts.max() > records.shape[1]):
        raise ValueError("counts out of range [0, w_max]")
    return records, counts


def save_records(path, records, counts, meta=None):
    """Write the padded records of `n_images` images (device or host tensors / arrays) to `path`.
    Only the valid rows are stored.  The file is written to `path + '.tmp'` and renamed, so a reader
    never sees a partial file."""
    records, counts = _to_numpy(records, counts)
    n_images, w_max, n_fields = records.shape
    meta_b = json.dumps(meta or {}, sort_keys=True).encode("utf-8")
    valid = np.arange(w_max)[None, :] < counts[:, None]
    rows = records[valid]  # image-major, [n_rows, n_fields]
    # image_id, label_id and region_idx travel as fp32: exact only below 2^24
    if rows.size and float(np.abs(rows[:, [0, 1, 7]]).max()) >= float(1 << 24):
        raise ValueError("image_id / label_id / region_idx must be below 2^24 (they are stored as fp32)")
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_HEADER.pack(MAGIC, VERSION, n_images, w_max, n_fields, rows.shape[0], len(meta_b)))
        f.write(meta_b.ljust(_pad64(len(meta_b)), b"\0"))
        f.write(counts.tobytes().ljust(_pad64(counts.nbytes), b"\0"))
        f.write(rows.tobytes())
    os.replace(tmp, path)
    return rows.shape[0]


def _read_header(f):
    raw = f.read(_HEADER.size)
    if len(raw) != _HEADER.size:
        raise ValueError("truncated pseudo-label file")
    magic, version, n_images, w_max, n_fields, n_rows, meta_len = _HEADER.unpack(raw)
    if magic != MAGIC:
        raise ValueError("not a pseudo-label record file (bad magic %r)" % magic)
    if version != VERSION:
        raise ValueError("pseudo-label file version %d, this reader understands %d" % (version, VERSION))
    if n_fields != len(FIELDS):
        raise ValueError("pseudo-label file has %d fields per row, expected %d" % (n_fields, len(FIELDS)))
    return n_images, w_max, n_fields, n_rows, meta_len


def load_records(path, padded=True, mmap=False):
    """-> (records, counts, meta).  padded=True: records [n_images, w_max, 8] fp32 torch tensor (zero
    rows past counts), exactly what was saved; padded=False: the packed rows [n_rows, 8] (with
    mmap=True a read-only np.memmap, nothing is copied) and counts give the image boundaries."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        n_images, w_max, n_fields, n_rows, meta_len = _read_header(f)
        meta_raw = f.read(_pad64(meta_len))
        off_counts = _HEADER.size + _pad64(meta_len)
        off_rows = off_counts + _pad64(4 * n_images)
        if size != off_rows + 4 * n_fields * n_rows:
            raise ValueError("pseudo-label file size does not match its header (truncated or corrupt)")
        meta = json.loads(meta_raw[:meta_len].decode("utf-8")) if meta_len else {}
        counts = np.frombuffer(f.read(_pad64(4 * n_images))[: 4 * n_images], dtype="<i4").copy()
    if int(counts.sum()) != n_rows or (counts.size and (counts.min() < 0 or counts.max() > w_max)):
        raise ValueError("pseudo-label file: counts do not match the row block")
    if mmap:
        rows = np.memmap(path, dtype="<f4", mode="r", offset=off_rows, shape=(n_rows, n_fields))
    else:
        rows = np.fromfile(path, dtype="<f4", offset=off_rows, count=n_rows * n_fields).reshape(n_rows, n_fields)
    if not padded:
        return rows, torch.from_numpy(counts), meta
    out = np.zeros((n_images, w_max, n_fields), dtype=np.float32)
    valid = np.arange(w_max)[None, :] < counts[:, None]
    out[valid] = rows
    return torch.from_numpy(out), torch.from_numpy(counts), meta


def merge_record_files(paths, out_path, meta=None):
    """Concatenate per-rank files (rank-major = image order of `parallel.shard_range`) into one file --
    the offline counterpart of the all-gather.  w_max becomes the largest of the inputs."""
    parts = [load_records(p, padded=True) for p in paths]
    if not parts:
        raise ValueError("nothing to merge")
    w_max = max(int(r.shape[1]) for r, _, _ in parts)
    recs = []
    for r, _, _ in parts:
        pad = torch.zeros((r.shape[0], w_max, r.shape[2]), dtype=r.dtype)
        pad[:, : r.shape[1]] = r
        recs.append(pad)
    merged_meta = dict(meta or {})
    merged_meta.setdefault("shards", [m for _, _, m in parts])
    return save_records(out_path, torch.cat(recs), torch.cat([c for _, c, _ in parts]), merged_meta)


def records_to_boxlists(records, counts, image_sizes, device=None):
    """The student's view of stored pseudo-labels: one BoxList per image (mode xyxy) with the fields
    generate_pseudo_label attaches (reference st_generalized_rcnn.py:257-262): labels (int64), scores,
    plus region_idx and image_id.  image_sizes: (width, height) per image, or one pair for all."""
    records = records if isinstance(records, torch.Tensor) else torch.as_tensor(np.asarray(records))
    counts = counts if isinstance(counts, torch.Tensor) else torch.as_tensor(np.asarray(counts))
    if device is not None:
        records = records.to(device)
    n = records.shape[0]
    if len(image_sizes) == 2 and not isinstance(image_sizes[0], (tuple, list)):
        image_sizes = [tuple(image_sizes)] * n
    if len(image_sizes) != n:
        raise ValueError("need one (width, height) per image")
    out = []
    for i in range(n):
        k = int(counts[i])
        rows = records[i, :k]
        bl = BoxList(rows[:, 2:6].contiguous(), tuple(image_sizes[i]), mode="xyxy")
        bl.add_field("labels", rows[:, 1].round().to(torch.int64))
        bl.add_field("scores", rows[:, 6].contiguous())
        bl.add_field("region_idx", rows[:, 7].round().to(torch.int64))
        bl.add_field("image_id", rows[:, 0].round().to(torch.int64))
        out.append(bl)
    return out
