"""Pseudo-label record files: round trips, packed / memory-mapped reads, shard merging =
the all-gather order, the student's BoxList view, and loud failures on foreign / truncated files."""
import os
import struct

import numpy as np
import pytest
import torch

from cvpr22_cross_modal_pseudo_labeling_b200 import records as R
from cvpr22_cross_modal_pseudo_labeling_b200.parallel import shard_range


def _make(n_images, w_max, seed, first_id=0):
    g = torch.Generator().manual_seed(seed)
    rec = torch.zeros((n_images, w_max, 8))
    cnt = torch.randint(0, w_max + 1, (n_images,), generator=g, dtype=torch.int32)
    for i in range(n_images):
        k = int(cnt[i])
        rec[i, :k, 0] = first_id + i
        rec[i, :k, 1] = torch.randint(1, 66, (k,), generator=g).float()
        xy = torch.rand((k, 2), generator=g) * 600
        rec[i, :k, 2:4] = xy
        rec[i, :k, 4:6] = xy + torch.rand((k, 2), generator=g) * 200
        rec[i, :k, 6] = torch.rand((k,), generator=g)
        rec[i, :k, 7] = torch.randint(0, 1000, (k,), generator=g).float()
    return rec, cnt


def test_round_trip_padded_packed_and_mmap(tmp_path):
    rec, cnt = _make(7, 10, 1)
    p = str(tmp_path / "pl.b2pl")
    n_rows = R.save_records(p, rec, cnt, meta={"classes": ["__bg__", "cat"], "rank": 0})
    assert n_rows == int(cnt.sum()) and not os.path.exists(p + ".tmp")
    assert os.path.getsize(p) == 64 + 64 + 64 + 32 * n_rows          # header, meta, counts, rows
    rec2, cnt2, meta = R.load_records(p)
    assert torch.equal(rec2, rec) and torch.equal(cnt2, cnt) and rec2.dtype == torch.float32
    assert meta == {"classes": ["__bg__", "cat"], "rank": 0}
    rows, cnt3, _ = R.load_records(p, padded=False, mmap=True)
    assert isinstance(rows, np.memmap) and rows.shape == (n_rows, 8) and torch.equal(cnt3, cnt)
    off = np.concatenate([[0], np.cumsum(cnt.numpy())])
    for i in range(7):
        assert np.array_equal(rows[off[i]:off[i + 1]], rec[i, :int(cnt[i])].numpy())
    # empty sets and images without labels
    R.save_records(p, torch.zeros((0, 4, 8)), torch.zeros((0,), dtype=torch.int32))
    e, c, m = R.load_records(p)
    assert e.shape == (0, 4, 8) and c.numel() == 0 and m == {}
    R.save_records(p, torch.zeros((3, 4, 8)), torch.zeros((3,), dtype=torch.int32))
    assert R.load_records(p, padded=False)[0].shape == (0, 8)


def test_merge_of_rank_shards_equals_the_all_gather_order(tmp_path):
    n_img, ws = 11, 4
    full, cnt = _make(n_img, 6, 2)
    paths = []
    for r in range(ws):
        lo, hi = shard_range(n_img, r, ws)
        paths.append(str(tmp_path / ("shard%d.b2pl" % r)))
        R.save_records(paths[-1], full[lo:hi], cnt[lo:hi], {"rank": r})
    out = str(tmp_path / "all.b2pl")
    assert R.merge_record_files(paths, out, {"split": "train"}) == int(cnt.sum())
    rec, c, meta = R.load_records(out)
    assert torch.equal(rec, full) and torch.equal(c, cnt)
    assert meta["split"] == "train" and [m["rank"] for m in meta["shards"]] == list(range(ws))
    # shards written with different w_max are padded to the widest
    a, ca = _make(2, 3, 5)
    b, cb = _make(2, 9, 6)
    pa, pb = str(tmp_path / "a"), str(tmp_path / "b")
    R.save_records(pa, a, ca)
    R.save_records(pb, b, cb)
    R.merge_record_files([pa, pb], out)
    rec, c, _ = R.load_records(out)
    assert rec.shape == (4, 9, 8) and torch.equal(rec[:2, :3], a) and torch.equal(rec[2:], b)
    assert float(rec[:2, 3:].abs().sum()) == 0 and c.tolist() == ca.tolist() + cb.tolist()


def test_boxlist_view_matches_pack_records(tmp_path):
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling.pseudo_label import pack_records
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    g = torch.Generator().manual_seed(3)
    pls = []
    for n in (3, 0, 5):
        xy = torch.rand((n, 2), generator=g) * 300
        bl = BoxList(torch.cat([xy, xy + 50], 1), (640, 480))
        bl.add_field("labels", torch.randint(1, 66, (n,), generator=g))
        bl.add_field("scores", torch.rand((n,), generator=g))
        bl.add_field("region_idx", torch.randint(0, 1000, (n,), generator=g))
        pls.append(bl)
    rec, cnt = pack_records(pls, [10, 11, 12], 4)                      # the third image is cut to w_max
    assert cnt.tolist() == [3, 0, 4]
    p = str(tmp_path / "x.b2pl")
    R.save_records(p, rec, cnt)
    back = R.records_to_boxlists(*R.load_records(p)[:2], (640, 480))
    for bl, src, k, img in zip(back, pls, (3, 0, 4), (10, 11, 12)):
        assert len(bl) == k and bl.size == (640, 480) and bl.mode == "xyxy"
        assert torch.equal(bl.bbox, src.bbox[:k]) and torch.equal(bl.get_field("scores"), src.get_field("scores")[:k])
        assert torch.equal(bl.get_field("labels"), src.get_field("labels")[:k])
        assert torch.equal(bl.get_field("region_idx"), src.get_field("region_idx")[:k])
        assert bl.get_field("image_id").tolist() == [img] * k
    with pytest.raises(ValueError):
        R.records_to_boxlists(rec, cnt, [(640, 480)])


def test_bad_files_fail_loudly(tmp_path):
    rec, cnt = _make(4, 5, 4)
    p = str(tmp_path / "pl.b2pl")
    R.save_records(p, rec, cnt)
    raw = open(p, "rb").read()
    bad = str(tmp_path / "bad")
    open(bad, "wb").write(b"XXXX" + raw[4:])
    with pytest.raises(ValueError, match="magic"):
        R.load_records(bad)
    open(bad, "wb").write(raw[:4] + struct.pack("<I", 99) + raw[8:])
    with pytest.raises(ValueError, match="version"):
        R.load_records(bad)
    open(bad, "wb").write(raw[:-8])
    with pytest.raises(ValueError, match="size"):
        R.load_records(bad)
    open(bad, "wb").write(raw[:20])
    with pytest.raises(ValueError, match="truncated"):
        R.load_records(bad)
    with pytest.raises(ValueError):
        R.save_records(bad, rec[:, :, :7], cnt)
    with pytest.raises(ValueError):
        R.save_records(bad, rec, cnt + 9)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from cvpr22_cross_modal_pseudo_labeling_b200 import records as R
from cvpr22_cross_modal_pseudo_labeling_b200.parallel import all_gather_records, shard_range
dist.init_process_group("gloo", init_method="env://")
rank, ws = dist.get_rank(), dist.get_world_size()
n_img, w_max, out = 7, 5, os.environ["OUT_DIR"]
g = torch.Generator().manual_seed(99)                 # every rank draws the same full set, keeps its shard
full = torch.rand((n_img, w_max, 8), generator=g)
cnt = torch.randint(0, w_max + 1, (n_img,), generator=g, dtype=torch.int32)
full = full * (torch.arange(w_max)[None, :, None] < cnt[:, None, None])
lo, hi = shard_range(n_img)
R.save_records(os.path.join(out, "shard%d.b2pl" % rank), full[lo:hi], cnt[lo:hi], {"rank": rank})
allr, allc = all_gather_records(full[lo:hi].contiguous(), cnt[lo:hi].contiguous())
dist.barrier()
if rank == 0:
    R.merge_record_files([os.path.join(out, "shard%d.b2pl" % r) for r in range(ws)], os.path.join(out, "all.b2pl"))
    rec, c, meta = R.load_records(os.path.join(out, "all.b2pl"))
    assert torch.equal(rec, allr) and torch.equal(c, allc) and torch.equal(rec, full) and torch.equal(c, cnt)
    assert [m["rank"] for m in meta["shards"]] == list(range(ws))
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_shard_files_merge_to_what_the_all_gather_returns_gloo(tmp_path):
    """world_size 2 (gloo): every rank writes its shard file and joins the all-gather; the merged file
    equals the gathered records -- the wire format and the disk format are the same thing."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, REPO_ROOT=root, OUT_DIR=str(tmp_path), MASTER_ADDR="127.0.0.1", MASTER_PORT="29633",
               WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_ids_beyond_fp32_exactness_are_refused(tmp_path):
    rec, cnt = _make(2, 3, 9)
    cnt[:] = 3
    rec[0, 0, 0] = float(1 << 24)
    with pytest.raises(ValueError):
        R.save_records(str(tmp_path / "x.b2pl"), rec, cnt)
