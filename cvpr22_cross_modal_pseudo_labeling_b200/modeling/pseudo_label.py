"""Caption-guided pseudo-labeling of region proposals.

Mirrors the alignment half of STGeneralizedRCNN.generate_pseudo_label (reference
modeling/detector/st_generalized_rcnn.py:218-275): for every image, every caption noun is
aligned to the region whose projected embedding scores highest against the noun's word
embedding; the aligned (teacher-refined) box becomes a pseudo ground-truth box with the noun's
label id and score sigmoid(max score).  The einsum / max / sigmoid chain (:245-255) is one
tcgen05 launch per batch (layers.caption_align) instead of a Python loop over images.
"""
import torch

from ..layers import caption_align
from ..structures import BoxList


def generate_pseudo_labels(cls_embs, results, word_embs, ids_cap, words=None):
    """
    cls_embs : [sum(P_i), D] projected region embeddings (emb_pred output), image-major
    results  : list[BoxList] teacher-refined boxes per image, len(results[i]) == P_i
               (PostProcessor(is_teacher=True) with one dummy class, reference :220-223)
    word_embs: list of [W_i, D] L2-normalised noun embeddings (reference extract_emb :202-209)
    ids_cap  : list of int64 [W_i] label ids of the nouns (reference :257)
    Returns list[BoxList] with fields labels, scores, consistencies, embs (and joined_words).
    """
    rows = [len(r) for r in results]
    aligned = caption_align(cls_embs, rows, word_embs)
    out, o = [], 0
    for i, (res, (idx, _mx, sig)) in enumerate(zip(results, aligned)):
        w = int(word_embs[i].shape[0])
        if w == 0 or rows[i] == 0:  # no noun phrase found (reference :236-239)
            out.append(BoxList(torch.zeros((0, 4), device=cls_embs.device), res.size, mode=res.mode))
            o += rows[i]
            continue
        pl = res[idx]
        if words is not None:
            pl.add_field("joined_words", "/".join(words[i]))
        pl.add_field("labels", ids_cap[i])
        pl.add_field("scores", sig)
        pl.add_field("consistencies", torch.ones_like(sig))
        pl.add_field("embs", cls_embs[o + idx])
        pl.add_field("region_idx", idx)
        out.append(pl)
        o += rows[i]
    return out


def pack_records(pseudo_labels, image_ids, w_max):
    """Fixed-size wire format of the collected pseudo-labels: a padded
    [B, w_max, 8] fp32 tensor (img_id, label, x1, y1, x2, y2, score, region_idx) and a
    [B] int32 count -- what the ranks all-gather over NCCL."""
    dev = pseudo_labels[0].bbox.device if pseudo_labels else torch.device("cuda")
    if any(abs(int(i)) >= (1 << 24) for i in image_ids):
        raise ValueError("image ids must be below 2^24 (the record stores them as fp32)")
    # one concatenation + one indexed store for the whole batch (no per-image device writes: at 64 images per
    # GPU a loop of scalar assignments costs more than the kernels that produced the labels)
    ns = [min(len(pl), w_max) for pl in pseudo_labels]
    rec = torch.zeros((len(pseudo_labels), w_max, 8), dtype=torch.float32, device=dev)
    cnt = torch.tensor(ns, dtype=torch.int32, device=dev)
    rows, img_idx, slot_idx = [], [], []
    for i, (pl, n) in enumerate(zip(pseudo_labels, ns)):
        if n == 0:
            continue
        region = pl.get_field("region_idx")[:n].float() if pl.has_field("region_idx") else torch.full((n,), -1.0, device=dev)
        rows.append(torch.cat([torch.full((n, 1), float(image_ids[i]), device=dev), pl.get_field("labels")[:n].float()[:, None],
                               pl.convert("xyxy").bbox[:n], pl.get_field("scores")[:n][:, None], region[:, None]], dim=1))
        img_idx += [i] * n
        slot_idx += list(range(n))
    if rows:
        rec[torch.tensor(img_idx, device=dev), torch.tensor(slot_idx, device=dev)] = torch.cat(rows)
    return rec, cnt
