"""Region -> class-embedding scoring (C ABI b200_embed_match: tcgen05 bf16 GEMM, fused epilogue).

Replaces, in the reference:
  roi_box_predictors.py:67   cls_logit = einsum('pe,ce->pc', cls_emb, cls_score)
  box_head/inference.py:62   class_prob = softmax(class_logits)
  st_generalized_rcnn.py:245-255   region_scores = einsum('pd,wd->pw'); max over regions; sigmoid
"""
import torch
from torch.autograd import Function

from .. import _ext

MAX_COLS = 512


def _bf16(t, name):
    _ext.require_cuda(t, name)
    if t.dim() != 2:
        raise ValueError("%s must be 2-D" % name)
    return t.to(torch.bfloat16).contiguous()


def embed_match_softmax(A, E, score_thresh=0.05, want_probs=True, want_logits=False, want_top=True):
    """A [R,D], E [C,D] (row 0 = background) -> dict with any of
    probs [R,C] fp32 (row softmax), logits [R,C] fp32, top_label [R] int32 (best class >= 1,
    0 if its probability <= score_thresh), top_prob [R] fp32."""
    A, E = _bf16(A, "A"), _bf16(E, "E")
    r, d = A.shape
    c = E.shape[0]
    if E.shape[1] != d:
        raise ValueError("A and E must share the embedding dimension")
    if c > MAX_COLS:
        raise ValueError("at most %d classes per call (got %d)" % (MAX_COLS, c))
    dev = A.device
    out = {}
    probs = torch.empty((r, c), dtype=torch.float32, device=dev) if want_probs else None
    logits = torch.empty((r, c), dtype=torch.float32, device=dev) if want_logits else None
    top_label = torch.empty((r,), dtype=torch.int32, device=dev) if want_top else None
    top_prob = torch.empty((r,), dtype=torch.float32, device=dev) if want_top else None
    if r > 0 and c > 0:
        with torch.cuda.device(dev):
            rc = _ext.lib().b200_embed_match(_ext.ptr(A), _ext.ptr(E), r, c, d, _ext.B200_MATCH_SOFTMAX,
                                             float(score_thresh), _ext.ptr(probs), _ext.ptr(logits),
                                             _ext.ptr(top_label), _ext.ptr(top_prob), None, None, None, None,
                                             _ext.stream_ptr(dev))
        _ext.check(rc, "b200_embed_match")
    for k, v in (("probs", probs), ("logits", logits), ("top_label", top_label), ("top_prob", top_prob)):
        if v is not None:
            out[k] = v
    return out


class _EmbedLogits(Function):
    """cls_logit = cls_emb . E^T with gradient w.r.t. cls_emb (the class matrix is a constant,
    reference roi_box_predictors.py:84-92)."""

    @staticmethod
    def forward(ctx, cls_emb, E):
        ctx.save_for_backward(E)
        ctx.in_dtype = cls_emb.dtype
        return embed_match_softmax(cls_emb, E, want_probs=False, want_logits=True, want_top=False)["logits"]

    @staticmethod
    def backward(ctx, grad_logits):
        (E,) = ctx.saved_tensors
        return (grad_logits.float() @ E.float()).to(ctx.in_dtype), None


def embed_logits(cls_emb, E):
    return _EmbedLogits.apply(cls_emb, E)


def caption_align(emb, rows_per_image, word_embs):
    """Caption-noun alignment for a batch (reference st_generalized_rcnn.py:243-255).

    emb [R,D]: projected region embeddings of all images, image-major;
    rows_per_image: list[int]; word_embs: list of [W_i, D] tensors (may be empty).
    Returns per image (region_idx int64 [W_i] local to the image, max_score [W_i], sigmoid [W_i]).
    """
    A = _bf16(emb, "emb")
    dev = A.device
    nimg = len(rows_per_image)
    assert len(word_embs) == nimg
    d = A.shape[1]
    results = [None] * nimg
    row_start = [0]
    for n in rows_per_image:
        row_start.append(row_start[-1] + int(n))
    # group consecutive images so that each call sees at most MAX_COLS words
    groups, cur, cur_w = [], [], 0
    for i in range(nimg):
        w = int(word_embs[i].shape[0])
        if w > MAX_COLS:
            raise ValueError("more than %d words in one caption" % MAX_COLS)
        if cur and cur_w + w > MAX_COLS:
            groups.append(cur)
            cur, cur_w = [], 0
        cur.append(i)
        cur_w += w
    if cur:
        groups.append(cur)
    lib = _ext.lib()
    for g in groups:
        wcounts = [int(word_embs[i].shape[0]) for i in g]
        wt = sum(wcounts)
        r0, r1 = row_start[g[0]], row_start[g[-1] + 1]
        if wt == 0 or r1 == r0:
            for i in g:
                w = int(word_embs[i].shape[0])
                results[i] = (torch.full((w,), -1, dtype=torch.int64, device=dev),
                              torch.full((w,), float("-inf"), device=dev), torch.zeros((w,), device=dev))
            continue
        E = torch.cat([_bf16(word_embs[i], "word_embs") for i in g if word_embs[i].shape[0] > 0], dim=0)
        rcounts = torch.tensor([rows_per_image[i] for i in g], device=dev)
        row_seg = torch.repeat_interleave(torch.arange(len(g), dtype=torch.int32, device=dev), rcounts,
                                          output_size=r1 - r0)
        col_seg = torch.repeat_interleave(torch.arange(len(g), dtype=torch.int32, device=dev),
                                          torch.tensor(wcounts, device=dev), output_size=wt)
        seg_start = torch.tensor([row_start[i] - r0 for i in g], dtype=torch.int32, device=dev)
        best = torch.zeros((wt,), dtype=torch.int64, device=dev)   # uint64 keys
        ridx = torch.empty((wt,), dtype=torch.int32, device=dev)
        mx = torch.empty((wt,), dtype=torch.float32, device=dev)
        sg = torch.empty((wt,), dtype=torch.float32, device=dev)
        Ag = A[r0:r1]
        with torch.cuda.device(dev):
            rc = lib.b200_embed_match(_ext.ptr(Ag), _ext.ptr(E), r1 - r0, wt, d, _ext.B200_MATCH_COLMAX, 0.0,
                                      None, None, None, None, _ext.ptr(row_seg), _ext.ptr(col_seg),
                                      _ext.ptr(seg_start), _ext.ptr(best), _ext.stream_ptr(dev))
            _ext.check(rc, "b200_embed_match")
            rc = lib.b200_colmax_decode(_ext.ptr(best), wt, _ext.ptr(ridx), _ext.ptr(mx), _ext.ptr(sg),
                                        _ext.stream_ptr(dev))
            _ext.check(rc, "b200_colmax_decode")
        o = 0
        for i, w in zip(g, wcounts):
            results[i] = (ridx[o:o + w].to(torch.int64), mx[o:o + w], sg[o:o + w])
            o += w
    return results
