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
    pre = getattr(t, "b200_bf16", None)   # TensorCoreLinear hands its bf16 output along with the fp32 one
    if pre is not None and pre.shape == t.shape and pre.device == t.device:
        return pre
    return t.to(torch.bfloat16).contiguous()


def embed_match_softmax(A, E, score_thresh=0.05, want_probs=True, want_logits=False, want_top=True):
    """A [R,D], E [C,D] (row 0 = background; any C: wider than 512 runs by column blocks) -> dict with any of
    probs [R,C] fp32 (row softmax), logits [R,C] fp32, top_label [R] int32 (best class >= 1,
    0 if its probability <= score_thresh), top_prob [R] fp32."""
    A, E = _bf16(A, "A"), _bf16(E, "E")
    r, d = A.shape
    c = E.shape[0]
    if E.shape[1] != d:
        raise ValueError("A and E must share the embedding dimension")
    dev = A.device
    out = {}
    wide = c > MAX_COLS   # b200_embed_match_wide: statistics pass + probability pass, no logits round trip
    probs = torch.empty((r, c), dtype=torch.float32, device=dev) if want_probs else None
    logits = torch.empty((r, c), dtype=torch.float32, device=dev) if want_logits else None
    top_label = torch.empty((r,), dtype=torch.int32, device=dev) if want_top else None
    top_prob = torch.empty((r,), dtype=torch.float32, device=dev) if want_top else None
    if r > 0 and c > 0 and wide:
        with torch.cuda.device(dev):
            rc = _ext.lib().b200_embed_match_wide(_ext.ptr(A), _ext.ptr(E), r, c, d, float(score_thresh),
                                                  _ext.ptr(probs), _ext.ptr(logits), _ext.ptr(top_label),
                                                  _ext.ptr(top_prob), _ext.stream_ptr(dev))
        _ext.check(rc, "b200_embed_match_wide")
    elif r > 0 and c > 0:
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
    """cls_logit = cls_emb . E^T (reference roi_box_predictors.py:67, an einsum that differentiates
    into both operands).  The gradient w.r.t. cls_emb is always produced; the gradient w.r.t. the class
    matrix is produced when E requires grad -- in the reference it does once exemplars are loaded:
    `set_class_embeddings(self.combine_embs(...))` (st_generalized_rcnn.py:173, :372-375) makes E a
    function of the learnable `lambda_exemplar`."""

    @staticmethod
    def forward(ctx, cls_emb, E):
        ctx.save_for_backward(cls_emb, E)
        return embed_match_softmax(cls_emb, E, want_probs=False, want_logits=True, want_top=False)["logits"]

    @staticmethod
    def backward(ctx, grad_logits):
        cls_emb, E = ctx.saved_tensors
        g_emb = g_E = None
        if ctx.needs_input_grad[0] and cls_emb.dtype == torch.float32:
            g_emb = grad_logits.float() @ E.float()        # fp32 training keeps the fp32 product
        elif ctx.needs_input_grad[0]:
            # bf16 (AMP / config #4): grad_emb [R, D] = g [R, C] . E [C, D] on the tensor cores; the contraction
            # runs over the C classes, padded to a multiple of 8 (16-byte rows for TMA)
            from .linear import linear_bf16
            c = grad_logits.shape[1]
            pad = (-c) % 8
            g16 = grad_logits.to(torch.bfloat16)
            et = E.detach().to(torch.bfloat16).t()
            if pad:
                g16 = torch.nn.functional.pad(g16, (0, pad))
                et = torch.nn.functional.pad(et, (0, pad))
            g_emb = linear_bf16(g16.contiguous(), et.contiguous(), None)[0].to(cls_emb.dtype)
        if ctx.needs_input_grad[1]:
            g_E = (grad_logits.float().t() @ cls_emb.float()).to(E.dtype)
        return g_emb, g_E


def embed_logits(cls_emb, E):
    return _EmbedLogits.apply(cls_emb, E)


_plan_cache = {}


def _align_plan(rows_per_image, wcounts, dev):
    """Index tensors of caption_align for one (rows, words) signature, built once: groups of
    consecutive images with at most MAX_COLS words, and per group the image id of every row /
    word and the first row of every image.  Cached so that steady-state calls issue no host to
    device copies (and can be captured in a CUDA graph)."""
    key = (tuple(rows_per_image), tuple(wcounts), str(dev))
    plan = _plan_cache.get(key)
    if plan is not None:
        return plan
    nimg = len(rows_per_image)
    row_start = [0]
    for n in rows_per_image:
        row_start.append(row_start[-1] + int(n))
    groups, cur, cur_w = [], [], 0
    for i in range(nimg):
        w = wcounts[i]
        if w > MAX_COLS:
            raise ValueError("more than %d words in one caption" % MAX_COLS)
        if cur and cur_w + w > MAX_COLS:
            groups.append(cur)
            cur, cur_w = [], 0
        cur.append(i)
        cur_w += w
    if cur:
        groups.append(cur)
    plan = []
    for g in groups:
        wc = [wcounts[i] for i in g]
        wt = sum(wc)
        r0, r1 = row_start[g[0]], row_start[g[-1] + 1]
        entry = dict(images=g, wc=wc, wt=wt, r0=r0, r1=r1)
        if wt > 0 and r1 > r0:
            ar = torch.arange(len(g), dtype=torch.int32, device=dev)
            entry["row_seg"] = torch.repeat_interleave(ar, torch.tensor([rows_per_image[i] for i in g], device=dev),
                                                       output_size=r1 - r0)
            entry["col_seg"] = torch.repeat_interleave(ar, torch.tensor(wc, device=dev), output_size=wt)
            entry["seg_start"] = torch.tensor([row_start[i] - r0 for i in g], dtype=torch.int32, device=dev)
        plan.append(entry)
    if len(_plan_cache) > 64:
        _plan_cache.clear()
    _plan_cache[key] = plan
    return plan


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
    plan = _align_plan([int(n) for n in rows_per_image], [int(w.shape[0]) for w in word_embs], dev)
    lib = _ext.lib()
    for e in plan:
        g, wc, wt, r0, r1 = e["images"], e["wc"], e["wt"], e["r0"], e["r1"]
        if wt == 0 or r1 == r0:
            for i, w in zip(g, wc):
                results[i] = (torch.full((w,), -1, dtype=torch.int64, device=dev),
                              torch.full((w,), float("-inf"), device=dev), torch.zeros((w,), device=dev))
            continue
        parts = [_bf16(word_embs[i], "word_embs") for i in g if word_embs[i].shape[0] > 0]
        E = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        best = torch.zeros((wt,), dtype=torch.int64, device=dev)   # uint64 keys
        ridx = torch.empty((wt,), dtype=torch.int32, device=dev)
        mx = torch.empty((wt,), dtype=torch.float32, device=dev)
        sg = torch.empty((wt,), dtype=torch.float32, device=dev)
        Ag = A[r0:r1]
        with torch.cuda.device(dev):
            rc = lib.b200_embed_match(_ext.ptr(Ag), _ext.ptr(E), r1 - r0, wt, d, _ext.B200_MATCH_COLMAX, 0.0,
                                      None, None, None, None, _ext.ptr(e["row_seg"]), _ext.ptr(e["col_seg"]),
                                      _ext.ptr(e["seg_start"]), _ext.ptr(best), _ext.stream_ptr(dev))
            _ext.check(rc, "b200_embed_match")
            rc = lib.b200_colmax_decode(_ext.ptr(best), wt, _ext.ptr(ridx), _ext.ptr(mx), _ext.ptr(sg),
                                        _ext.stream_ptr(dev))
            _ext.check(rc, "b200_colmax_decode")
        ridx64 = ridx.to(torch.int64)
        o = 0
        for i, w in zip(g, wc):
            results[i] = (ridx64[o:o + w], mx[o:o + w], sg[o:o + w])
            o += w
    return results
