"""CPU simulation: bytes into shared memory for the row-streaming RoIAlign when G RoIs of one
(image, level) share one row stream (union of tap rows, per-row union column span)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth

def level_of(r):
    w = r[:, 3] - r[:, 1] + 1; h = r[:, 4] - r[:, 2] + 1
    s = np.sqrt(w * h)
    l = np.floor(4 + np.log2(s / 224 + 1e-6))
    return (np.clip(l, 2, 5) - 2).astype(int)

def spans(r, lv, shapes):
    sc = np.array(synth.FPN_SCALES)[lv]
    H = np.array([s[0] for s in shapes])[lv]; W = np.array([s[1] for s in shapes])[lv]
    def ax(a, b, ext):
        st = a * sc; en = b * sc
        sz = np.maximum(en - st, 1.0); bin_ = sz / 7
        first = st + 0.25 * bin_; last = st + 6 * bin_ + 0.75 * bin_
        lo = np.clip(np.floor(np.maximum(first, 0)), 0, ext - 1)
        hi = np.clip(np.floor(np.maximum(last, 0)) + 1, 0, ext - 1)
        return lo.astype(int), np.maximum(hi, lo).astype(int)
    y0, y1 = ax(r[:, 2], r[:, 4], H); x0, x1 = ax(r[:, 1], r[:, 3], W)
    return y0, y1, x0, x1

def sim(rois, G, key):
    shapes = synth.fpn_shapes()
    lv = level_of(rois)
    y0, y1, x0, x1 = spans(rois, lv, shapes)
    base = ((y1 - y0 + 1) * (x1 - x0 + 1)).sum()
    img = rois[:, 0].astype(int)
    cx = (rois[:, 1] + rois[:, 3]) / 2; cy = (rois[:, 2] + rois[:, 4]) / 2
    k = key(img, lv, cx, cy, rois)
    order = np.argsort(k, kind="stable")
    tot = 0
    i = 0; n = len(order); ngroups = 0
    il = img * 8 + lv
    while i < n:
        j = i + 1
        while j < n and j - i < G and il[order[j]] == il[order[i]]:
            j += 1
        g = order[i:j]
        ymin, ymax = y0[g].min(), y1[g].max()
        lo = np.full(ymax - ymin + 1, 1 << 30); hi = np.full(ymax - ymin + 1, -1)
        for t in g:
            sl = slice(y0[t] - ymin, y1[t] - ymin + 1)
            lo[sl] = np.minimum(lo[sl], x0[t]); hi[sl] = np.maximum(hi[sl], x1[t])
        m = hi >= 0
        tot += (hi[m] - lo[m] + 1).sum()
        ngroups += 1
        i = j
    return base, tot, ngroups

def keys():
    def k_y(img, lv, cx, cy, r): return (img * 8 + lv) * 4096.0 + cy
    def mk_band(bh):
        def k(img, lv, cx, cy, r):
            # band height in image px scaled with level: bh feature px
            stride = 4 * (2 ** lv)
            band = np.floor(cy / (bh * stride))
            return ((img * 8 + lv) * 256 + band) * 4096.0 + cx
        return k
    def mk_band_tl(bh):
        def k(img, lv, cx, cy, r):
            stride = 4 * (2 ** lv)
            band = np.floor(r[:, 2] / (bh * stride))
            return ((img * 8 + lv) * 256 + band) * 4096.0 + r[:, 1]
        return k
    def k_morton(img, lv, cx, cy, r):
        stride = 4 * (2 ** lv)
        xi = (cx / (stride * 4)).astype(np.int64); yi = (cy / (stride * 4)).astype(np.int64)
        m = np.zeros_like(xi)
        for b in range(8):
            m |= ((xi >> b) & 1) << (2 * b); m |= ((yi >> b) & 1) << (2 * b + 1)
        return (img * 8 + lv) * 65536.0 + m
    return {"y": k_y, "band8,x": mk_band(8), "band12,x": mk_band(12), "band16,x": mk_band(16), "band24,x": mk_band(24),
            "tl band12": mk_band_tl(12), "morton": k_morton}

if __name__ == "__main__":
    rng = np.random.default_rng(1236)
    which = sys.argv[1] if len(sys.argv) > 1 else "micro"
    if which == "micro":
        rois = synth.make_rois(rng, 1000, 4)
    else:
        import oracle
        sys.path.insert(0, ROOT)
        import bench
        out = []
        for b in range(3):
            boxes, scores = bench.make_rpn_candidates(rng, 1)
            kb, ks, o = [], [], 0
            for L in bench.RPN_LENS:
                k = oracle.nms(boxes[o:o + L], scores[o:o + L], 0.7)[:1000]
                kb.append(boxes[o:o + L][k]); ks.append(scores[o:o + L][k]); o += L
            kb, ks = np.concatenate(kb), np.concatenate(ks)
            top = np.argsort(-ks, kind="stable")[:1000]
            out.append(np.concatenate([np.full((len(top), 1), b, np.float32), kb[top]], 1))
        rois = np.concatenate(out)
    lv = level_of(rois)
    print("level histogram", np.bincount(lv, minlength=4))
    y0, y1, x0, x1 = spans(rois, lv, synth.fpn_shapes())
    px = (y1 - y0 + 1) * (x1 - x0 + 1)
    print("mean patch px", px.mean(), "per level", [px[lv == l].mean() for l in range(4)], "share of bytes", [px[lv == l].sum() / px.sum() for l in range(4)])
    for name, k in keys().items():
        for G in (2, 3, 4, 8):
            base, tot, ng = sim(rois, G, k)
            print("%-10s G=%d  in-bytes ratio %.3f  groups %d" % (name, G, tot / base, ng))
