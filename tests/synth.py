"""Seeded synthetic inputs shared by the parity tests and bench.py."""
import numpy as np

FPN_SCALES = (0.25, 0.125, 0.0625, 0.03125)
IMG_W, IMG_H = 1333, 800


def fpn_shapes(img_h=IMG_H, img_w=IMG_W, n_levels=4, divisibility=32):
    """P2.. feature map sizes of an image padded to a multiple of `divisibility`."""
    ph = (img_h + divisibility - 1) // divisibility * divisibility
    pw = (img_w + divisibility - 1) // divisibility * divisibility
    return [(ph // (4 << l), pw // (4 << l)) for l in range(n_levels)]


def make_rois(rng, n, batch, img_w=IMG_W, img_h=IMG_H, smin=16.0, smax=700.0, degenerate=0.01):
    """[n*batch, 5] fp32 rois: sqrt(area) log-uniform in [smin, smax], aspect log-uniform
    [1/2, 2], centre uniform, clipped to the image; ~1% degenerate (sub-pixel / border)."""
    out = []
    for b in range(batch):
        s = np.exp(rng.uniform(np.log(smin), np.log(smax), n))
        ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
        w, h = s * np.sqrt(ar), s / np.sqrt(ar)
        cx, cy = rng.uniform(0, img_w, n), rng.uniform(0, img_h, n)
        x1, y1, x2, y2 = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2
        nd = int(n * degenerate)
        if nd:
            idx = rng.choice(n, nd, replace=False)
            x2[idx[: nd // 2]] = x1[idx[: nd // 2]] + rng.uniform(0, 1, nd // 2)       # sub-pixel wide
            y1[idx[nd // 2:]] = 0.0                                                   # touching the border
        x1, x2 = np.clip(x1, 0, img_w - 1), np.clip(x2, 0, img_w - 1)
        y1, y2 = np.clip(y1, 0, img_h - 1), np.clip(y2, 0, img_h - 1)
        out.append(np.stack([np.full(n, b), x1, y1, x2, y2], 1))
    return np.concatenate(out, 0).astype(np.float32)


def make_nms_boxes(rng, n, img_w=IMG_W, img_h=IMG_H, size=64.0, jitter=0.35, clusters=None):
    """n boxes around `clusters` centres (heavy overlap, like RPN output), unique scores
    sorted descending when sort=True is wanted by the caller."""
    clusters = clusters or max(1, n // 12)
    ccx, ccy = rng.uniform(0, img_w, clusters), rng.uniform(0, img_h, clusters)
    cs = np.exp(rng.uniform(np.log(size / 2), np.log(size * 4), clusters))
    k = rng.integers(0, clusters, n)
    s = cs[k] * np.exp(rng.normal(0, jitter, n))
    ar = np.exp(rng.normal(0, jitter, n))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    cx = ccx[k] + rng.normal(0, 1, n) * cs[k] * jitter
    cy = ccy[k] + rng.normal(0, 1, n) * cs[k] * jitter
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h - 1)
    scores = (rng.permutation(n).astype(np.float64) + rng.uniform(0.1, 0.9, n)) / n   # unique
    return b.astype(np.float32), scores.astype(np.float32)
