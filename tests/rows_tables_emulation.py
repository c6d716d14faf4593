"""CPU emulation of the table builder + consumer arithmetic of csrc/roi_align_fwd_rows.cu.

Transcribes build_rows_tab() lane by lane (numpy, fp32) and evaluates the x pass / y pass exactly
as the consumer threads do, then compares with the C oracle (oracle.pooler_forward).  It checks the
ALGORITHM (list construction, tap merging, ring entries and their grouping, runs, ring placement), not the CUDA transcription; the
GPU parity tests do that.  Test infrastructure only (imports oracle/).
    python tests/rows_tables_emulation.py [n_rois]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import synth  # noqa: E402

f32 = np.float32
P, NS = 7, 14


def axis_sample(start, p, binsz, i, grid, extent):
    pos = f32(start + f32(f32(p) * binsz))
    pos = f32(pos + f32(f32(f32(f32(i) + f32(0.5)) * binsz) / f32(grid)))
    ok = not (pos < f32(-1.0) or pos > f32(extent))
    if pos <= 0:
        pos = f32(0)
    lo = int(pos)
    if lo >= extent - 1:
        hi = lo = extent - 1
        pos = f32(lo)
    else:
        hi = lo + 1
    l = f32(pos - f32(lo))
    h = f32(f32(1.0) - l)
    return lo, hi, l, h, ok


def build_axis(start, binsz, extent):
    """one half-warp of build_rows_tab: returns (list coords, per-bin merged taps [(idx, w)])"""
    lo = [0] * 16; hi = [0] * 16; wl = [f32(0)] * 16; wh = [f32(0)] * 16
    for hl in range(NS):
        a, b, l, h, ok = axis_sample(start, hl >> 1, binsz, hl & 1, 2, extent)
        if ok:
            lo[hl], hi[hl], wl[hl], wh[hl] = a, b, l, h
    nnew = [0] * 16
    for hl in range(16):
        n = 1 if lo[hl] == hi[hl] else 2
        if hl > 0:
            if lo[hl] == lo[hl - 1] and hi[hl] == hi[hl - 1]:
                n = 0
            elif lo[hl] == hi[hl - 1]:
                n = 1
        if hl >= NS:
            n = 0
        nnew[hl] = n
    scan = list(np.cumsum(nnew))
    nlist = scan[15]
    coords = [None] * nlist
    for hl in range(NS):
        if nnew[hl] >= 1:
            coords[scan[hl] - 1] = hi[hl]
        if nnew[hl] == 2:
            coords[scan[hl] - 2] = lo[hl]
    jhi = [scan[hl] - 1 for hl in range(16)]
    jlo = [jhi[hl] if lo[hl] == hi[hl] else jhi[hl] - 1 for hl in range(16)]
    w_lo = [f32(0) if lo[hl] == hi[hl] else wh[hl] for hl in range(16)]
    w_hi = [f32(wl[hl] + wh[hl]) if lo[hl] == hi[hl] else wl[hl] for hl in range(16)]
    bins = []
    for b in range(P):
        sa, sb = 2 * b, 2 * b + 1
        idx = [jlo[sa], jhi[sa], jlo[sb], jhi[sb]]
        w = [w_lo[sa], w_hi[sa], w_lo[sb], w_hi[sb]]
        for k in range(1, 4):
            for j in range(k):
                if idx[k] == idx[j]:
                    w[j] = f32(w[j] + w[k])
                    w[k] = f32(0)
        bins.append([(idx[k], w[k]) for k in range(4) if w[k] != 0])
    assert all(c is not None for c in coords)
    return coords, bins


def runs_of(cols):
    runs = []
    for j, x in enumerate(cols):
        if j == 0 or x != cols[j - 1] + 1:
            runs.append([j, x, 1])
        else:
            runs[-1][2] += 1
    return runs


def emulate_roi(feats, roi, scales, k_min, k_max, stats):
    """feats: list of [B, H, W, C] fp32 arrays (NHWC)"""
    b = int(roi[0])
    x1, y1, x2, y2 = [f32(v) for v in roi[1:]]
    lvl = int(oracle.level_map(np.asarray([roi], dtype=np.float32), k_min, k_max)[0]) if len(feats) > 1 else 0
    C = feats[0].shape[-1]
    if lvl < 0:
        return np.zeros((C, P, P), np.float32), lvl
    f = feats[lvl]
    H, W = f.shape[1], f.shape[2]
    sc = f32(scales[lvl])
    sw, sh = f32(x1 * sc), f32(y1 * sc)
    rw = max(f32(f32(x2 * sc) - sw), f32(1.0))
    rh = max(f32(f32(y2 * sc) - sh), f32(1.0))
    bh, bw = f32(rh / f32(P)), f32(rw / f32(P))
    rows, ybins = build_axis(sh, bh, H)
    cols, xbins = build_axis(sw, bw, W)
    # slot image of every list row: the runs copied to compact positions
    runs = runs_of(cols)
    assert sum(r[2] for r in runs) == len(cols) and len(runs) <= 28
    stats["runs"] += len(runs); stats["rows"] += len(rows); stats["cols"] += len(cols)
    wy = np.zeros((len(rows), 8), np.float32)
    for ph, taps in enumerate(ybins):
        for idx, w in taps:
            assert wy[idx, ph] == 0
            wy[idx, ph] = f32(f32(0.25) * w)
    # ring entries: per list row the output rows it feeds, paired greedily from the lowest one
    # (b, b + 1); the pairs of all rows ordered by b (the kernel's counting sort is stable in the row)
    entries = []
    for i in range(len(rows)):
        left = {ph for ph in range(P) if wy[i, ph] != 0}
        for bb in range(P):
            if bb in left:
                entries.append((bb, i, wy[i, bb], wy[i, bb + 1] if bb + 1 < P else f32(0)))
                left.discard(bb)
                left.discard(bb + 1)
    entries.sort(key=lambda e: (e[0], e[1]))
    assert len(entries) <= 28
    stats["entries"] = stats.get("entries", 0) + len(entries)
    stats["generic"] += sum(1 for i in range(len(rows)) if sum(1 for e in entries if e[1] == i) > 1)
    acc = np.zeros((P, P, C), np.float32)
    for bb, i, w0, w1 in entries:  # the consumers: one ring entry = one copy of tap row i
        y = rows[i]
        slot = np.zeros((28, C), np.float32)
        for pos, col, ln in runs:
            slot[pos:pos + ln] = f[b, y, col:col + ln]
        for pw, taps in enumerate(xbins):
            u = np.zeros(C, np.float32)
            for idx, w in taps:
                u = (u + w * slot[idx]).astype(np.float32)
            acc[bb, pw] += w0 * u
            if bb + 1 < P:
                acc[bb + 1, pw] += w1 * u
    return acc.transpose(2, 0, 1), lvl


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(7)
    B, C = 2, 8
    shapes = synth.fpn_shapes(400, 672)
    feats_nchw = [rng.standard_normal((B, C, h, w)).astype(np.float32) for (h, w) in shapes]
    feats_nhwc = [np.ascontiguousarray(f.transpose(0, 2, 3, 1)) for f in feats_nchw]
    rois = synth.make_rois(rng, n, B, 672, 400, smin=2, smax=700, degenerate=0.05)
    # extra edge cases: outside the image, tiny, huge, touching borders, malformed
    extra = np.array([
        [0, -50, -50, -20, -20], [0, 700, 450, 800, 500], [1, 660, 390, 690, 420], [0, -30, 100, 40, 180],
        [1, 100, 100, 100.5, 100.5], [0, 0, 0, 671, 399], [1, 300, 200, 290, 190], [0, 10.3, 20.7, 12.1, 22.2],
        [0, -200, -200, 900, 700], [1, 640, 0, 900, 30], [0, 0, 380, 30, 600], [1, 5, 5, 9, 300],
    ], dtype=np.float32)
    rois = np.concatenate([rois, extra]).astype(np.float32)
    scales = synth.FPN_SCALES
    k_min, k_max = -np.log2(scales[0]), -np.log2(scales[-1])
    want, wl = oracle.pooler_forward(feats_nchw, rois, scales, P, P, 2)
    stats = dict(runs=0, rows=0, cols=0, generic=0)
    worst = 0.0
    for i, roi in enumerate(rois):
        got, lvl = emulate_roi(feats_nhwc, roi, scales, k_min, k_max, stats)
        assert lvl == wl[i], (i, lvl, wl[i])
        err = np.abs(got - want[i]).max() / max(1e-6, np.abs(want[i]).max())
        assert np.allclose(got, want[i], rtol=1e-5, atol=2e-6), (i, roi, err)
        worst = max(worst, err)
    m = len(rois)
    print("ok: %d RoIs, worst rel err %.2e; per RoI: rows %.1f (ring entries %.1f) cols %.1f runs %.2f; rows copied more than once %d" %
          (m, worst, stats["rows"] / m, stats["entries"] / m, stats["cols"] / m, stats["runs"] / m, stats["generic"]))
    # single-level pooler with wide RoIs (sparse taps -> many runs)
    f1 = [feats_nchw[1]]
    big = synth.make_rois(rng, 60, B, 672, 400, smin=100, smax=700, degenerate=0.0)
    want1, _ = oracle.pooler_forward(f1, big, (scales[1],), P, P, 2)
    stats = dict(runs=0, rows=0, cols=0, generic=0)
    for i, roi in enumerate(big):
        got, _ = emulate_roi([feats_nhwc[1]], roi, (scales[1],), 0, 0, stats)
        assert np.allclose(got, want1[i], rtol=1e-5, atol=2e-6), (i, roi)
    print("ok: single level, %d wide RoIs; runs per RoI %.1f" % (len(big), stats["runs"] / len(big)))




def check_ring_planner(n_rois=3000, seed=3):
    """The planner's byte-ring placement (build_rows_tab, 'ring placement'): no entry may overwrite an
    earlier entry that is not guaranteed to have been given back (sequence number >= dep), entries stay
    inside the ring, and dep - 1 >= g - kNBar (the parity of the empty barrier is then unambiguous)."""
    RING, NBAR, HIST = 168, 16, 64
    rng = np.random.default_rng(seed)
    g0, head = 0, 0
    hist_place = [0] * HIST
    hist_size = [0] * HIST
    placed = []  # (place, size) by sequence number
    for _ in range(n_rois):
        s = int(rng.integers(1, 29))
        nent = int(rng.integers(0, 29))
        k0, kfit = (RING - head) // s, RING // s
        places = []
        for i in range(nent):
            place = head + i * s if i < k0 else ((i - k0) % kfit) * s
            places.append(place)
            hist_place[(g0 + i) % HIST] = place
            hist_size[(g0 + i) % HIST] = s
        for i in range(nent):
            g = g0 + i
            place = places[i]
            dep = g - NBAR + 1 if g >= NBAR else 0
            for d in range(1, NBAR):
                if d > g:
                    break
                e = g - d
                pe, se = hist_place[e % HIST], hist_size[e % HIST]
                if pe < place + s and place < pe + se:
                    dep = max(dep, e + 1)
                    break
            assert 0 <= place and place + s <= RING
            assert dep <= g and (dep == 0 or dep - 1 >= g - NBAR)
            for e in range(dep, g):
                pe, se = placed[e]
                assert not (pe < place + s and place < pe + se), (g, e)
            placed.append((place, s))
        if nent:
            head = places[-1] + s
        g0 += nent
    print("ok: ring planner, %d entries" % len(placed))


if __name__ == "__main__":
    main()
    check_ring_planner()
