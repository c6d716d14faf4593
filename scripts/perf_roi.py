"""Kernel-level timing sweep on one B200 (development aid; bench.py is the contract)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvpr22_cross_modal_pseudo_labeling_b200 import _ext  # noqa: E402
from cvpr22_cross_modal_pseudo_labeling_b200.layers import nms_batched  # noqa: E402
from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward, _backward  # noqa: E402
from tests import synth  # noqa: E402


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--rois", type=int, default=1000)
    ap.add_argument("--channels", type=int, default=256)
    ap.add_argument("--what", default="fwd,bwd,nms,layout")
    args = ap.parse_args()
    what = set(args.what.split(","))
    B, n, C = args.batch, args.rois, args.channels
    rng = np.random.default_rng(1236)
    shapes = synth.fpn_shapes()
    res = {}
    g = torch.Generator(device="cuda").manual_seed(1236)
    feats_nchw = [torch.randn((B, C, h, w), device="cuda", generator=g) for (h, w) in shapes]
    feats = [f.contiguous(memory_format=torch.channels_last) for f in feats_nchw]
    rois = torch.from_numpy(synth.make_rois(rng, n, B)).cuda()
    R = rois.shape[0]
    pyr_bytes = sum(f.numel() * 4 for f in feats)
    print("features %.1f MB, R=%d" % (pyr_bytes / 1e6, R), flush=True)

    if "fwd" in what:
        for (res_, name) in ((7, "box7"), (14, "mask14")):
            out_bytes = R * C * res_ * res_ * 4
            algo = pyr_bytes + out_bytes + R * 20
            for exact in (True, False):
                # fast math: variant 0 = the default (box 7x7 at 256 channels: the row-streaming kernel),
                # 16 + v = the separable marching kernel with occupancy variant v
                for pk in ((0, 1, 2) if exact else (0, 16, 17, 18)):
                    _ext.debug_set(False, exact, pk)
                    med, best = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (res_, res_), 2))
                    tag = "v%d" % pk if pk < 16 else "sep_v%d" % (pk - 16)
                    if not exact and pk == 0 and res_ == 7 and C == 256:
                        tag = "rows"
                    key = "fwd_%s_%s_%s" % (name, "exact" if exact else "fma", tag)
                    res[key] = dict(ms=med, best_ms=best, GBs=algo / med / 1e6, mroi_s=R / med / 1e3)
                    print(key, res[key], flush=True)
            _ext.debug_set(True, True, 0)
            med, best = timeit(lambda: _forward(feats, synth.FPN_SCALES, rois, (res_, res_), 2), iters=5)
            res["fwd_%s_generic_nhwc" % name] = dict(ms=med, GBs=algo / med / 1e6)
            print("fwd_%s_generic_nhwc" % name, res["fwd_%s_generic_nhwc" % name], flush=True)
            med, best = timeit(lambda: _forward(feats_nchw, synth.FPN_SCALES, rois, (res_, res_), 2), iters=5)
            res["fwd_%s_generic_nchw" % name] = dict(ms=med, GBs=algo / med / 1e6)
            print("fwd_%s_generic_nchw" % name, res["fwd_%s_generic_nchw" % name], flush=True)
            _ext.debug_set(False, True, 0)

    if "layout" in what:
        dst = [torch.empty_like(f).contiguous(memory_format=torch.channels_last) for f in feats_nchw]
        lib = _ext.lib()

        def tr():
            for s, d in zip(feats_nchw, dst):
                lib.b200_nchw_to_nhwc(_ext.ptr(s), _ext.ptr(d), B, C, s.shape[2], s.shape[3], _ext.stream_ptr())
        med, best = timeit(tr)
        res["nchw_to_nhwc"] = dict(ms=med, GBs=2 * pyr_bytes / med / 1e6)
        print("nchw_to_nhwc", res["nchw_to_nhwc"], flush=True)
        del dst

    if "bwd" in what:
        for (res_, name) in ((7, "box7"), (14, "mask14")):
            go = torch.randn((R, C, res_, res_), device="cuda", generator=g)
            shp = [tuple(f.shape) for f in feats]
            algo = pyr_bytes + go.numel() * 4 + R * 20
            for nhwc in (True, False):
                med, best = timeit(lambda: _backward(go, rois, shp, nhwc, synth.FPN_SCALES, (res_, res_), 2), iters=5)
                key = "bwd_%s_%s" % (name, "nhwc" if nhwc else "nchw")
                res[key] = dict(ms=med, GBs=algo / med / 1e6, note="includes zero-fill of grads")
                print(key, res[key], flush=True)
            del go

    if "nms" in what:
        for (nper, label) in ((6000, "rpn6000"), (1000, "rpn1000")):
            lens = [min(nper, 3 * h * w) for (h, w) in shapes + [(13, 21)]] * B
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            bs, ss = [], []
            for L in lens:
                b_, s_ = synth.make_nms_boxes(rng, L)
                o = np.argsort(-s_, kind="stable")
                bs.append(b_[o]); ss.append(s_[o])
            boxes = torch.from_numpy(np.concatenate(bs)).cuda()
            scores = torch.from_numpy(np.concatenate(ss)).cuda()
            offs = torch.from_numpy(off).cuda()
            pairs = sum(L * (L - 1) // 2 for L in lens)
            for mk in (-1, 1000):
                med, best = timeit(lambda: nms_batched(boxes, scores, offs, 0.7, mk, max(lens)))
                key = "nms_%s_keep%d" % (label, mk)
                res[key] = dict(ms=med, segs=len(lens), boxes=int(off[-1]), mboxes_s=off[-1] / med / 1e3,
                                gpairs_s=pairs / med / 1e6)
                print(key, res[key], flush=True)
            # unsorted input exercises the bitonic sort
            perm = torch.cat([torch.randperm(L, device="cuda") + int(o) for L, o in zip(lens, off[:-1])])
            b2, s2 = boxes[perm], scores[perm]
            med, best = timeit(lambda: nms_batched(b2, s2, offs, 0.7, -1, max(lens)))
            res["nms_%s_unsorted" % label] = dict(ms=med)
            print("nms_%s_unsorted" % label, res["nms_%s_unsorted" % label], flush=True)

    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/perf_roi.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
