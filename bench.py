#!/usr/bin/env python
"""bench.py -- RoI pseudo-labeling hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1     # the reference's CPU path, host cores

Workload (BASELINE.json configs[1]): 16 synthetic COCO-shaped images per GPU (800x1333 padded
to 800x1344), 256-channel fp32 FPN features P2-P5 (channels_last), RPN candidates on 5 levels
(6000/6000/6000/3150/819 per image), 1000 proposals per image.

One step = one pass of the hot path over the batch:
  1. batched RPN NMS (80 segments, IoU 0.7, keep <= 1000 per level)      b200_nms_batched
  2. per-image top-1000 over levels -> RoI format (rpn/inference.py:173-180)   b200_select_topk
  3. fused 4-level RoIAlign 7x7 on 16000 RoIs                            b200_roi_align_forward
  4. head: mean-pool (fused into 3) -> Linear(256->1024) standing in for fc6/fc7 (out of scope) ->
     emb_pred Linear(1024->768) at the reference's FPN width, both on tcgen05    b200_linear_bf16
  5. class-embedding match, 66 classes, softmax + top-1 fused              b200_embed_match
  6. caption alignment, 1-10 nouns per image (column max + sigmoid)        b200_embed_match + decode
  7. fused RoIAlign 14x14 (mask pooler) on the aligned pseudo-label boxes  b200_roi_align_forward
  8. pack fixed-size pseudo-label records; N > 1: NCCL all-gather of the records

Prints ONE JSON line (see the field notes in DESIGN.md, "Measurement").
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402

METRIC = "RoIs/sec pooled+scored (images/sec pseudo-labeled = value/1000)"
UNIT = "RoIs/s"
B_IMG = 16
R_IMG = 1000
C_FEAT = 256
N_CLASSES = 66
EMB_DIM = 768
HEAD_DIM = 1024   # representation size of the FPN box head (MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM), the input width of emb_pred
RPN_LENS = [6000, 6000, 6000, 3150, 819]  # min(6000, 3*H*W) on P2..P6 (reference PRE_NMS_TOP_N_TEST)
WORKLOAD = ("roi_hot_path microbench: %d img/GPU x %d proposals, 5-level RPN NMS (6000/level, thr 0.7, keep 1000), "
            "4-level RoIAlign 7x7 sr2 on 256ch fp32, 66-class embedding match D=768, caption alignment, "
            "14x14 mask pooler on aligned boxes" % (B_IMG, R_IMG))


# --------------------------------------------------------------------------------------------
# synthetic inputs (seeded)
# --------------------------------------------------------------------------------------------
def make_rpn_candidates(rng, n_img):
    bs, ss = [], []
    for _ in range(n_img):
        for L in RPN_LENS:
            b, s = synth.make_nms_boxes(rng, L)
            o = np.argsort(-s, kind="stable")  # RPN top-k output arrives score-sorted
            bs.append(b[o])
            ss.append(s[o])
    return np.concatenate(bs), np.concatenate(ss)


def make_text(seed, n_img):
    import torch
    g = torch.Generator().manual_seed(seed)
    E = torch.nn.functional.normalize(torch.randn((N_CLASSES, EMB_DIM), generator=g), dim=-1)
    E[0] = 0
    nw = torch.randint(1, 11, (n_img,), generator=g).tolist()
    words = [torch.nn.functional.normalize(torch.randn((w, EMB_DIM), generator=g), dim=-1) for w in nw]
    Wfc = torch.randn((HEAD_DIM, C_FEAT), generator=g) * (1.0 / C_FEAT ** 0.5)   # fc6/fc7 stand-in
    Wemb = torch.randn((EMB_DIM, HEAD_DIM), generator=g) * (3.0 / HEAD_DIM ** 0.5)  # emb_pred (roi_box_predictors.py:63-66)
    return E, words, Wfc, Wemb


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref, compiled from /root/reference in the
# build container) on the host cores, one image per worker process
# --------------------------------------------------------------------------------------------
_REF_TREE = "/root/reference"
_ref_classes = None


def _reference_classes():
    """The reference's own Python glue (Pooler, boxlist_nms, BoxList) imported from /root/reference with the
    compiled reference kernels plugged in as maskrcnn_benchmark._C -- the stock code path of BASELINE.md 5.2.
    Only possible where the reference tree exists (the build container); the GPU box has no /root/reference and
    runs the numpy glue around the same compiled kernels (oracle/_ref)."""
    global _ref_classes
    if _ref_classes is None:
        _ref_classes = False
        import oracle
        if os.path.isdir(os.path.join(_REF_TREE, "maskrcnn_benchmark")) and oracle.ref_lib() is not None:
            try:
                from tests.golden.make_golden import install_reference
                install_reference()
                from maskrcnn_benchmark.modeling.poolers import Pooler
                from maskrcnn_benchmark.structures.bounding_box import BoxList
                from maskrcnn_benchmark.structures.boxlist_ops import boxlist_nms, cat_boxlist
                _ref_classes = dict(Pooler=Pooler, BoxList=BoxList, boxlist_nms=boxlist_nms, cat_boxlist=cat_boxlist)
            except Exception:
                _ref_classes = False
    return _ref_classes


def _ref_inputs(seed):
    rng = np.random.default_rng(seed)
    feats = [rng.standard_normal((1, C_FEAT, h, w), dtype=np.float32) for (h, w) in synth.fpn_shapes()]
    boxes, scores = make_rpn_candidates(rng, 1)
    g = np.random.default_rng(seed + 1)
    E = g.standard_normal((N_CLASSES, EMB_DIM)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    E[0] = 0
    n_words = 1 + (seed % 10)   # 1..10 nouns per caption, as in the GPU arm
    W = g.standard_normal((n_words, EMB_DIM)).astype(np.float32)
    Wfc = (g.standard_normal((HEAD_DIM, C_FEAT)) / C_FEAT ** 0.5).astype(np.float32)
    Wemb = (g.standard_normal((EMB_DIM, HEAD_DIM)) * 3.0 / HEAD_DIM ** 0.5).astype(np.float32)
    return feats, boxes, scores, E, W, Wfc, Wemb


def _ref_image_classes(seed, cls):
    """One image through the reference's Python classes (boxlist_nms, Pooler) on its compiled CPU kernels."""
    import torch
    feats, boxes, scores, E, W, Wfc, Wemb = _ref_inputs(seed)
    ft = [torch.from_numpy(f) for f in feats]
    Et, Wt, Wfct, Wembt = torch.from_numpy(E), torch.from_numpy(W), torch.from_numpy(Wfc), torch.from_numpy(Wemb)
    BoxList, boxlist_nms, cat_boxlist, Pooler = cls["BoxList"], cls["boxlist_nms"], cls["cat_boxlist"], cls["Pooler"]
    size = (synth.IMG_W, synth.IMG_H)
    t0 = time.perf_counter()
    with torch.no_grad():
        per_level, o = [], 0
        for L in RPN_LENS:      # rpn/inference.py:111-122
            bl = BoxList(torch.from_numpy(boxes[o:o + L]), size, mode="xyxy")
            bl.add_field("objectness", torch.from_numpy(scores[o:o + L]))
            per_level.append(boxlist_nms(bl, 0.7, max_proposals=1000, score_field="objectness"))
            o += L
        bl = cat_boxlist(per_level)
        obj = bl.get_field("objectness")
        if len(bl) > R_IMG:     # select_over_all_levels, test mode (rpn/inference.py:173-180)
            _, idx = torch.topk(obj, R_IMG, dim=0, sorted=True)
            bl = bl[idx]
        pooled = Pooler((7, 7), synth.FPN_SCALES, 2)(ft, [bl])
        emb = torch.nn.functional.linear(torch.nn.functional.linear(pooled.mean(dim=(2, 3)), Wfct), Wembt)
        probs = torch.softmax(torch.einsum("pe,ce->pc", emb, Et), -1)     # roi_box_predictors.py:67, inference.py:62
        sc, idx = torch.max(torch.einsum("pd,wd->pw", emb, Wt), dim=0)      # st_generalized_rcnn.py:245-247
        _ = torch.sigmoid(sc)
        Pooler((14, 14), synth.FPN_SCALES, 2)(ft, [bl[idx]])
    dt = time.perf_counter() - t0
    assert probs.shape[1] == N_CLASSES
    return dt


def _ref_image(seed):
    """The same hot path for ONE image through the reference CPU kernels.  Returns seconds."""
    import oracle
    cls = _reference_classes()
    if cls:
        return _ref_image_classes(seed, cls)
    feats, boxes, scores, E, W, Wfc, Wemb = _ref_inputs(seed)
    use_ref = oracle.ref_lib() is not None
    nms = oracle.ref_nms if use_ref else oracle.nms
    ra = oracle.ref_roi_align_forward if use_ref else oracle.roi_align_forward
    t0 = time.perf_counter()
    # RPN NMS per level (rpn/inference.py:111-122), then per-image top-1000 (:173-180)
    kb, ks, o = [], [], 0
    for L in RPN_LENS:
        k = nms(boxes[o:o + L], scores[o:o + L], 0.7)[:1000]
        kb.append(boxes[o:o + L][k])
        ks.append(scores[o:o + L][k])
        o += L
    kb, ks = np.concatenate(kb), np.concatenate(ks)
    top = np.argsort(-ks, kind="stable")[:R_IMG]
    rois = np.concatenate([np.zeros((len(top), 1), np.float32), kb[top]], 1)
    # Pooler.forward (poolers.py:91-121)
    lv = oracle.level_map(rois, 2.0, 5.0)
    pooled = np.zeros((len(rois), C_FEAT, 7, 7), np.float32)
    for l, (f, s) in enumerate(zip(feats, synth.FPN_SCALES)):
        idx = np.nonzero(lv == l)[0]
        if len(idx):
            pooled[idx] = ra(f, rois[idx], s, 7, 7, 2)
    emb = (pooled.mean(axis=(2, 3)) @ Wfc.T) @ Wemb.T
    probs = oracle.softmax_rows(emb @ E.T)                       # roi_box_predictors.py:67, inference.py:62
    sc = emb @ W.T                                               # st_generalized_rcnn.py:245-255
    idx = sc.argmax(0)
    _ = 1 / (1 + np.exp(-sc[idx, np.arange(W.shape[0])]))
    sel = rois[idx]
    lv = oracle.level_map(sel, 2.0, 5.0)
    for l, (f, s) in enumerate(zip(feats, synth.FPN_SCALES)):
        j = np.nonzero(lv == l)[0]
        if len(j):
            ra(f, sel[j], s, 14, 14, 2)
    dt = time.perf_counter() - t0
    assert probs.shape == (R_IMG, N_CLASSES)
    return dt


def bench_config(**kw):
    """The `config` object of the JSON line: the same keys in both arms (the driver compares them)."""
    cfg = {"workload": WORKLOAD, "images_per_gpu": B_IMG, "rois_per_image": R_IMG,
           "feature_layout": "channels_last (NHWC memory, logical [B,C,H,W])",
           "roi_align_math": "fast", "l2": "inputs (1.46 GB features/GPU) exceed the 126 MB L2; no flush needed",
           "launch": "n/a", "sample": "n/a"}
    cfg.update(kw)
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    import oracle
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    kind = "reference" if oracle.ref_lib() is not None else "port"
    glue = ("reference Python classes (Pooler, boxlist_nms) from /root/reference on its compiled CPU kernels"
            if _reference_classes() else "numpy glue around the compiled reference CPU kernels (no /root/reference on this box)")
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(workers) as pool:
        for it in range(args.warmup + args.steps):
            # workers run concurrently; a step lasts as long as its slowest image (input
            # generation inside the worker is outside its timed region)
            dts = pool.map(_ref_image, [1000 + 97 * it + w for w in range(workers)])
            if it >= args.warmup:
                times.append(max(dts))
    ms = 1e3 * float(np.mean(times)) if times else float("nan")
    value = workers * R_IMG / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(images_per_gpu=workers, feature_layout="NCHW (the reference's layout)",
                               roi_align_math="exact (the reference's own ROIAlign_cpu)", l2="n/a (host caches)",
                               launch=glue, sample="%d images per step, one per worker process" % workers),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": "%d images/step (1 per worker), csrc kernels single-threaded as shipped; %s" % (workers, glue)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# clocks sampler (NVML, background thread)
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def window_begin(self):
        """Only what is sampled from here on is reported.  The thread is started BEFORE the warm-up: its first NVML
        calls cost milliseconds when eight ranks make them at once, and inside a 17 ms timed region that showed up as
        one 3 ms step (bench line at N = 8: 1.09 instead of 0.98 ms/step)."""
        self.begin = len(self.samples)
        self.reasons = set()

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        mine = self.samples[getattr(self, "begin", 0):]
        return {"sm_mhz": float(np.median(mine)) if mine else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(mine)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def f_touched_bytes(rois, levels, shapes, scales, channels):
    """Exact union, per level, of the feature pixels inside the tap rectangle of the RoIs assigned
    to the level, in bytes (fp32, all images)."""
    import torch
    total = 0
    for l, ((h, w), s) in enumerate(zip(shapes, scales)):
        m = levels == l
        if not bool(m.any()):
            continue
        r = rois[m]
        b = r[:, 0].long()
        x0 = (r[:, 1] * s).floor().clamp(0, w - 1).long()
        y0 = (r[:, 2] * s).floor().clamp(0, h - 1).long()
        x1 = ((r[:, 3] * s).floor() + 1).clamp(0, w - 1).long()
        y1 = ((r[:, 4] * s).floor() + 1).clamp(0, h - 1).long()
        x1, y1 = torch.maximum(x1, x0), torch.maximum(y1, y0)
        nb = int(b.max().item()) + 1
        d = torch.zeros((nb, h + 1, w + 1), dtype=torch.int32, device=rois.device)
        one = torch.ones_like(b, dtype=torch.int32)
        d.index_put_((b, y0, x0), one, accumulate=True)
        d.index_put_((b, y1 + 1, x0), -one, accumulate=True)
        d.index_put_((b, y0, x1 + 1), -one, accumulate=True)
        d.index_put_((b, y1 + 1, x1 + 1), one, accumulate=True)
        cov = d.cumsum(1).cumsum(2)[:, :h, :w] > 0
        total += int(cov.sum().item()) * channels * 4
    return total


_ORIG_AFFINITY = None


def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs NVML names as local to GPU `index` BEFORE the pinned host buffers are
    allocated (first touch decides the NUMA node of pinned pages): with 8 ranks each copying 1.47 GB per step the
    e2e number is a host-memory / PCIe-uplink number, and remote pages make it worse.  Returns a short note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        if cpus:
            global _ORIG_AFFINITY
            _ORIG_AFFINITY = os.sched_getaffinity(0)
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d (NVML affinity of GPU %d)" % (cpus[0], cpus[-1], index)
    except Exception as e:
        return "not bound (%s)" % type(e).__name__
    return "not bound"


# the library kernels one step launches (own kernels only; torch glue such as index_select / cat is not counted)
STEP_KERNELS = ["nms_fused_kernel", "select_topk_kernel", "roi_order_kernel", "roi_align_fwd_rows",
                "tc_gemm_kernel<linear> (fc stand-in)", "tc_gemm_kernel<linear> (emb_pred)", "tc_gemm_kernel<softmax>",
                "embed_match_kernel (caption alignment)", "colmax_decode_kernel", "roi_align_fwd_sep (mask pooler)"]


def roofline_passes(torch, dev, feats, scales, shapes, state, cand_boxes, cand_scores, seg_off, args):
    """One timed pass per hot kernel at the BASELINE config sizes -> list of {kernel, bound, achieved, peak, unit,
    frac, ms, ...}.  Peaks: MEASURED_PEAKS.json (hbm_gbs; bf16_tflops_sustained).  Bytes / flops: SURVEY 8(d)."""
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import embed_match_softmax, linear_bf16, nms_batched
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _backward, _forward
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm, tf = float(peaks.get("hbm_gbs", 6650.0)), float(peaks.get("bf16_tflops_sustained", 1400.0))

    def timeit(fn, iters=7, reps=1):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) / reps)
        return float(np.median(ts))

    def graphed(fn, reps=8):
        """device time per call with launch gaps removed (kernels of tens of microseconds)"""
        fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        return timeit(g.replay, iters=7) / reps

    out = []
    rois, levels = state["rois"], state["levels"]
    n = rois.shape[0]
    ft = f_touched_bytes(rois, levels, shapes, scales, C_FEAT)
    f_all = sum(int(f.numel()) * 4 for f in feats)

    def hbm_entry(kernel, nbytes, ms, **kw):
        e = {"kernel": kernel, "bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm, "unit": "GB/s",
             "frac": nbytes / ms / 1e6 / hbm, "ms": ms, "algorithmic_bytes": int(nbytes)}
        e.update(kw)
        return e

    def tensor_entry(kernel, flops, ms, **kw):
        e = {"kernel": kernel, "bound": "tensor", "achieved": flops / ms / 1e9, "peak": tf, "unit": "TFLOP/s",
             "frac": flops / ms / 1e9 / tf, "ms": ms, "algorithmic_flops": float(flops)}
        e.update(kw)
        return e

    # RoIAlign forward, box 7x7 and mask 14x14, on the step's 16000 RoIs (bytes: touched features + rois + output)
    for res, name in ((7, "roi_align fwd 7x7 (box pooler, %s)" % args.math), (14, "roi_align fwd 14x14 (mask pooler at R=16000, %s)" % args.math)):
        ms = timeit(lambda: _forward(feats, scales, rois, (res, res), 2, math=args.math))
        out.append(hbm_entry(name, ft + n * (C_FEAT * res * res * 4 + 20), ms, rois=int(n)))
    # RoIAlign backward (bytes: grad_out read + rois + every gradient element written once, zero fill included)
    fshapes = [tuple(f.shape) for f in feats]
    for res in (7, 14):
        g = torch.randn((n, C_FEAT, res, res), device=dev)
        ms = timeit(lambda: _backward(g, rois, fshapes, True, scales, (res, res), 2), iters=5)
        out.append(hbm_entry("roi_align bwd %dx%d (incl. zero fill of the gradient pyramid)" % (res, res),
                             n * (C_FEAT * res * res * 4 + 20) + f_all, ms, rois=int(n)))
        del g
    # BASELINE config #1, the reference's shipped pooler: C4 map [1, 1024, 50, 84] NCHW, 14x14 bins, adaptive
    # sampling (sampling_ratio 0), scale 1/16, 1000 RoIs (config/defaults.py:301-305) -- exact arithmetic, both layouts
    try:
        gen1 = torch.Generator(device=dev).manual_seed(1235)
        c4 = torch.randn((1, 1024, 50, 84), device=dev, generator=gen1)
        r1 = torch.from_numpy(synth.make_rois(np.random.default_rng(1235), 1000, 1)).to(dev)
        flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)    # the 17 MB map would sit in the L2
        from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import nhwc_cache
        for layout, x in (("NCHW (the reference's layout; timed with the NHWC re-layout the Pooler stages)", c4),
                          ("channels_last", c4.contiguous(memory_format=torch.channels_last))):
            ts = []
            for _ in range(5):
                flush.zero_()
                nhwc_cache.entries.clear()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                # (what modeling.Pooler does with an NCHW-contiguous map: b200_nchw_to_nhwc, then roi_align_fwd_tile)
                _forward([nhwc_cache.get(x) if x.is_contiguous() else x], (1.0 / 16,), r1, (14, 14), 0, math="exact")
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ms = float(np.median(ts[1:]))
            out.append(hbm_entry("roi_align fwd config #1: C4 1024 ch, 14x14, sampling_ratio 0, 1000 RoIs, %s, exact" % layout,
                                 c4.numel() * 4 + 1000 * (1024 * 196 * 4 + 20), ms, rois=1000, l2="flushed between iterations"))
        del c4, flush
    except Exception as e:
        out.append({"kernel": "roi_align fwd config #1", "error": "%s: %s" % (type(e).__name__, e)})
    # RPN NMS: 80 segments, keep <= 1000 (pairwise-IoU tests are data dependent: boxes/s and the bitmask convention)
    lens = np.array(RPN_LENS * B_IMG, dtype=np.float64)
    ms = timeit(lambda: nms_batched(cand_boxes, cand_scores, seg_off, 0.7, 1000, max(RPN_LENS)))
    bm_bytes = float(np.sum(20 * lens + 8 * lens * np.ceil(lens / 64) + 8 * 1000))
    out.append(hbm_entry("nms_fused (RPN, 80 segments, keep 1000)", bm_bytes, ms, convention="materialised 64x64 bitmask (SURVEY 8d)",
                         boxes_per_s=float(lens.sum() / ms * 1e3), segments=int(len(lens))))
    ms = timeit(lambda: nms_batched(cand_boxes, cand_scores, seg_off, 0.7, -1, max(RPN_LENS)), iters=3)
    out.append(hbm_entry("nms (RPN, 80 segments, keep all: bitmask path)", bm_bytes, ms, convention="materialised 64x64 bitmask (SURVEY 8d)",
                         pair_tests_per_s=float(np.sum(lens * (lens - 1) / 2) / ms * 1e3)))
    # box-head per-class NMS: 16 images x 65 classes, candidates = prob > 0.05 of the step's scores (thr 0.5)
    try:
        from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
        # class probabilities with a few confident classes per RoI (the step's own scores come from random
        # embeddings and are nearly flat): a dominant class shared by neighbouring boxes plus two weak ones
        gp = torch.Generator(device=dev).manual_seed(7)
        probs = torch.full((n, N_CLASSES), 1e-4, device=dev)
        ctr = ((rois[:, 1] + rois[:, 3]) * (0.5 / 160.0)).long() + 16 * ((rois[:, 2] + rois[:, 4]) * (0.5 / 160.0)).long()
        main = 1 + (ctr * 7 + rois[:, 0].long() * 13) % (N_CLASSES - 1)
        probs[torch.arange(n, device=dev), main] = 0.3 + 0.6 * torch.rand((n,), device=dev, generator=gp)
        for _ in range(2):
            other = 1 + torch.randint(0, N_CLASSES - 1, (n,), device=dev, generator=gp)
            probs[torch.arange(n, device=dev), other] = 0.05 + 0.1 * torch.rand((n,), device=dev, generator=gp)
        reg = torch.randn((n, 8), device=dev, generator=gp) * 0.1
        offs = torch.arange(0, n + 1, R_IMG, dtype=torch.int32, device=dev)
        im = torch.tensor([[float(synth.IMG_W), float(synth.IMG_H)]] * B_IMG, device=dev)
        cap = n * 8
        nseg = B_IMG * (N_CLASSES - 1)
        so = torch.empty((nseg + 1,), dtype=torch.int32, device=dev)
        cb = torch.empty((cap, 4), device=dev)
        cs = torch.empty((cap,), device=dev)
        cr = torch.empty((cap,), dtype=torch.int32, device=dev)
        st = torch.empty((2,), dtype=torch.int32, device=dev)
        boxes_xyxy = rois[:, 1:].contiguous()
        sl = torch.empty((nseg,), dtype=torch.int32, device=dev)
        rc = _ext.lib().b200_box_candidates(_ext.ptr(probs), _ext.ptr(reg), _ext.ptr(boxes_xyxy), _ext.ptr(offs), _ext.ptr(im),
                                            B_IMG, n, N_CLASSES, 8, 1, 10.0, 10.0, 5.0, 5.0, 0.05, cap, _ext.ptr(sl), _ext.ptr(so),
                                            _ext.ptr(cb), _ext.ptr(cs), _ext.ptr(cr), _ext.ptr(st), _ext.stream_ptr(dev))
        _ext.check(rc, "b200_box_candidates")
        total = int(st[0].item())
        seg_len = (so[1:] - so[:-1]).cpu().numpy().astype(np.float64)
        ms = graphed(lambda: nms_batched(cb[:total], cs[:total], so, 0.5, -1, int(seg_len.max()) if total else 1))
        out.append(hbm_entry("nms_fused (box head, %d segments = 16 images x 65 classes)" % nseg,
                             float(np.sum(20 * seg_len + 8 * seg_len * np.ceil(seg_len / 64) + 8 * seg_len)), ms,
                             convention="materialised 64x64 bitmask (SURVEY 8d)", boxes_per_s=float(total / ms * 1e3), candidates=total))
    except Exception as e:
        out.append({"kernel": "nms_fused (box head)", "error": "%s: %s" % (type(e).__name__, e)})
    # scoring: config #3 shape (HBM-bound on the embeddings) and config #5 (tensor-bound)
    gen = torch.Generator(device=dev).manual_seed(99)
    for name, (r, c, d), bound in (("embed_match cfg#3 (64000 x 66, D=768)", (64000, 66, 768), "hbm"),
                                   ("embed_match cfg#5 (262144 x 501, D=512)", (262144, 501, 512), "tensor")):
        A = (torch.randn((r, d), device=dev, generator=gen) * (3.0 / d ** 0.5)).to(torch.bfloat16)
        Em = torch.nn.functional.normalize(torch.randn((c, d), device=dev, generator=gen), dim=-1).to(torch.bfloat16)
        for probs_out in (False, True):
            ms = graphed(lambda: embed_match_softmax(A, Em, 0.05, want_probs=probs_out), reps=4)
            nbytes = r * d * 2 + c * d * 2 + r * 8 + (r * c * 4 if probs_out else 0)
            tag = name + (", probabilities written" if probs_out else ", top-1 only")
            if bound == "hbm":
                out.append(hbm_entry(tag, nbytes, ms, tflops=2.0 * r * c * d / ms / 1e9))
            else:
                out.append(tensor_entry(tag, 2.0 * r * c * d, ms, gbs=nbytes / ms / 1e6, hbm_frac=nbytes / ms / 1e6 / hbm))
        del A, Em
    # emb_pred projection at the reference's widths
    for r, k in ((64000, 1024), (64000, 2048)):
        x = torch.randn((r, k), device=dev, generator=gen).to(torch.bfloat16)
        w = (torch.randn((EMB_DIM, k), device=dev, generator=gen) * 0.02).to(torch.bfloat16)
        b = torch.zeros((EMB_DIM,), device=dev)
        ms = graphed(lambda: linear_bf16(x, w, b, want_f32=False, want_bf16=True), reps=4)
        out.append(tensor_entry("emb_pred linear (%d x %d -> 768, b200_linear_bf16)" % (r, k), 2.0 * r * k * EMB_DIM, ms))
        del x, w
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import (caption_align, embed_match_softmax, linear_bf16, nms_batched,
                                                                select_topk)
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _backward as roi_align_backward
    from cvpr22_cross_modal_pseudo_labeling_b200.layers.roi_align import _forward as roi_align_forward
    from cvpr22_cross_modal_pseudo_labeling_b200.parallel import all_gather_records

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (the B200 path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _ext.lib()
    _ext.debug_set(False, True, 0)
    numa = bind_to_gpu_numa_node(local)

    # every rank draws its own synthetic batch unless --same-seed (then the max over ranks carries no
    # data-dependent spread: what is left of the N = 1 -> N step is the machine, not the boxes)
    seed = 1236 + (0 if args.same_seed else 1000 * rank)
    rng = np.random.default_rng(seed)
    shapes = synth.fpn_shapes()
    scales = synth.FPN_SCALES
    gen = torch.Generator(device=dev).manual_seed(seed)
    # ---- host (pinned) copies of every input the step consumes ----
    feats_h = [torch.empty((B_IMG, h, w, C_FEAT), dtype=torch.float32).pin_memory() for (h, w) in shapes]
    feats = []
    for fh, (h, w) in zip(feats_h, shapes):
        f = torch.randn((B_IMG, h, w, C_FEAT), device=dev, generator=gen)
        fh.copy_(f)
        feats.append(f.permute(0, 3, 1, 2))  # logical [B,C,H,W], channels_last memory
    cb, cs = make_rpn_candidates(rng, B_IMG)
    K = sum(RPN_LENS)
    cand_boxes_h = torch.from_numpy(cb).pin_memory()
    cand_scores_h = torch.from_numpy(cs).pin_memory()
    seg_off = torch.from_numpy(np.concatenate([[0], np.cumsum(RPN_LENS * B_IMG)]).astype(np.int32)).to(dev)
    E, words, Wfc, Wemb = make_text(seed, B_IMG)
    E_h = E.to(torch.bfloat16).pin_memory()
    words_h = [w.to(torch.bfloat16).pin_memory() for w in words]
    n_words = [int(w.shape[0]) for w in words]
    Wfc_bf = Wfc.to(dev).to(torch.bfloat16)
    Wemb_bf = Wemb.to(dev).to(torch.bfloat16)
    bemb = torch.zeros((EMB_DIM,), device=dev)
    img_of_word = torch.repeat_interleave(torch.arange(B_IMG, device=dev), torch.tensor(n_words, device=dev))
    w_max = 10
    h2d_bytes = sum(f.numel() * 4 for f in feats_h) + cand_boxes_h.numel() * 4 + cand_scores_h.numel() * 4 + \
        E_h.numel() * 2 + sum(w.numel() * 2 for w in words_h)

    state = {}
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for k in ("nms", "pool7", "head", "match", "pool14")}
    acc = {k: [] for k in ev}

    # static index helpers of the step (built once, so the step issues no host->device copies
    # and can be replayed as a CUDA graph)
    word_slot = torch.cat([torch.arange(w, device=dev) for w in n_words])
    cnt_words = torch.tensor(n_words, dtype=torch.int32, device=dev)
    word_base = img_of_word * R_IMG
    rec_img = (img_of_word.float() + B_IMG * rank)[:, None]
    rec_slot = word_slot.float()[:, None]
    rows_per_image = [R_IMG] * B_IMG

    def step(cand_boxes, cand_scores, feats_d, E_d, words_d, timed=False):
        def mark(k, i):
            if timed:
                ev[k][i].record()
        # 1. batched RPN NMS
        mark("nms", 0)
        keep_idx, keep_cnt = nms_batched(cand_boxes, cand_scores, seg_off, 0.7, 1000, max(RPN_LENS))
        mark("nms", 1)
        # 2. per-image top-1000 over the levels, straight into RoI format (no host sync)
        rois, _, _ = select_topk(cand_boxes, cand_scores, seg_off, keep_idx, keep_cnt, B_IMG, R_IMG, 5 * 1000)
        # 3. box pooler
        mark("pool7", 0)
        # (+ the per-channel mean of every pooled block, FastRCNNPredictor's AvgPool2d, taken from
        # the pooler's shared-memory tile instead of a second pass over the 803 MB pooled tensor)
        pooled_mean = torch.empty((B_IMG * R_IMG, C_FEAT), dtype=torch.float32, device=dev)
        pooled, levels = roi_align_forward(feats_d, scales, rois, (7, 7), 2, want_levels=True, math=args.math,
                                           mean_out=pooled_mean)
        mark("pool7", 1)
        # 4. head: Linear(256 -> 1024) standing in for fc6/fc7, then emb_pred Linear(1024 -> 768) at the
        # reference's FPN width (roi_box_predictors.py:63-66) -- both on the tcgen05 GEMM, bf16 embeddings out
        mark("head", 0)
        hid = linear_bf16(pooled_mean, Wfc_bf, None, want_f32=False, want_bf16=True)[1]
        emb = linear_bf16(hid, Wemb_bf, bemb, want_f32=False, want_bf16=True)[1]
        mark("head", 1)
        # 5-6. scoring + caption alignment
        mark("match", 0)
        cls = embed_match_softmax(emb, E_d, 0.05, want_probs=True)
        aligned = caption_align(emb, rows_per_image, words_d)
        mark("match", 1)
        idx = torch.cat([a[0] for a in aligned])
        sig = torch.cat([a[2] for a in aligned])
        sel = rois[word_base + idx]
        # 7. mask pooler on the aligned boxes
        mark("pool14", 0)
        mask_feat, _ = roi_align_forward(feats_d, scales, sel, (14, 14), 2, math=args.math)
        mark("pool14", 1)
        # 8. records (img, word slot, box, score, region)
        rec = torch.zeros((B_IMG, w_max, 8), dtype=torch.float32, device=dev)
        rec[img_of_word, word_slot] = torch.cat([rec_img, rec_slot, sel[:, 1:], sig[:, None], idx[:, None].float()], dim=1)
        state.update(pooled_mean=pooled_mean, rois=rois, levels=levels, rec=rec, probs=cls["probs"], mask_feat=mask_feat, sel=sel)
        return rec

    def gather(rec):
        # N > 1: the one collective of the path -- NCCL all-gather of the fixed-size records
        if world > 1:
            return all_gather_records(rec, cnt_words, sizes=[B_IMG] * world)
        return rec, cnt_words

    class GraphedStep(object):
        """The step captured once as a CUDA graph and replayed (static shapes, no host syncs)."""

        def __init__(self, fn):
            self.fn = fn
            self.graph = None
            self.out = None
            if args.no_graph:
                return
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = fn()
            self.graph = g

        def __call__(self):
            if self.graph is None:
                return self.fn()
            self.graph.replay()
            return self.out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident inputs ----
    cand_boxes = cand_boxes_h.to(dev)
    cand_scores = cand_scores_h.to(dev)
    E_d = E_h.to(dev)
    words_d = [w.to(dev) for w in words_h]

    run_step = GraphedStep(lambda: step(cand_boxes, cand_scores, feats, E_d, words_d))

    # N > 1: the all-gather of step i runs on its own stream and overlaps the compute of step i+1.
    # The step's (static) record buffer is first copied into one of two staging buffers, so the
    # next graph replay may overwrite it; a staging buffer is reused only after its gather is done.
    gstream = torch.cuda.Stream() if world > 1 else None
    staging, gather_done = [None, None], [None, None]

    def gather_overlapped(rec, i):
        if world == 1:
            return rec, cnt_words
        main = torch.cuda.current_stream()
        slot = i & 1
        if staging[slot] is None:
            staging[slot] = torch.empty_like(rec)
        if gather_done[slot] is not None:
            main.wait_event(gather_done[slot])
        staging[slot].copy_(rec)
        ready = torch.cuda.Event()
        ready.record(main)
        gstream.wait_event(ready)
        with torch.cuda.stream(gstream):
            out = all_gather_records(staging[slot], cnt_words, sizes=[B_IMG] * world)
            done = torch.cuda.Event()
            done.record(gstream)
        gather_done[slot] = done
        return out

    def drain_gathers():
        for e in gather_done:
            if e is not None:
                torch.cuda.current_stream().wait_event(e)

    sampler = ClockSampler(local)
    if not os.environ.get("B200_BENCH_NO_SAMPLER"):
        sampler.start()
    for i in range(args.warmup):
        gather_overlapped(run_step(), i)
    drain_gathers()
    sync_all()
    sampler.window_begin()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]   # (diagnostic: per-step device times)
    t0.record()
    for i in range(args.steps):
        gather_overlapped(run_step(), i)
        marks[i].record()
    drain_gathers()   # the timed region ends when the last records have arrived everywhere
    t1.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = t0.elapsed_time(t1)
    step_marks = [t0.elapsed_time(m) for m in marks]
    step_ms = [b - a for a, b in zip([0.0] + step_marks[:-1], step_marks)]

    # sustained: the same step replayed back to back for >= --sustain seconds (the timed K steps above last
    # ~20 ms; this shows whether the number holds once clocks and power settle)
    sustained = None
    if args.sustain > 0:
        # the step count must be the same on every rank (each replay issues one all-gather): agree on the slowest
        ms_any = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_any, op=dist.ReduceOp.MAX)
        n_sus = max(args.steps, int(args.sustain * 1e3 / max(float(ms_any[0]) / args.steps, 1e-3)))
        sync_all()
        sampler2 = ClockSampler(local)
        sampler2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_sus):
            gather_overlapped(run_step(), i)
        drain_gathers()
        s1.record()
        sync_all()
        c2 = sampler2.stop()
        sustained = {"steps": n_sus, "ms_per_step": s0.elapsed_time(s1) / n_sus, "seconds": s0.elapsed_time(s1) / 1e3,
                     "sm_mhz": c2["sm_mhz"], "reasons": c2["reasons"], "samples": c2["samples"]}

    # per-kernel durations (separate pass so the events do not perturb the headline loop)
    for _ in range(max(3, min(args.steps, 10))):
        step(cand_boxes, cand_scores, feats, E_d, words_d, timed=True)
        torch.cuda.synchronize()
        for k in ev:
            acc[k].append(ev[k][0].elapsed_time(ev[k][1]))
    kms = {k: float(np.median(v)) for k, v in acc.items()}
    # the box pooler in the other arithmetic mode, same RoIs, for the record
    other = "exact" if args.math == "fast" else "fast"
    ts = []
    for _ in range(5):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        roi_align_forward(feats, scales, state["rois"], (7, 7), 2, math=other, mean_out=state["pooled_mean"])
        a1.record()
        a1.synchronize()
        ts.append(a0.elapsed_time(a1))
    kms["pool7_%s_math" % other] = float(np.median(ts[1:]))

    # cost of the one-off NCHW -> NHWC re-layout a caller with NCHW-contiguous maps pays per
    # batch (shared by both poolers and the backward); reported, not part of `value`
    relayout_ms = None
    try:
        lib = _ext.lib()
        src = [torch.empty((B_IMG, C_FEAT, h, w), device=dev) for (h, w) in shapes[:2]]
        dst = [torch.empty((B_IMG, C_FEAT, h, w), device=dev, memory_format=torch.channels_last) for (h, w) in shapes[:2]]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(4):
            if it == 1:
                e0.record()
            for a_, d_ in zip(src, dst):
                lib.b200_nchw_to_nhwc(_ext.ptr(a_), _ext.ptr(d_), B_IMG, C_FEAT, a_.shape[2], a_.shape[3],
                                      _ext.stream_ptr(dev))
        e1.record()
        torch.cuda.synchronize()
        frac = sum(h * w for (h, w) in shapes[:2]) / float(sum(h * w for (h, w) in shapes))
        relayout_ms = e0.elapsed_time(e1) / 3 / frac   # P2+P3 measured, scaled to the whole pyramid
        del src, dst
    except Exception:
        pass

    # ---- roofline of every hot kernel at the BASELINE config sizes (one extra timed pass each) ----
    roofline_all = []
    if rank == 0 and not args.no_roofline_all:
        try:
            roofline_all = roofline_passes(torch, dev, feats, scales, shapes, state, cand_boxes, cand_scores, seg_off, args)
        except Exception as e:   # a failed extra pass must not lose the headline line
            roofline_all = [{"kernel": "roofline_all", "error": "%s: %s" % (type(e).__name__, e)}]

    # ---- end to end: host buffers in, host records out, copies inside the timed region ----
    feats_e = [torch.empty(f.shape, dtype=torch.float32, device=dev) for f in feats_h]
    cb_e, cs_e = torch.empty_like(cand_boxes), torch.empty_like(cand_scores)
    E_e = torch.empty_like(E_d)
    words_e = [torch.empty_like(w) for w in words_d]
    rec_h = torch.empty((B_IMG * world, w_max, 8), dtype=torch.float32).pin_memory()
    cnt_h = torch.empty((B_IMG * world,), dtype=torch.int32).pin_memory()

    feats_e_view = [f.permute(0, 3, 1, 2) for f in feats_e]

    def e2e_device_part():
        for d, s_ in zip(feats_e, feats_h):
            d.copy_(s_, non_blocking=True)
        cb_e.copy_(cand_boxes_h, non_blocking=True)
        cs_e.copy_(cand_scores_h, non_blocking=True)
        E_e.copy_(E_h, non_blocking=True)
        for d, s_ in zip(words_e, words_h):
            d.copy_(s_, non_blocking=True)
        return step(cb_e, cs_e, feats_e_view, E_e, words_e)

    run_e2e = GraphedStep(e2e_device_part)

    def step_e2e():
        rec, cnt = gather(run_e2e())
        rec_h.copy_(rec, non_blocking=True)
        cnt_h.copy_(cnt, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller holds the records on the host

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_e2e()
    sync_all()
    t2, t3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2.record()
    for _ in range(e2e_steps):
        step_e2e()
    t3.record()
    sync_all()
    ms_e2e_total = t2.elapsed_time(t3)

    # ---- max over ranks ----
    tms = torch.tensor([ms_total, ms_e2e_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms[0]) / args.steps
    ms_e2e = float(tms[1]) / e2e_steps
    rois_per_step = world * B_IMG * R_IMG

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ft = f_touched_bytes(state["rois"], state["levels"], shapes, scales, C_FEAT)
        out_bytes = B_IMG * R_IMG * C_FEAT * 49 * 4
        algo = ft + out_bytes + B_IMG * R_IMG * 20
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roi_align_fwd_ncu.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        achieved = algo / (kms["pool7"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": rois_per_step / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(
                roi_align_math=("fast: separable FMA evaluation (row-streaming kernel), <= 1e-5 rel (+ 1e-6 abs on unit-variance "
                                "features) of ROIAlign_cpu (b200_roi_align_forward_ws)" if args.math == "fast"
                                else "exact: bit-identical to ROIAlign_cpu (b200_roi_align_forward)"),
                launch=("eager" if args.no_graph else "CUDA graph replay of the step") +
                       ("; NCCL all-gather of step i on a side stream, overlapping step i+1" if world > 1 else ""),
                sample="%d images per GPU per step%s" % (B_IMG, "; same seed on every rank" if args.same_seed else "")),
            "images_per_sec": world * B_IMG / (ms_step * 1e-3),
            # the reference's backbones emit NCHW: the same step with the one-off NCHW -> NHWC re-layout of the
            # pyramid added (b200_nchw_to_nhwc, shared by both poolers and the backward)
            "nchw_value": (rois_per_step / ((ms_step + relayout_ms) * 1e-3)) if relayout_ms else None,
            "nchw_input_relayout_ms_per_step": relayout_ms,
            "exact_math_value": (rois_per_step / ((ms_step - kms["pool7"] + kms.get("pool7_exact_math", kms["pool7"])) * 1e-3)
                                 if args.math == "fast" else None),
            "sustained": sustained,
            "step_ms_rank0": {"first3": [round(v, 4) for v in step_ms[:3]], "median": float(np.median(step_ms)),
                              "max": float(np.max(step_ms)), "drain_ms": ms_total - step_marks[-1]},
            "clocks": clocks,
            "e2e": {"value": rois_per_step / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "host_numa": numa,
                    "bound": "PCIe: %.2f GB of fp32 features cross the host link per step and GPU (the device part is %.2f ms)" % (h2d_bytes / 1e9, ms_step),
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(rec_h.numel() * 4 + cnt_h.numel() * 4) // world},
            "gpu_launches": args.steps * len(STEP_KERNELS),
            "step_kernels": STEP_KERNELS,
            "kernel_ms": kms,
            "roofline": {"kernel": "%s (box pooler 7x7)" % ("roi_align_fwd_rows" if args.math == "fast" else "roi_align_fwd_march"), "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes": int(algo), "f_touched_bytes": int(ft), "traffic": traffic},
            "roofline_all": roofline_all,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            if _ORIG_AFFINITY is not None:   # the reference arm gets every host core back
                os.sched_setaffinity(0, _ORIG_AFFINITY)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2",
                                      "--warmup", "1"], capture_output=True, text=True, timeout=600)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "failed: %s" % e}
        else:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "only measured at N=1 (run --impl reference)"}
        print(json.dumps(line), flush=True)
    return 0


def run_train(args):
    """BASELINE config #4: one student training step of the RoI path in bf16 -- box pooler forward + backward on
    512 sampled RoIs per image (ROI_HEADS.BATCH_SIZE_PER_IMAGE, config/defaults.py:282) and the embedding head
    (fc stand-in, emb_pred, class scoring) forward + backward on the tensor cores -- with the head gradients
    all-reduced over NCCL as DDP does (tools/train_net.py:65-71).  Prints one JSON line (not the headline metric)."""
    import torch
    import torch.distributed as dist
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import TensorCoreLinear, embed_logits, roi_align_multilevel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _ext.lib()
    r_img = 512
    seed = 1238 + 1000 * rank
    rng = np.random.default_rng(seed)
    gen = torch.Generator(device=dev).manual_seed(seed)
    shapes, scales = synth.fpn_shapes(), synth.FPN_SCALES
    feats = [torch.randn((B_IMG, h, w, C_FEAT), device=dev, generator=gen).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
             for (h, w) in shapes]
    rois = torch.from_numpy(synth.make_rois(rng, r_img, B_IMG)).to(dev)
    n = rois.shape[0]
    labels = torch.randint(0, N_CLASSES, (n,), device=dev, generator=gen)
    E = torch.nn.functional.normalize(torch.randn((N_CLASSES, EMB_DIM), device=dev, generator=gen), dim=-1)
    E[0] = 0
    E = E.to(torch.bfloat16)
    torch.manual_seed(1238)     # same initial head on every rank
    fc = TensorCoreLinear(C_FEAT, HEAD_DIM).to(dev)
    emb_pred = TensorCoreLinear(HEAD_DIM, EMB_DIM).to(dev)
    params = list(fc.parameters()) + list(emb_pred.parameters())
    flat = torch.zeros((sum(p.numel() for p in params),), device=dev)

    def step():
        for f in feats:
            f.grad = None
        for p_ in params:
            p_.grad = None
        pooled = roi_align_multilevel(feats, rois, (7, 7), scales, 2, math="fast")          # bf16 in, bf16 out
        x = pooled.mean(dim=(2, 3))
        hid = torch.relu(fc(x))
        emb = emb_pred(hid)
        logits = embed_logits(emb, E)
        loss = torch.nn.functional.cross_entropy(logits.float(), labels)
        loss.backward()
        if world > 1:      # DDP-style: one all-reduce of the flattened head gradients
            torch.cat([p_.grad.reshape(-1) for p_ in params], out=flat)
            dist.all_reduce(flat)
        return loss

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sampler.window_begin()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step()
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    tms = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms[0]) / args.steps
    # per-stage device times (one extra pass)
    ev = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        b.synchronize()
        ev[name] = a.elapsed_time(b)
        return out
    for f in feats:
        f.grad = None
    g = None
    for _ in range(3):   # (the last pass is the one reported; gradients dropped first, so that the caching allocator
        for f in feats:  # re-uses the blocks of the pass before instead of calling cudaMalloc inside the timed region)
            f.grad = None
        pooled = timed("pool7_fwd_bf16", lambda: roi_align_multilevel(feats, rois, (7, 7), scales, 2, math="fast"))
        if g is None:
            g = torch.randn(pooled.shape, device=dev, generator=gen).to(torch.bfloat16)
            torch.cuda.synchronize()
        timed("pool7_bwd_incl_casts_and_zero_fill", lambda: pooled.backward(g))
    if rank == 0:
        line = {"metric": "RoIs/sec trained (box pooler fwd+bwd + embedding head fwd+bwd), BASELINE config #4",
                "value": world * n / (ms * 1e-3), "unit": "RoIs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "loss": float(loss.detach()),
                "config": bench_config(workload="student training step of the RoI path: %d img/GPU x %d sampled RoIs, bf16 NHWC features, "
                                                "4-level RoIAlign 7x7 fwd+bwd, fc(256->1024) + emb_pred(1024->768) + 66-class scoring fwd+bwd on "
                                                "tcgen05, cross-entropy; head gradients all-reduced over NCCL" % (B_IMG, r_img),
                                       rois_per_image=r_img, feature_layout="channels_last bf16", launch="eager autograd",
                                       sample="%d images per GPU per step" % B_IMG),
                "clocks": clocks, "stage_ms": ev, "gpu_launches": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="fast", choices=["fast", "exact"],
                    help="RoIAlign forward arithmetic: fast = separable FMA evaluation (<= 1e-5 rel, the tolerance "
                         "BASELINE.json states); exact = the reference's operation order, bit-identical")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="2: the headline inference step (BASELINE configs #2/#3); 4: the bf16 student training step (config #4)")
    ap.add_argument("--same-seed", action="store_true", help="every rank draws the same synthetic batch")
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back replay for the `sustained` record (0 = skip)")
    ap.add_argument("--no-roofline-all", action="store_true", help="skip the extra per-kernel roofline passes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        return run_train(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
