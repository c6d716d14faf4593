"""CPU suite: the C-ABI library exports what include/b200det.h declares, the host-side
mirrors behave like the reference containers, the product path refuses CPU tensors and never
touches the oracle, and the N>1 sharding logic works under gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cvpr22_cross_modal_pseudo_labeling_b200")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200det.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    assert os.path.exists(_ext.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_ext.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_ext.SIGNATURES), "ctypes signatures out of sync with the header"
    assert _ext.lib().b200_version() == 100


def test_library_is_sm100a_with_tcgen05_and_tma():
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", _ext.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):       # tcgen05.mma, TMA tensor load, tcgen05.ld
        assert mnemonic in out, mnemonic


def test_exact_roi_align_kernels_contain_no_packed_fma():
    """ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although both carry a rounding mode (it never does
    for the scalar forms), which silently breaks bit-exactness: the exact (<true, ...>) instantiations of the
    RoIAlign forward kernels must not contain an FFMA2."""
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", _ext.LIB_PATH], capture_output=True, text=True).stdout
    seen = 0
    for block in out.split("Function : ")[1:]:
        name = block.split("\n", 1)[0]
        if ("roi_align_fwd_tileILb1" in name or "roi_align_fwd_marchILb1" in name or "roi_align_fwd_genericILb1" in name):
            seen += 1
            assert "FFMA2" not in block, name
    assert seen >= 3


def test_product_never_imports_the_oracle():
    for dp, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt and "libref_cpu" not in txt, f


def test_cpu_tensors_are_rejected_everywhere():
    from cvpr22_cross_modal_pseudo_labeling_b200.layers import ROIAlign, ROIPool, embed_match_softmax, nms
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import Pooler
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList
    x = torch.zeros(1, 64, 8, 8)
    rois = torch.tensor([[0, 1.0, 1.0, 5.0, 5.0]])
    with pytest.raises(RuntimeError):
        ROIAlign((7, 7), 0.25, 2)(x, rois)
    with pytest.raises(RuntimeError):
        ROIPool((7, 7), 0.25)(x, rois)
    with pytest.raises(RuntimeError):
        nms(rois[:, 1:], torch.ones(1), 0.5)
    with pytest.raises(RuntimeError):
        embed_match_softmax(torch.zeros(4, 8), torch.zeros(2, 8))
    with pytest.raises(RuntimeError):
        Pooler((7, 7), (0.25, 0.125), 2)([x, x[:, :, ::2, ::2]], [BoxList(rois[:, 1:], (32, 32))])
    # the embedding predictor scores through the tensor-core kernel only: no einsum on the CPU
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import FastRCNNPredictor
    pred = FastRCNNPredictor({"MODEL": {"ROI_BOX_HEAD": {"EMB_DIM": 8}}}, 16).eval()
    pred.set_class_embeddings(torch.zeros(3, 8))
    with pytest.raises(RuntimeError), torch.no_grad():
        pred(torch.zeros(4, 16))
    with pytest.raises(RuntimeError):
        pred(torch.zeros(4, 16))


def test_boxlist_semantics():
    from cvpr22_cross_modal_pseudo_labeling_b200.structures import BoxList, boxlist_iou, cat_boxlist, remove_small_boxes
    b = BoxList(torch.tensor([[0., 0., 9., 9.], [5., 0., 14., 9.], [-5., -5., 700., 3.]]), (640, 480))
    b.add_field("scores", torch.tensor([0.9, 0.8, 0.7]))
    assert b.area().tolist()[:2] == [100.0, 100.0]                       # legacy +1 (bounding_box.py:226-230)
    xywh = b.convert("xywh")
    assert xywh.bbox[0].tolist() == [0, 0, 10, 10] and xywh.convert("xyxy").bbox[1].tolist() == [5, 0, 14, 9]
    iou = boxlist_iou(b, b)
    assert abs(float(iou[0, 1]) - 1 / 3) < 1e-6
    c = BoxList(b.bbox.clone(), b.size).clip_to_image(remove_empty=False)
    assert c.bbox[2].tolist() == [0, 0, 639, 3]
    assert len(remove_small_boxes(c, 5)) == 2
    both = cat_boxlist([b, b[torch.tensor([1])]])
    assert len(both) == 4 and both.get_field("scores").tolist()[-1] == pytest.approx(0.8)
    assert len(b[torch.tensor([True, False, True])]) == 2
    r = b.resize((320, 240))
    assert r.size == (320, 240) and r.bbox[0].tolist() == [0, 0, 4.5, 4.5]


def test_box_coder_matches_oracle_decode():
    import oracle
    from cvpr22_cross_modal_pseudo_labeling_b200.modeling import BoxCoder
    rng = np.random.default_rng(1)
    anchors = np.abs(rng.standard_normal((500, 4)).astype(np.float32)) * 50
    anchors[:, 2:] += anchors[:, :2] + 4
    codes = (rng.standard_normal((500, 4)) * 2).astype(np.float32)
    codes[:5, 2:] = 50.0                                                     # hits the log(1000/16) clamp
    got = BoxCoder((10., 10., 5., 5.)).decode(torch.from_numpy(codes), torch.from_numpy(anchors)).numpy()
    want = oracle.box_decode(codes, anchors, (10., 10., 5., 5.))
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-3)
    enc = BoxCoder((10., 10., 5., 5.)).encode(torch.from_numpy(want), torch.from_numpy(anchors))
    np.testing.assert_allclose(enc.numpy()[5:], codes[5:], rtol=1e-3, atol=1e-3)


def test_shard_range_partitions_everything():
    from cvpr22_cross_modal_pseudo_labeling_b200.parallel import shard_range
    for n in (0, 1, 7, 64, 65):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from cvpr22_cross_modal_pseudo_labeling_b200.parallel import all_gather_records, shard_range
dist.init_process_group("gloo", init_method="env://")
rank, ws = dist.get_rank(), dist.get_world_size()
n_img = 5
lo, hi = shard_range(n_img)
rec = torch.zeros((hi - lo, 4, 8)); cnt = torch.zeros((hi - lo,), dtype=torch.int32)
for i in range(lo, hi):
    rec[i - lo, :, 0] = i; rec[i - lo, :, 6] = 0.1 * i; cnt[i - lo] = i % 4
allr, allc = all_gather_records(rec, cnt)
assert allr.shape == (n_img, 4, 8) and allc.tolist() == [i % 4 for i in range(n_img)], (allr.shape, allc)
assert allr[:, 0, 0].tolist() == list(range(n_img))
# equal shards with known sizes: the single-collective path
rec2 = torch.full((3, 4, 8), float(rank)); cnt2 = torch.tensor([rank, rank + 1, rank + 2], dtype=torch.int32)
r2, c2 = all_gather_records(rec2, cnt2, sizes=[3] * ws)
assert r2.shape == (3 * ws, 4, 8) and c2.dtype == torch.int32
assert c2.tolist() == [0, 1, 2, 1, 2, 3] and r2[:, 0, 0].tolist() == [0.0] * 3 + [1.0] * 3
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_all_gather_records_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, REPO_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CUDA extension => ImportError naming the build command; never a silent CPU path."""
    from cvpr22_cross_modal_pseudo_labeling_b200 import _ext
    monkeypatch.setattr(_ext, "_lib", None)
    monkeypatch.setattr(_ext, "LIB_PATH", str(tmp_path / "libb200det.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _ext.lib()


def test_install_patches_the_reference_when_present():
    """install() swaps the hot-path objects into an imported maskrcnn_benchmark (only possible
    where /root/reference exists; the GPU box skips)."""
    import types
    if not os.path.isdir("/root/reference/maskrcnn_benchmark"):
        pytest.skip("reference tree not on this machine")
    amp = types.ModuleType("apex.amp")
    amp.float_function = lambda f: f
    apex = types.ModuleType("apex")
    apex.amp = amp
    saved = {k: sys.modules.get(k) for k in ("apex", "apex.amp", "maskrcnn_benchmark._C")}
    sys.modules["apex"], sys.modules["apex.amp"] = apex, amp
    sys.path.insert(0, "/root/reference")
    try:
        import maskrcnn_benchmark
        stub = types.ModuleType("maskrcnn_benchmark._C")
        stub.nms = lambda *a: None
        sys.modules["maskrcnn_benchmark._C"] = stub
        maskrcnn_benchmark._C = stub
        from cvpr22_cross_modal_pseudo_labeling_b200 import layers, modeling
        from cvpr22_cross_modal_pseudo_labeling_b200.install import install
        done = install()
        import maskrcnn_benchmark.modeling.poolers as ref_poolers
        import maskrcnn_benchmark.structures.boxlist_ops as ref_ops
        assert ref_poolers.Pooler is modeling.Pooler
        assert ref_ops._box_nms is layers.nms
        assert any(d.endswith("rpn.inference.RPNPostProcessor") for d in done)
    finally:
        sys.path.remove("/root/reference")
        for k in [m for m in sys.modules if m.startswith("maskrcnn_benchmark")]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
