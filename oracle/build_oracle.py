"""Build recipe for the oracle (TEST INFRASTRUCTURE, never shipped).

  oracle/liboracle.so        gcc build of roi_oracle.c, the plain-C restatement.
  oracle/_ref/libref_cpu.so  the REFERENCE's own csrc/cpu/{ROIAlign_cpu,nms_cpu}.cpp,
                             compiled from where they lie under /root/reference
                             (only possible in the build container; the GPU box
                             uses the prebuilt file that travels with the repo).

No -march / -ffast-math: x86-64 baseline has no FMA, so both builds round every
multiply and add separately, like the reference's shipped CPU build
(/root/reference/setup.py:20-54 passes no arch flags either).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_CSRC = "/root/reference/maskrcnn_benchmark/csrc"


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(verbose=False):
    src = os.path.join(HERE, "roi_oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if _newer(out, [src]):
        return out
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off",
           "-fno-fast-math", "-Wall", "-Wextra", src, "-o", out, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


def build_ref(verbose=False):
    """Returns the path of libref_cpu.so, or None when it neither exists nor
    can be built (no /root/reference on this machine)."""
    out_dir = os.path.join(HERE, "_ref")
    out = os.path.join(out_dir, "libref_cpu.so")
    shim = os.path.join(HERE, "ref_shim.cpp")
    ref_srcs = [os.path.join(REFERENCE_CSRC, "cpu", f)
                for f in ("ROIAlign_cpu.cpp", "nms_cpu.cpp", "vision.h")]
    have_ref = all(os.path.exists(p) for p in ref_srcs)
    if not have_ref:
        return out if os.path.exists(out) else None
    if _newer(out, [shim] + ref_srcs):
        return out
    import torch
    from torch.utils import cpp_extension
    import pybind11
    os.makedirs(out_dir, exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-w",
           "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           "-I" + REFERENCE_CSRC]
    for inc in cpp_extension.include_paths():
        cmd.append("-I" + inc)
    cmd += ["-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(),
            shim, "-o", out, "-L" + tlib, "-ltorch", "-ltorch_cpu", "-lc10",
            "-Wl,-rpath," + tlib]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_oracle(verbose=True))
    print(build_ref(verbose=True))
    sys.exit(0)
