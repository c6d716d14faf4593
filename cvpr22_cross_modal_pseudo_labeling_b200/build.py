"""In-tree nvcc build of csrc/*.cu -> libb200det.so (sm_100a only)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200det.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wall",
    "-DB200DET_BUILD=1",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force=False, verbose=False, extra=()):
    """Compile every kernel for sm_100a and link the C-ABI shared library."""
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in sources():
        obj = os.path.splitext(src)[0] + ".o"
        cmd = [NVCC, *FLAGS, *extra, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out and (verbose or p.returncode != 0):
            sys.stderr.write(out.decode(errors="replace"))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    # share torch's libcudart.so.12 (one runtime instance => one notion of "current device")
    cmd = [NVCC, "-shared", "--cudart", "shared", "-o", LIB, *objs,
           "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    ex = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    print(build(force=True, verbose=True, extra=ex))
