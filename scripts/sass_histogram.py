"""Per-kernel SASS mnemonic histogram of the in-tree libb200det.so (cuobjdump -sass): which kernels carry tcgen05
(UTCHMMA / UTCBAR / LDTM / STTM), TMA (UTMALDG / UBLKCP / UBLKRED), mbarrier (SYNCS), packed fp32 (FFMA2 / FMUL2 /
FADD2), reductions (RED) ...  Writes profiles/<tag>_sass_histogram.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cvpr22_cross_modal_pseudo_labeling_b200", "libb200det.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "UBLKPF", "SYNCS", "FFMA2",
         "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "RED", "ATOMS", "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "BAR",
         "HMMA", "IMAD", "LEA", "BRA"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", out)
    if names:
        d = subprocess.run(["/usr/local/cuda/bin/cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, d))
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    lines = ["SASS mnemonic histogram of libb200det.so (%s); columns: instructions, then the watched mnemonics that occur"
             % ", ".join(arch), ""]
    total = collections.Counter()
    for block in out.split("Function : ")[1:]:
        name = block.split("\n", 1)[0].strip()
        ops = collections.Counter()
        n = 0
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", block, flags=re.M):
            op = m.group(1)
            n += 1
            ops[op] += 1
        total.update(ops)
        short = demangle.get(name, name)
        short = re.sub(r"\(anonymous namespace\)::|b200::|<unnamed>::", "", short)
        cut = short.find(">(")
        short = (short[:cut + 1] if cut >= 0 else short.split("(")[0]).replace("(int)", "").replace("(bool)", "")[:100]
        watched = ", ".join("%s %d" % (w, ops[w]) for w in WATCH if ops.get(w))
        lines.append("%-100s %6d  %s" % (short, n, watched))
    lines.append("")
    lines.append("whole library: " + ", ".join("%s %d" % (w, total[w]) for w in WATCH if total.get(w)))
    path = os.path.join(ROOT, "profiles", "%s_sass_histogram.txt" % tag)
    open(path, "w").write("\n".join(lines) + "\n")
    print(path, len(lines) - 4, "kernels")


if __name__ == "__main__":
    main()
