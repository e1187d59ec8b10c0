"""Opcode histogram of every kernel in libegonn_b200.so (cuobjdump -sass): the evidence that the hot kernels are
Blackwell-native (UTCHMMA = tcgen05.mma, STTM / LDTM = tcgen05.st / ld, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, LDG.E.*256 = 256-bit gathers) and that no library kernel is linked in.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "egonn_b200", "csrc", "libegonn_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "STTM", "LDTM", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "LDG.E.ENL2.256", "LDG", "STG", "LDS", "STS", "ATOM",
         "ATOMS", "RED", "MATCH", "SHFL", "BAR", "FFMA", "HMMA", "IMAD"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op] += 1
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(LIB, REPO)}: {len(kernels)} kernels, opcode counts per kernel (cuobjdump -sass, sm_100a)")
    print("# columns: total instructions | " + " ".join(WATCH))
    total = collections.Counter()
    for k, c in kernels.items():
        row = []
        for w in WATCH:
            n = sum(v for op, v in c.items() if op == w or op.startswith(w + "."))
            row.append(n)
            total[w] += n
        short = re.sub(r"\((?!anonymous).*", "", names.get(k, k))
        print(f"{short:90s} {sum(c.values()):6d} | " + " ".join(f"{w}={n}" for w, n in zip(WATCH, row) if n))
    print("# whole library: " + " ".join(f"{w}={total[w]}" for w in WATCH if total[w]))
    libs = [k for k in kernels if "cub" in k.lower() or "thrust" in k.lower() or "cutlass" in k.lower()]
    print(f"# library kernels (cub / thrust / cutlass) in the image: {len(libs)}")


if __name__ == "__main__":
    main()
