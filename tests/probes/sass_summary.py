"""Static SASS evidence of the shipped library: tcgen05 / TMEM / TMA mnemonics per kernel.

    python tests/probes/sass_summary.py > profiles/r2_sass_summary.md      (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "slowfast_b200", "csrc", "libsfb200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "LDGSTS", "UBLKCP", "HMMA")
TOTAL_KEYS = KEYS + ("REDG", "SYNCS", "FFMA", "MUFU.EX2")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o).replace("void ", "") for o in out[:len(names)]]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout.split("\n")
    kernels, cur = collections.OrderedDict(), None
    totals = collections.Counter()
    for line in sass:
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            if op.startswith(KEYS):
                kernels[cur][op] += 1
            if op.startswith(TOTAL_KEYS):
                totals[".".join(op.split(".")[:4])] += 1
    names = demangle(list(kernels))
    print("# SASS evidence of the shipped `libsfb200.so` (sm_100a): tcgen05 / TMEM / TMA mnemonics per kernel\n")
    print("Produced by `python tests/probes/sass_summary.py` (`cuobjdump -sass slowfast_b200/csrc/libsfb200.so`, after the last "
          "rebuild); counts are static instruction counts per kernel.  `UTCHMMA` = `tcgen05.mma kind::f16`, `LDTM` = "
          "`tcgen05.ld`, `UTMALDG` = TMA loads (`cp.async.bulk.tensor`, `.IM2COL` = im2col mode), `UTCBAR` = `tcgen05.commit` "
          "-> mbarrier, `LDGSTS` = `cp.async`.  No `HMMA` (legacy `mma.sync`) anywhere.\n")
    print("| kernel | tensor-core / TMEM / TMA mnemonics |\n|---|---|")
    for (mangled, ops), name in zip(kernels.items(), names):
        if ops:
            print(f"| `{name}` | " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())) + " |")
    print("\nTotals over the library: " + ", ".join(f"{k} x{v}" for k, v in sorted(totals.items())) + ".")
    print(f"\nKernels in the library: {len(kernels)}.")
    assert not any(k.startswith("HMMA") for k in totals), "legacy mma.sync found"


if __name__ == "__main__":
    sys.exit(main())
