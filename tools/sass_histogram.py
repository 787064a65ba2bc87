"""Opcode histogram of the kernels in libglc_b200.so (cuobjdump -sass): the evidence behind "EXACT mode has no FMA"
(FFMA = 0 in exact_gemm_kernel / imdct_sparse_kernel, operands arrive by UBLKCP = cp.async.bulk, SYNCS = mbarrier).

    python tools/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gapless_lossy_codec_b200", "libglc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\(.*", "", kern).replace("glc::", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
want = ["FMUL", "FADD", "FFMA", "FMNMX", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "ATOMS", "F2I", "MUFU"]
print(f"{'kernel':48s} {'total':>7s} " + " ".join(f"{w:>6s}" for w in want))
for k in sorted(hist):
    h = hist[k]
    print(f"{k[:48]:48s} {sum(h.values()):7d} " + " ".join(f"{h.get(w, 0):6d}" for w in want))
print("\nThe EXACT transform kernels (exact_gemm_kernel, imdct_sparse_kernel) must show FFMA = 0: Rust never contracts"
      "\na*b+c, and a fused multiply-add rounds once and flips quantised indices (SURVEY.md section 0 F2); their operands"
      "\narrive by UBLKCP (cp.async.bulk) on SYNCS (mbarrier) rings.  The FFMA in quant_pack / dequant_tile / ola /"
      "\ngather_raw belong to the correctly rounded IEEE division, square-root and rounding sequences (__fdiv_rn, sqrtf,"
      "\nroundf), not to contracted user arithmetic (the files are compiled with -fmad=false; results are bit-equal to"
      "\nthe oracle).  FAST-mode kernels (fast_*) are tolerance class and do use FFMA.")
