"""Opcode histogram of the kernels in libglc_b200.so (cuobjdump -sass): the evidence behind "EXACT mode never
contracts a multiply-add" (exact_gemm_kernel / imdct_sparse_kernel: FFMA = 0, and every packed FFMA2 is either the
exact multiply a * t + (-0.0) or the exact add p * 1.0 + acc, its constant in a uniform register; operands arrive by
UBLKCP = cp.async.bulk, SYNCS = mbarrier).

    python tools/sass_histogram.py > profiles/r2_sass_opcode_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gapless_lossy_codec_b200", "libglc_b200.so")


def histogram(so=SO):
    """{kernel name: Counter of opcodes (+ the FFMA2 forms x2mul / x2add / x2fused)}"""
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
            if m.group(1) == "FFMA2":
                # FFMA2 Rd, Ra, Rb, Rc: the kernels' two exact forms carry their constant in a uniform register
                ops = [o.strip() for o in line.split("FFMA2", 1)[1].split(";")[0].split(",")]
                if len(ops) == 4 and ops[3].startswith("UR"):
                    hist[kern]["x2mul"] += 1  # a * t + UR(-0.0)
                elif len(ops) == 4 and ops[2].startswith("UR"):
                    hist[kern]["x2add"] += 1  # p * UR(1.0) + acc
                else:
                    hist[kern]["x2fused"] += 1  # three vector operands: a contracted multiply-add
    return hist


def main():
    hist = histogram()
    want = ["FMUL", "FADD", "FFMA", "FFMA2", "x2mul", "x2add", "x2fused", "FMNMX", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "ATOMS", "F2I", "MUFU"]
    print(f"{'kernel':48s} {'total':>7s} " + " ".join(f"{w:>6s}" for w in want))
    for k in sorted(hist):
        h = hist[k]
        print(f"{k[:48]:48s} {sum(h.values()):7d} " + " ".join(f"{h.get(w, 0):6d}" for w in want))
    print("\nThe EXACT transform kernels (exact_gemm_kernel, imdct_sparse_kernel) must show FFMA = 0 and x2fused = 0: Rust never"
          "\ncontracts a*b+c, and a fused multiply-add rounds once and flips quantised indices (SURVEY.md section 0 F2).  Their"
          "\nmultiply-adds are packed f32x2 instructions in two exact forms: x2mul = FFMA2 a, t, UR(-0.0) (the product, rounded,"
          "\nsign of zero kept) and x2add = FFMA2 p, UR(1.0), acc (the sum, rounded) -- two roundings like FMUL + FADD; ptxas"
          "\ncontracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 when it can see the constants, so they arrive as kernel"
          "\nparameters.  x2mul must equal x2add.  The operands"
          "\narrive by UBLKCP (cp.async.bulk) on SYNCS (mbarrier) rings.  The FFMA in quant_pack / dequant_tile / ola /"
          "\ngather_raw belong to the correctly rounded IEEE division, square-root and rounding sequences (__fdiv_rn, sqrtf,"
          "\nroundf), not to contracted user arithmetic (the files are compiled with -fmad=false; results are bit-equal to"
          "\nthe oracle).  FAST-mode kernels (fast_*) are tolerance class and do use FFMA.")


if __name__ == "__main__":
    main()
