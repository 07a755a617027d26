"""SASS evidence for a likelihood kernel: TMA / mbarrier instruction counts, the instruction mix of every innermost
evaluation loop, and the full text of the loop with the fewest FP64 instructions per evaluation (what adapted chains run).

  python tools/sass_listing.py lisp_mcmc_b200/csrc/build/eval_builtin_normal.o _Z16mfit_eval_kernelI12MfitGaussLinLi0EEv13MfitMultiArgs \
      "mfit_eval_kernel<MfitGaussLin, log-normal>" 16 > profiles/r02_sass_gauss_lin.txt
(last argument: evaluations per loop iteration, to turn FP64 counts into FP64 instructions per evaluation)"""
import re
import subprocess
import sys
from collections import Counter

obj, fun, title = sys.argv[1], sys.argv[2], sys.argv[3]
per_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 16
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in out.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
addr = lambda l: int(re.match(r"\s*/\*([0-9a-f]{4,5})\*/", l).group(1), 16)
op = lambda l: re.sub(r"\s*/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?", "", l).split(" ")[0].split(".")[0]
bra = lambda l: re.search(r"BRA(\.U)? (!?U?P\d+, )?0x([0-9a-f]+)", l)
print("%s (sm_100a, cuobjdump -sass of %s): %d instructions" % (title, obj.split("csrc/")[-1], len(lines)))
print("TMA / mbarrier: %d UBLKCP, %d SYNCS" % (sum("UBLKCP" in l for l in lines), sum("SYNCS" in l for l in lines)))
loops = []
for l in lines:
    m = bra(l)
    if m and int(m.group(3), 16) < addr(l):
        tgt = int(m.group(3), 16)
        body = [b for b in lines if tgt <= addr(b) <= addr(l)]
        if any(bra(b) and int(bra(b).group(3), 16) < addr(b) for b in body[:-1]):
            continue
        fp = sum(op(b) in ("DFMA", "DADD", "DMUL", "DSETP") for b in body)
        if fp >= 8 * per_iter // 2:
            loops.append((tgt, fp, body))
print("\ninnermost evaluation loops (%d evaluations per iteration)" % per_iter)
for tgt, fp, body in loops:
    c = Counter(op(b) for b in body)
    print("  loop at 0x%05x: %d instructions, %d FP64 (%.2f per evaluation, %.2f instructions per evaluation): %s"
          % (tgt, len(body), fp, fp / per_iter, len(body) / per_iter, ", ".join("%s %d" % kv for kv in c.most_common())))
if loops:
    tgt, fp, body = min(loops, key=lambda t: (t[1], len(t[2])))
    print("\nthe loop with the fewest FP64 instructions (0x%05x), in full:" % tgt)
    print("\n".join(body))
