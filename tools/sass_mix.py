"""Instruction mix of the hottest loop of a kernel: python tools/sass_mix.py <object> <mangled kernel name>
(the hottest loop = the backward branch whose body holds the most FP64 instructions)."""
import re, subprocess, sys
from collections import Counter
out = subprocess.run(["cuobjdump", "-sass", "-fun", sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", l).strip() for l in out.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
addr = lambda l: int(re.match(r"/\*([0-9a-f]{4})\*/", l).group(1), 16)
best = None
cands = []
for i, l in enumerate(lines):
    m = re.search(r"BRA(\.U)? (!?U?P\d+, )?0x([0-9a-f]+)", l)
    if m and int(m.group(3), 16) < addr(l):
        tgt = int(m.group(3), 16)
        body = [b for b in lines if tgt <= addr(b) <= addr(l)]
        fp = sum(1 for b in body if re.search(r"\b(DFMA|DADD|DMUL|DSETP)\b", b))
        inner = sum(1 for b in body[:-1] if (lambda mm: mm and int(mm.group(3), 16) < addr(b))(re.search(r"BRA(\.U)? (!?U?P\d+, )?0x([0-9a-f]+)", b)))
        if inner:
            continue
        cands.append((fp, body))
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cands = [c for c in cands if c[0] >= 8]
print("innermost FP64 loops:", [(hex(addr(b[0])), len(b), f) for f, b in cands])
fp, body = cands[which]
c = Counter(re.sub(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?", "", b).split(" ")[0].split(".")[0] for b in body)
lds = sum("LDS.128" in b for b in body)
print("loop: %d instr, fp64 %d, other %d, LDS.128 %d -> points/iter ~%d" % (len(body), fp, len(body) - fp, lds, lds * 2 // 3 if lds else 0))
print(c.most_common())
if len(sys.argv) > 4:
    print("\n".join(body))
