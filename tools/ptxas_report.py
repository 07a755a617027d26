"""Print registers / spills per kernel from csrc/build/*.ptxas.log"""
import re, subprocess, sys
t = open(sys.argv[1]).read()
pat = sys.argv[2:] 
for b in re.split(r"ptxas info\s+: Compiling entry function '", t)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat and not any(k in dem for k in pat):
        continue
    regs = re.search(r"Used (\d+) registers", b).group(1)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", b)
    print(regs, sp.group(1), sp.group(2), dem[:90])
