"""print the headline fields of a bench.py JSON line read from stdin: python bench.py ... | python tools/bench_line.py [label]"""
import json
import sys

line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
r = d.get("roofline", {})
print(" ".join(sys.argv[1:]), "value %.4g  ms/step %.4f  e2e %.4g  kernel_ms %.4f  frac %.3f  launches %s"
      % (d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_ms", float("nan")), r.get("frac", float("nan")),
         d.get("gpu_launches")))
