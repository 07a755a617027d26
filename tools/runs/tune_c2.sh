# in-step sweep of the segment size and an unroll variant on C2 (one box): tools/runs/tune_c2.sh
run() { python bench.py --steps 300 --warmup 5 --no-cpu --no-extras --repeat-s 0.3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1: value %.4g  step %.4f ms (min %.4f)  kernel alone %.4f ms  frac %.3f' % (d['value'], d['ms_per_step'], d['blocks']['min_ms']/d['steps'], d['roofline']['kernel_ms'], d['roofline']['frac']))"; }
for t in 3 4 5 6 8; do MFIT_SEG_TILES=$t run "seg_tiles=$t"; done
for g in 2 3 4; do MFIT_GROUPS=$g run "groups=$g"; done
for v in "$@"; do MFIT_LIB=$PWD/gpurun_variants/libmfit_$v.so run "variant $v"; done
