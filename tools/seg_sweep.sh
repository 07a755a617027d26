for t in 8 6 5 7 10 3; do
  MFIT_SEG_TILES=$t python bench.py --steps 600 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tiles $t: step %.4f ms  kernel alone %.4f ms  frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
