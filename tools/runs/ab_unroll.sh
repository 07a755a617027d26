for v in default u8 u6 default; do
  if [ $v = default ]; then unset MFIT_LIB; else export MFIT_LIB=$PWD/gpurun_variants/libmfit_$v.so; fi
  echo "== $v"
  python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g  step %.4f ms  kernel alone %.4f ms  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
