for v in default w4c3 w6c2; do
  if [ $v = default ]; then unset MFIT_LIB; else export MFIT_LIB=$PWD/gpurun_variants/libmfit_$v.so; fi
  echo "== $v"
  python tools/kbench.py --config C2 --reps 10 2>&1 | tail -1
  for st in 1000 20; do
  python bench.py --steps $st --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('steps $st: value %.4g  step %.4f ms  kernel alone %.4f ms  frac %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))"
  done
done
unset MFIT_LIB
python tools/kbench.py --config C2 --reps 3 > gpurun_out/plain_kb.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mfit_eval_kernel -s 4 -c 1 -o gpurun_out/prof_r2a python tools/kbench.py --config C2 --reps 3 > gpurun_out/ncu_r2a.log 2>&1; tail -2 gpurun_out/ncu_r2a.log
