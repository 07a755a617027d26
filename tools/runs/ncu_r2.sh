# ncu evidence for round 2 (one box): plain runs first, then the same command lines under ncu; reports are exported to CSV
# on the box (raw page + per-instruction source page) so that what comes back stays small
set -x
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --repeat-s 0 > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --repeat-s 0 > gpurun_out/r2_ncu_launch.log 2>&1
python tools/kbench.py --config C2 --reps 3 > gpurun_out/r2_plain_kb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mfit_eval_kernel -s 4 -c 1 -o /tmp/prof_r2_final python tools/kbench.py --config C2 --reps 3 > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/prof_r2_final.ncu-rep --page raw --csv > gpurun_out/r2_eval_kernel_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2_final.ncu-rep --page source --csv > gpurun_out/r2_eval_kernel_ncu_source.csv 2>/dev/null
python tools/kbench.py --config NVSET --reps 3 > gpurun_out/r2_plain_kb_nvset.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:perchain -s 2 -c 1 -o /tmp/prof_r2_perchain python tools/kbench.py --config NVSET --reps 3 > gpurun_out/r2_ncu_perchain.log 2>&1
ncu -i /tmp/prof_r2_perchain.ncu-rep --page raw --csv > gpurun_out/r2_perchain_kernel_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2_perchain.ncu-rep --page source --csv > gpurun_out/r2_perchain_kernel_ncu_source.csv 2>/dev/null
tail -n 2 gpurun_out/r2_plain_kb.log gpurun_out/r2_plain_kb_nvset.log
ls -la gpurun_out
