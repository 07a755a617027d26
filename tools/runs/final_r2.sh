# Round-2 final evidence run on one B200 (every ncu line follows a plain run of the same command in the same call)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
python bench.py > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_steps20.log 2> gpurun_out/r2f_bench_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.log 2>&1
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --repeat-s 0 > gpurun_out/r2f_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --repeat-s 0 > gpurun_out/r2f_ncu_launch.log 2>&1
python tools/kbench.py --config C2 --reps 3 > gpurun_out/r2f_plain_kb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mfit_eval_kernel -s 4 -c 1 -o /tmp/prof_r2f python tools/kbench.py --config C2 --reps 3 > gpurun_out/r2f_ncu_full.log 2>&1
ncu -i /tmp/prof_r2f.ncu-rep --page raw --csv > gpurun_out/r2f_eval_kernel_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2f.ncu-rep --page source --csv > gpurun_out/r2f_eval_kernel_ncu_source.csv 2>/dev/null
python tools/kbench.py --config NVSET --reps 3 > gpurun_out/r2f_plain_kb_nvset.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:perchain -s 2 -c 1 -o /tmp/prof_r2f_pc python tools/kbench.py --config NVSET --reps 3 > gpurun_out/r2f_ncu_perchain.log 2>&1
ncu -i /tmp/prof_r2f_pc.ncu-rep --page raw --csv > gpurun_out/r2f_perchain_kernel_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2f_pc.ncu-rep --page source --csv > gpurun_out/r2f_perchain_kernel_ncu_source.csv 2>/dev/null
for d in 1 2; do MFIT_PC_DIAG=$d python tools/kbench.py --config NVSET --reps 5 2>&1 | head -1; done > gpurun_out/r2f_nvset_diag.log
tail -n 2 gpurun_out/r2f_plain_kb.log gpurun_out/r2f_plain_kb_nvset.log
tail -c 600 gpurun_out/r2f_bench_steps20.log
