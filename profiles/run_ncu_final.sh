# ncu evidence of the final build: one --set full capture per headline kernel + the launch list of a short bench run
set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc3_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_mc3_final python profiles/kernel_once.py mc > gpurun_out/ncu_mc.log 2>&1; tail -2 gpurun_out/ncu_mc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_res_ts --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_wide_res_final python profiles/kernel_once.py wide 262144 > gpurun_out/ncu_wide.log 2>&1; tail -2 gpurun_out/ncu_wide.log
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-eager > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -c 300 gpurun_out/bench_short.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-eager > gpurun_out/ncu_bench.log 2>&1; tail -c 300 gpurun_out/ncu_bench.log
