set -x
timeout 300 python profiles/wide_res_check.py > gpurun_out/wide_res_check.log 2>&1; tail -8 gpurun_out/wide_res_check.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "wide or long_sweep or golden or oracle or edge" > gpurun_out/pytest_wide.log 2>&1; tail -5 gpurun_out/pytest_wide.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wide_res_ts --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_wide_res python profiles/kernel_once.py wide 262144 > gpurun_out/ncu_wide.log 2>&1; tail -3 gpurun_out/ncu_wide.log
