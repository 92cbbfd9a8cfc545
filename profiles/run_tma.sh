set -x
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "tma or golden or tile_edges or shard or wide" > gpurun_out/pytest_tma.log 2>&1; tail -5 gpurun_out/pytest_tma.log
timeout 300 python profiles/quick_time.py 1000000 mc 2>&1 | tail -6
timeout 200 python profiles/wide_res_check.py time 2>&1 | tail -2
