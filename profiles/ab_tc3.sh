set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "golden or oracle or tile_edges or shard or tma or pipelined or full_size or edge or ffma_path" > gpurun_out/pytest_tc3.log 2>&1; tail -4 gpurun_out/pytest_tc3.log
for i in 1 2; do
timeout 300 python profiles/quick_time.py 1000000 mc,fwd 2>&1 | tail -6
B200PINN_NO_TC3=1 timeout 300 python profiles/quick_time.py 1000000 mc,fwd 2>&1 | tail -6
done
