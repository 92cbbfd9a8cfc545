set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -c 300 gpurun_out/bench_8gpu.err; wc -c gpurun_out/bench_8gpu.json
