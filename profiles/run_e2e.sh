set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined or get_mc" > gpurun_out/pytest_pipe.log 2>&1; tail -3 gpurun_out/pytest_pipe.log
timeout 900 python bench.py --no-c1 --no-eager > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1]);print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
