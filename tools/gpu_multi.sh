#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multi_device" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2>/dev/null; cat gpurun_out/bench_n1.json | python -c "import json,sys; b=json.load(sys.stdin); print('N=1', b['value'], b['e2e']['value'])"
