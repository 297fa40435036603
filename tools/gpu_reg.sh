#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "registered or fused" 2>&1 | tail -2
timeout 600 python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.read())
print('value',b['value'],'e2e',b['e2e']['value'],'lat',b['step_latency_ms'],'reg',b['registered_keys'])"
