#!/bin/bash
for v in 2 1; do
SBV_P256_VARIANT=$v timeout 600 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.read())
print('variant $v value',b['value'],'e2e',b['e2e']['value'],b['e2e']['single_caller_value'],'lat',b['step_latency_ms'],'roof',b['roofline']['frac'],b['roofline']['kernel_ms'])"
done
SBV_P256_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "p256 or c2 or exceptional or rfc" 2>&1 | tail -2
