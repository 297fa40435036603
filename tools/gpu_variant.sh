#!/bin/bash
# A/B of the k_verify_coz variants (SBV_P256_VARIANT): 1 default, 2 lockstep, 3 doubling as one out-of-line unit,
# 4 six blocks/SM (no spills), 5 sixteen signatures per inversion in k_prep
for v in ${VARIANTS:-1 3 4 5}; do
SBV_P256_VARIANT=$v timeout 600 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.read())
print('variant $v value',round(b['value']/1e6,2),'e2e',round(b['e2e']['value']/1e6,2),round(b['e2e']['single_caller_value']/1e6,2),'lat',round(b['step_latency_ms'],3),'roof',round(b['roofline']['frac'],3),'k_ms',round(b['roofline']['kernel_ms'],3),'prep',round(b['roofline']['prep_kernel_ms'],3))"
done
for v in ${VARIANTS:-1 3 4 5}; do SBV_P256_VARIANT=$v timeout 300 python -m pytest tests/test_golden.py tests/test_gpu_parity.py -m gpu -x -q -k "golden or p256 or exceptional or rfc" 2>&1 | tail -1; done
