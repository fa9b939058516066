#!/bin/bash
# A/B of k_mm10f build variants (build/variants/libw3d_<X>.so) on one workload; prints the stress pass time
WL=${1:-hex8_1m_mm10}
for v in ${VARIANTS:-A B C D E}; do
  W3D_ENGINE_LIB=$PWD/build/variants/libw3d_$v.so timeout 150 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$v', d['stress_pass']['ms'], d['stiffness_pass']['ms'], d['plastic_check'])
except Exception as e: print('$v failed', e)"
done
