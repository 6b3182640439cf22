#!/bin/bash
mkdir -p gpurun_out
PREFIX=i_ bash tools/ab_ncu_inv.sh 2>&1 | tee gpurun_out/ab_ncu_inv4.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 -k "inverse or full_size or edge" 2>&1 | tail -3
timeout 300 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); s = d.get('sampling') or d
print('config 4: %.1f M samples/s' % (s['value'] / 1e6), s.get('roofline', {}).get('evaluations_per_sample_layer'), s.get('check'))"
