#!/bin/bash
mkdir -p gpurun_out
PREFIX=i_ bash tools/ab_ncu_inv.sh 2>&1 | tee gpurun_out/ab_ncu_inv3.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 200 -k "inverse or full_size" 2>&1 | tail -12
python - <<'PY'
import json
d = json.load(open('gpurun_out/error_table.json'))
for k, v in d.items():
    if 'inverse' in k and ('symsol' in k or 'modelnet' in k or 'raw' in k) and 'tc]' in k or (isinstance(v, dict) and 'frac' in str(v) and 'inverse' in k and 'tc' in k):
        print(k, json.dumps(v)[:260])
PY
