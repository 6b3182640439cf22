#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s --timeout 300 > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest8.log; grep "full_grid\|config2\|config3" gpurun_out/r02_pytest8.log | head
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench8.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM frac %.3f wall %.0fs' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['wall_s_total']))
print(json.dumps(d.get('e2e_dropin'), indent=1))
PY
