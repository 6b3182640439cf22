#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke36.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke36.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest36.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest36.log
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02_bench36.json 2> gpurun_out/r02_bench36.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench36.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench36.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM frac %.3f clocks %s wall %.0fs' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'], d['wall_s_total']))
s = d['sampling']; print('sampling %.1fM' % (s['value']/1e6), 'frac %.3f' % s['roofline']['frac'], s.get('check'))
for k, v in d.get('configs', {}).items():
    print('cfg', k, '%.1fM' % (v['value']/1e6), 'ms %.1f' % v['ms_per_step'], v.get('sampling', {}).get('value'))
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
