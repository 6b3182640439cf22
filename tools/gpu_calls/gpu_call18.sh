#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r02_pytest18.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest18.log
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02_bench18.json 2> gpurun_out/r02_bench18.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench18.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench18.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM frac %.3f clocks %s wall %.0fs' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'], d['wall_s_total']))
s = d['sampling']; print('sampling %.1fM' % (s['value']/1e6), 'frac %.3f' % s['roofline']['frac'])
for k, v in d.get('configs', {}).items():
    print('cfg', k, '%.1fM' % (v['value']/1e6), 'ms %.1f' % v['ms_per_step'])
print(json.dumps(d.get('train_step')))
PY
