#!/bin/bash
mkdir -p gpurun_out
echo skip pytest

timeout 900 python bench.py > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench7.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench7.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM frac %.3f clocks %s wall %.0fs' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'], d['wall_s_total']))
print('sampling %.1fM e2e %.1fM frac %.3f' % (d['sampling']['value']/1e6, d['sampling']['e2e']['value']/1e6, d['sampling']['roofline']['frac']), d['sampling']['check'])
for k, v in d.get('configs', {}).items():
    print('cfg', k, '%.1fM' % (v['value']/1e6), 'ms %.1f' % v['ms_per_step'], 'e2e %.1fM' % (v['e2e']['value']/1e6) if 'e2e' in v else '', v.get('check'), v.get('sampling', {}).get('value'))
print(d.get('cpu_baseline'), d.get('eager_torch_gpu_baseline'))
PY
