#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 600 > gpurun_out/r02_pytest13.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest13.log
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench13.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench13.json').read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM frac %.3f clocks %s wall %.0fs' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'], d['wall_s_total']))
s = d['sampling']; print('sampling %.1fM' % (s['value']/1e6), json.dumps(s['roofline'].get('evaluations_per_sample_layer')), 'frac %.3f' % s['roofline']['frac'], s.get('check'))
for k, v in d.get('configs', {}).items():
    print('cfg', k, '%.1fM' % (v['value']/1e6), 'ms %.1f' % v['ms_per_step'], 'e2e %.1fM' % (v['e2e']['value']/1e6) if 'e2e' in v else '', v.get('check'), v.get('sampling', {}).get('value'))
print(json.dumps(d.get('e2e_dropin'))[:600])
print(d.get('cpu_baseline'))
PY
MODE=tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4_cfg2 -f python tools/profile_grid.py > gpurun_out/r02_prof_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_row -s 1 -c 1 -o gpurun_out/r02_inv -f python tools/profile_inverse.py > gpurun_out/r02_prof_inv.log 2>&1; echo "ncu inv rc=$?"
CFG=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4_cfg3 -f python tools/profile_cfg.py > gpurun_out/r02_prof_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
CFG=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4_cfg1 -f python tools/profile_cfg.py > gpurun_out/r02_prof_cfg1.log 2>&1; echo "ncu cfg1 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
