#!/bin/bash
# final round-2 record: GPU tests, default bench, ncu captures, launch list, timeline (workers wait on the MMA mbarrier directly)
mkdir -p gpurun_out
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke33.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke33.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r02_pytest33.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest33.log
cp gpurun_out/error_table.json gpurun_out/r02_error_table.json 2>/dev/null
for rep in 1 2; do
timeout 900 python bench.py > gpurun_out/r02_bench33_$rep.json 2> gpurun_out/r02_bench33.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench33.err
python - $rep <<'PY'
import json, sys
d = json.loads(open('gpurun_out/r02_bench33_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
print('value %.1fM e2e %.1fM frac %.3f clocks %s wall %.0fs' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'], d['wall_s_total']))
s = d['sampling']; print('sampling %.1fM' % (s['value']/1e6), 'frac %.3f' % s['roofline']['frac'])
for k, v in d.get('configs', {}).items():
    print('cfg', k, '%.1fM' % (v['value']/1e6), 'ms %.1f' % v['ms_per_step'])
PY
done
MODE=tc timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4svc_cfg2 -f python tools/profile_grid.py > gpurun_out/r02_prof_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
CFG=3 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4svc_cfg3 -f python tools/profile_cfg.py > gpurun_out/r02_prof_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
CFG=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/r02_t4svc_cfg1 -f python tools/profile_cfg.py > gpurun_out/r02_prof_cfg1.log 2>&1; echo "ncu cfg1 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_service.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --config 2 > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so; cp tools/_build/t_trace.so rotationnormflow_b200/librnf_b200.so
timeout 120 python tools/tc_timeline.py > gpurun_out/r02_t4_timeline_service.txt 2>&1; tail -4 gpurun_out/r02_t4_timeline_service.txt | cut -c1-300
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
