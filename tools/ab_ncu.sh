#!/bin/bash
# A/B by CYCLES (clock independent, unlike a bench run under the power cap): one profiled flow_t4 launch (config 2, one image of the
# level-5 grid) per prebuilt library variant tools/_build/${PREFIX}*.so
PREFIX=${PREFIX:-y_}
mkdir -p gpurun_out
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so
for v in tools/_build/${PREFIX}*.so; do
  cp "$v" rotationnormflow_b200/librnf_b200.so
  MODE=tc timeout ${TMO:-90} ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:flow_t4 -s 1 -c 1 --csv python tools/profile_grid.py > gpurun_out/ab_ncu_tmp.csv 2>&1
  python - "$v" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open('gpurun_out/ab_ncu_tmp.csv')) if len(r) > 10]
vals = {r[-3]: r[-1] for r in rows[1:]} if rows else {}
last = [l for l in open('gpurun_out/ab_ncu_tmp.csv').read().splitlines() if l.startswith('tc ')]
try:
    cyc = float(vals['sm__cycles_elapsed.max'].replace(',', '')); ms = float(vals['gpu__time_duration.sum'].replace(',', ''))
    print(f"{sys.argv[1].split('/')[-1]:24s} cycles {cyc/1e6:8.3f} M  time {ms:8.4f} {''}  inst {float(vals['smsp__inst_executed.sum'].replace(',',''))/1e9:6.3f} G  tensor {vals['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']} issue {vals['smsp__issue_active.avg.pct_of_peak_sustained_active']}  {last[-1] if last else 'NO OUTPUT'}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/ab_ncu_tmp.csv').read()[-300:])
PY
done
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
