#!/bin/bash
# cycle-based A/B of the inverse kernel (tools/profile_inverse.py: symsol2, 16 x 32768 samples), one profiled launch per variant
PREFIX=${PREFIX:-i_}
mkdir -p gpurun_out
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so
for v in tools/_build/${PREFIX}*.so; do
  cp "$v" rotationnormflow_b200/librnf_b200.so
  timeout ${TMO:-90} ncu --metrics sm__cycles_elapsed.max,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:flow_row -s 1 -c 1 --csv python tools/profile_inverse.py > gpurun_out/ab_ncu_tmp.csv 2>&1
  python - "$v" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open('gpurun_out/ab_ncu_tmp.csv')) if len(r) > 10]
vals = {r[-3]: r[-1] for r in rows[1:]} if rows else {}
last = [l for l in open('gpurun_out/ab_ncu_tmp.csv').read().splitlines() if l.startswith('torch.Size')]
try:
    print(f"{sys.argv[1].split('/')[-1]:24s} cycles {float(vals['sm__cycles_elapsed.max'].replace(',',''))/1e6:8.3f} M  inst {float(vals['smsp__inst_executed.sum'].replace(',',''))/1e9:6.3f} G  issue {vals['smsp__issue_active.avg.pct_of_peak_sustained_active']}  {last[-1] if last else 'NO OUTPUT'}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('gpurun_out/ab_ncu_tmp.csv').read()[-300:])
PY
done
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
