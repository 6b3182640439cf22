#!/bin/bash
# A/B helper for the GPU box: bench once per prebuilt library variant tools/_build/${PREFIX}*.so in mode $MODE
PREFIX=${PREFIX:-x_}
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so
for v in tools/_build/${PREFIX}*.so; do
  cp "$v" rotationnormflow_b200/librnf_b200.so
  printf "%s [%s]: " "$(basename $v)" "${MODE:-tc}"
  timeout ${TMO:-120} python bench.py --mode ${MODE:-tc} --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_err.log | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M rot/s', 'e2e', round(d['e2e']['value']/1e6,2), 'sampling', round(d['sampling']['value']/1e6,2) if d.get('sampling') else None, 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('FAILED', e)"
done
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
