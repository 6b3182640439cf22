#!/bin/bash
# A/B helper for the GPU box: run the bench once per prebuilt library variant in tools/_build/v_*.so (built locally with
# RNF_NVCC_EXTRA=... python -m rotationnormflow_b200.build --force; cp rotationnormflow_b200/librnf_b200.so tools/_build/v_NAME.so)
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product.so
for v in tools/_build/v_*.so; do
  cp "$v" rotationnormflow_b200/librnf_b200.so
  printf "%s: " "$(basename $v)"
  timeout 200 python bench.py --mode ${MODE:-tc} --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2), 'M rot/s', 'sampling', round(d['sampling']['value']/1e6,2) if d.get('sampling') else None)"
done
cp tools/_build/.product.so rotationnormflow_b200/librnf_b200.so
