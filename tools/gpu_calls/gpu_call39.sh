#!/bin/bash
# ambiguity band of the replay: cycles + parity statistics per setting
mkdir -p gpurun_out
PREFIX=i_ bash tools/ab_ncu_inv.sh 2>&1 | tee gpurun_out/ab_ncu_inv5.txt
cp rotationnormflow_b200/librnf_b200.so tools/_build/.product2.so
for v in tools/_build/i_*.so; do
  cp "$v" rotationnormflow_b200/librnf_b200.so
  RNF_TEST_MODES=tc timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "inverse_parity" 2>&1 | tail -1
  python - "$v" <<'PY'
import json, sys
t = json.load(open('gpurun_out/error_table.json'))
rows = [e for e in t if e.get('test') == 'inverse' and e.get('mode') == 'tc' and e['case'] in ('raw', 'symsol2048', 'symsol2', 'modelnet', 's_unrot', 's_clu')]
print(sys.argv[1].split('/')[-1], ' | '.join('%s fp64 %.4f fp32 %.4f worst %.1e' % (e['case'], e['rows_within_1e5_vs_ref_fp64'], e['rows_within_1e5_vs_ref_fp32'], e['worst_row']) for e in rows))
PY
done
cp tools/_build/.product2.so rotationnormflow_b200/librnf_b200.so
