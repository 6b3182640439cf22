#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r02_bench_n${N}_service.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
tail -2 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_n${N}_service.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value %.1fM' % (d['value']/1e6), 'ms/step %.1f' % d['ms_per_step'], d['scaling'], 'e2e', d.get('e2e'), 'sampling %.1fM' % (d['sampling']['value']/1e6) if d.get('sampling') else None, d['clocks'])
PY
