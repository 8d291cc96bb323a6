#!/bin/bash
# tools/scale_run.sh MAXN: bench.py at N=1,2,4,..,MAXN for both workloads (run on a multi-GPU box)
MAXN=${1:-8}
mkdir -p gpurun_out
for wl in config2 config5; do
  for n in 1 2 4 8; do
    [ $n -gt $MAXN ] && continue
    extra="--no-extra"; [ $wl = config5 ] && extra="--workload config5 --log2-powers ${LOG2P:-26}"
    out=gpurun_out/scale_${wl}_n$n.json
    if [ $n = 1 ]; then timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 $extra > $out 2> gpurun_out/scale.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --steps 10 --warmup 3 $extra > $out 2> gpurun_out/scale.err; fi
    echo "$wl n=$n rc=$?"; grep '^{' $out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  value %.3e %s  ms_per_step %.2f  e2e %s' % (d['value'], d['unit'], d['ms_per_step'], (d['e2e'] or {}).get('value')))"
  done
done
