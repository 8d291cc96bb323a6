#!/bin/bash
# tools/scale_run2.sh: on an 8-GPU box, the headline workload at N=8 and the 2^26-power setup (with the end-to-end
# leg) at N=2,4,8.  N=1 lines come from a 1-GPU box.
mkdir -p gpurun_out
run() {  # n, tag, args...
  n=$1; tag=$2; shift 2
  out=gpurun_out/scale2_${tag}_n$n.json
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n$n bench.py --gpus $n "$@" > $out 2> gpurun_out/scale2_${tag}_n$n.err
  echo "$tag n=$n rc=$?"; grep '^{' $out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('  value %.4e %s  ms_per_step %.2f  e2e %s' % (d['value'], d['unit'], d['ms_per_step'], (d['e2e'] or {}).get('value')))"
}
run 8 config2 --steps 10 --warmup 3 --no-extra
for n in 8 4 2; do run $n config5 --workload config5 --log2-powers 26; done
