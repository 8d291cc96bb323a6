#!/bin/bash
# Round-2 scaling run on one multi-GPU box: the driver's own command line at N = 1, 2, 4, 8 (headline = BASELINE
# configs[2] sharded over the ranks, with the 2^26-power configs[4] leg inside the same line), then the in-process
# multi-GPU form (one context owning all GPUs): parity tests, the drop-in binary and the library call on a
# 2^24-power ceremony file.   usage: tools/scale_run3.sh [maxN=8] [outdir=gpurun_out]
MAXN=${1:-8}
OUT=${2:-gpurun_out}
mkdir -p $OUT
for N in 1 2 4 8; do
  [ $N -gt $MAXN ] && break
  if [ $N -eq 1 ]; then
    python bench.py --gpus 1 --steps 10 --warmup 3 --no-legs > $OUT/r02_scale_n1.json 2> $OUT/r02_scale_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r02_scale_n$N.json 2> $OUT/r02_scale_n$N.err
  fi
  echo "N=$N rc=$?"; tail -c 300 $OUT/r02_scale_n$N.json | head -c 300; echo
done
python -m pytest tests -x -q -m gpu -k "multi_gpu" > $OUT/r02_pytest_multigpu.log 2>&1; tail -3 $OUT/r02_pytest_multigpu.log
for G in 1 $MAXN; do
  PTAU_BENCH_DIR=/dev/shm PTAU_TRACE=1 python tools/cli_bench.py 24 3 $G > $OUT/r02_cli_2pow24_gpus$G.log 2>&1
  echo "cli_bench gpus=$G rc=$?"; grep -v "ptau trace" $OUT/r02_cli_2pow24_gpus$G.log
done
