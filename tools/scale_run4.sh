#!/bin/bash
# Round-2 evidence on one 8-GPU box, kept short (box time is charged per GPU): the driver's command under torchrun at N = 8,
# then the in-process multi-GPU form (ONE process, one context owning all GPUs): parity tests and the drop-in binary /
# library call on a 2^24-power ceremony file.   usage: tools/scale_run4.sh [N=8] [outdir=gpurun_out]
N=${1:-8}
OUT=${2:-gpurun_out}
mkdir -p $OUT
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
  bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r02_scale_n$N.json 2> $OUT/r02_scale_n$N.err
echo "N=$N rc=$?"; tail -c 300 $OUT/r02_scale_n$N.json; echo
python -m pytest tests -x -q -m gpu -k "multi_gpu" > $OUT/r02_pytest_multigpu_n$N.log 2>&1; tail -2 $OUT/r02_pytest_multigpu_n$N.log
PTAU_BENCH_DIR=/dev/shm PTAU_TRACE=1 python tools/cli_bench.py 24 3 $N > $OUT/r02_cli_2pow24_gpus$N.log 2>&1
echo "cli_bench gpus=$N rc=$?"; grep -v "ptau trace" $OUT/r02_cli_2pow24_gpus$N.log
