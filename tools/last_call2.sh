#!/bin/bash
# The round's very last GPU call (about 100 s of box time): the GPU suite and the point-kernel timings with the library as shipped
OUT=gpurun_out; TAG=r02d; mkdir -p $OUT
timeout 88 python -m pytest tests -x -q -m gpu -v > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 20 python tools/ab_g2.py > $OUT/${TAG}_ab_default.log 2>&1; cat $OUT/${TAG}_ab_default.log
echo "elapsed ${SECONDS}s"
