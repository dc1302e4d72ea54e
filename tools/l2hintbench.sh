#!/bin/bash
# TMA loads with an L2 evict_first hint (sweep build made with EXTRA=-DBVB_L2_EVICT_FIRST) against the product library
mkdir -p gpurun_out
{
for lib in brevitas_b200/libbrevitas_b200.so brevitas_b200/libbrevitas_b200_tuning.so; do
  echo "== $lib"
  for shape in "fwd bf16 4096 11008" "bwd bf16 4096 11008" "fwd bf16 16384 4096" "bwd bf16 16384 4096" "fwd f32 4096 11008" "bwd f32 4096 11008" "both f32 4096 11008"; do
    set -- $shape
    BREVITAS_B200_LIB=$lib python tools/kbench.py --kernel $1 --dtype $2 --rows $3 --cols $4 2>&1 | grep "us "
  done
done
} > gpurun_out/l2hint.log 2>&1
cat gpurun_out/l2hint.log
