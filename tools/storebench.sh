#!/bin/bash
# TMA-store variant of the per-row forward against the default kernel (sweep build); run on the GPU box
export BREVITAS_B200_LIB=brevitas_b200/libbrevitas_b200_tuning.so
mkdir -p gpurun_out
{
for shape in "bf16 4096 11008" "bf16 16384 4096" "f32 4096 11008" "f32 16384 4096" "f16 4096 11008"; do
  set -- $shape
  python tools/kbench.py --kernel fwd --dtype $1 --rows $2 --cols $3 --sweep --store
done
} > gpurun_out/storebench.log 2>&1
grep -n "BEST\|identical\|tuning=(0, 0, 0, 0, 0)\|Error\|error\|assert" gpurun_out/storebench.log
