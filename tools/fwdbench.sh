#!/bin/bash
# forward rows kernel at the north-star shapes (default geometry), graph-free CUDA-event timing via tools/kbench.py
export BREVITAS_B200_LIB=$PWD/brevitas_b200/libbrevitas_b200_tuning.so
for cfg in "bf16 4096 11008" "bf16 16384 4096" "f32 4096 11008" "f16 4096 11008"; do set -- $cfg
  python tools/kbench.py --kernel fwd --dtype $1 --rows $2 --cols $3 --reps 50 | tail -1
done
