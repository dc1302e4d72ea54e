#!/bin/bash
# ncu launch lists of 2 timed QAT steps (eager launches: the same kernels the CUDA graph replays), run on the GPU box:
# TAG=r02c bash tools/step_launches.sh ; copy gpurun_out/${TAG}_*_step_launches*.csv into profiles/raw/
TAG=${TAG:-r02c}
mkdir -p gpurun_out
run() {  # name model batch extra...
  name=$1; model=$2; batch=$3; shift 3
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/${TAG}_${name}.csv python -m qat.train --model $model --batch $batch --steps 2 --warmup 3 \
      --channels-last "$@" > gpurun_out/${TAG}_${name}.log 2>&1
  python tools/summarize_launches.py gpurun_out/${TAG}_${name}.csv 12 | head -16
}
run resnet18_b256_nhwc_step_launches resnet18 256
run resnet18_b256_nhwc_step_launches_fused_bn resnet18 256 --fuse-bn
run mobilenetv1_b128_nhwc_step_launches_fused_bn mobilenet_v1 128 --fuse-bn
