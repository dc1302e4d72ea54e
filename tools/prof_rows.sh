#!/bin/bash
# kernels profiled for profiles/: per-row fused fwd / bwd at the BASELINE shapes (2 launches each + 1 setup fwd)
python tools/kbench.py --kernel fwd --dtype f32 --reps 1 --warmup 1
python tools/kbench.py --kernel bwd --dtype f32 --reps 1 --warmup 1
python tools/kbench.py --kernel fwd --dtype bf16 --rows 16384 --cols 4096 --reps 1 --warmup 1
python tools/kbench.py --kernel bwd --dtype bf16 --rows 16384 --cols 4096 --masked 1 --reps 1 --warmup 1
