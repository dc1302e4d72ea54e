#!/bin/bash
# Final evidence set of a round (run on the GPU box): TAG=r01w bash tools/capture_round.sh; copy the files from
# gpurun_out/ into profiles/ (raw csv under profiles/raw/) and run tools/make_profile_summary.py $TAG
TAG=${TAG:-r02c}
set -x
python bench.py --qat-all > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err; echo bench rc=$?
(python tools/prof_all.py; python tools/prof_misc.py) > gpurun_out/${TAG}_kernel_timings_graph_replay.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --no-qat --no-extras --steps 20 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/${TAG}_all -f python tools/prof_all.py --ncu --reps 1 > gpurun_out/${TAG}_ncu_all.log 2>&1
ncu -i gpurun_out/${TAG}_all.ncu-rep --page raw --csv > gpurun_out/${TAG}_all_kernels_full_raw.csv
ncu -i gpurun_out/${TAG}_all.ncu-rep --page details --csv > gpurun_out/${TAG}_all_kernels_full_details.csv
rm -f gpurun_out/${TAG}_all.ncu-rep
tail -c 600 gpurun_out/${TAG}_bench_plain.json
