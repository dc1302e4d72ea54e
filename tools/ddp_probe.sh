#!/bin/bash
# N-GPU probe of the graph-mode data-parallel step: bucket sizes, and the step without the collective (rank skew only)
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 -m qat.train --model $1 --batch $2 --steps 20 --warmup 3 --graph --channels-last 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $3', d['ms_per_step'], d['samples_per_s'], d.get('allreduce'))"; }
for m in "mobilenet_v1 128" "resnet18 256"; do
  set -- $m
  QAT_PROBE_NO_ALLREDUCE=1 run $1 $2 "no-allreduce"
  QAT_BUCKET_MB=1 run $1 $2 "bucket=1MB"
  QAT_BUCKET_MB=8 run $1 $2 "bucket=8MB"
  QAT_BUCKET_MB=64 run $1 $2 "bucket=64MB"
done
python -m qat.train --model mobilenet_v1 --batch 128 --steps 20 --warmup 3 --graph --channels-last 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mobilenet N=1', d['ms_per_step'])"
