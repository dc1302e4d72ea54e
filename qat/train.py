"""Data-parallel QAT step harness: one process per GPU, weights and their fake-quant replicated,
activations / activation statistics rank-local, ONE gradient all-reduce per step (NCCL through
DistributedDataParallel, bucketed and overlapped with backward, learned scales included).

    python -m qat.train --model resnet18 --batch 256 --steps 20                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        -m qat.train --model resnet18 --batch 256 --steps 20                           # DDP, weak scaling

Prints one JSON line (rank 0): samples/s over all ranks, device-timed (CUDA events, max over ranks).
"""
import argparse
import json
import os
import sys
import time

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from qat import models  # noqa: E402

WORKLOADS = {
    # name: (factory, input shape per sample, classes, loss, optimizer)
    "tfc": dict(shape=(1, 28, 28), classes=10, loss="sqr_hinge", opt="adam", lr=0.02),
    "resnet18": dict(shape=(3, 224, 224), classes=1000, loss="ce", opt="sgd", lr=0.1),
    "mobilenet_v1": dict(shape=(3, 224, 224), classes=1000, loss="ce", opt="sgd", lr=0.05),
}


def build(name: str, device, collect_stats_steps: int = 300, channels_last: bool = False):
    spec = WORKLOADS[name]
    if name == "tfc":
        model = models.tfc()
    elif name == "resnet18":
        model = models.resnet18(collect_stats_steps=collect_stats_steps)
    else:
        model = models.mobilenet_v1()
    model = model.to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    if spec["loss"] == "ce":
        loss_fn = nn.CrossEntropyLoss()
    else:
        loss_fn = models.SqrHingeLoss()
    return model, loss_fn, spec


def make_optimizer(model, spec):
    if spec["opt"] == "adam":
        return torch.optim.Adam(model.parameters(), lr=spec["lr"])
    return torch.optim.SGD(model.parameters(), lr=spec["lr"], momentum=0.9, weight_decay=1e-4)


def make_batch(spec, batch, device, seed):
    g = torch.Generator(device=device).manual_seed(seed)
    if spec["loss"] == "ce":
        x = torch.randn(batch, *spec["shape"], device=device, generator=g)
        y = torch.randint(0, spec["classes"], (batch,), device=device, generator=g)
    else:
        x = torch.rand(batch, *spec["shape"], device=device, generator=g)               # MNIST range [0, 1]
        idx = torch.randint(0, spec["classes"], (batch,), device=device, generator=g)
        y = torch.full((batch, spec["classes"]), -1.0, device=device)
        y.scatter_(1, idx.view(-1, 1), 1.0)                                               # one-hot in {-1, +1}
    return x, y


def train_step(model, raw_model, x, y, loss_fn, opt):
    opt.zero_grad(set_to_none=True)
    out = model(x)
    loss = loss_fn(out, y)
    loss.backward()
    opt.step()
    if hasattr(raw_model, "clip_weights"):
        raw_model.clip_weights(-1, 1)                 # bnn_pynq trainer.py:246
    return loss


def run(name, batch, steps, warmup, collect_stats_steps=2, log=None):
    """returns a dict with samples/s (all ranks), ms/step, kernel-launch count per step of OUR kernels"""
    import brevitas_b200  # noqa: F401
    from brevitas_b200 import _kernels as K
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(1234)                               # identical initial weights on every rank
    raw, loss_fn, spec = build(name, device, collect_stats_steps)
    model = raw
    if world > 1:
        model = nn.parallel.DistributedDataParallel(raw, device_ids=[local], gradient_as_bucket_view=True,
                                                    broadcast_buffers=True)
    opt = make_optimizer(raw, spec)
    model.train()
    batches = [make_batch(spec, batch, device, 100 + rank * 7 + i) for i in range(2)]
    # warm-up runs past the statistics-collection phase of the activation quantizers (steady state)
    for i in range(max(warmup, collect_stats_steps + 2)):
        loss = train_step(model, raw, *batches[i % 2], loss_fn, opt)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = K.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()             # ncu --profile-from-start off: only the timed steps
    e0.record()
    for i in range(steps):
        loss = train_step(model, raw, *batches[i % 2], loss_fn, opt)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / steps
    launches = (K.launch_count - l0) / steps
    loss_val = float(loss.detach())
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"model": name, "per_gpu_batch": batch, "n_gpus": world, "ms_per_step": round(ms, 3),
            "samples_per_s": round(world * batch / (ms * 1e-3), 1), "fakequant_launches_per_step": launches,
            "final_loss": round(loss_val, 4), "dtype": "f32", "data": "synthetic",
            "phase": f"steady state (after {collect_stats_steps} collect steps)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--collect-stats-steps", type=int, default=2)
    a = ap.parse_args()
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    res = run(a.model, a.batch, a.steps, a.warmup, a.collect_stats_steps)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
