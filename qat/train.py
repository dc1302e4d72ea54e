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


def bind_reference(brevitas_src=None):
    """``--frontend reference``: bind an (unmodified) Brevitas installation to the kernels.  ``brevitas_src``: a source
    tree to import it from when it is not pip-installed (``$BREVITAS_SRC``)."""
    import brevitas_b200
    brevitas_b200.install(brevitas_src or os.environ.get("BREVITAS_SRC"), fuse=True)


def build(name: str, device, collect_stats_steps: int = 300, channels_last: bool = False, frontend: str = "mirror",
          fuse_bn: bool = False):
    """``frontend``: "mirror" = this repository's re-statement of the layers (brevitas_b200.nn); "reference" = the
    reference's own brevitas.nn / brevitas_examples model code after bind_reference()."""
    spec = WORKLOADS[name]
    M = models
    if frontend == "reference":
        from qat import ref_models as M
    kw = {"fuse_bn": True} if (fuse_bn and frontend == "mirror") else {}
    if name == "tfc":
        model = M.tfc()
    elif name == "resnet18":
        model = M.resnet18(collect_stats_steps=collect_stats_steps, **kw)
    else:
        model = M.mobilenet_v1(**kw)
    model = model.to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
        for prm in model.parameters():          # [1,C,1,1] scales are not images: keep the default strides DDP buckets expect
            if prm.dim() == 4 and prm.shape[0] == 1 and prm.shape[2:] == (1, 1):
                prm.data = prm.data.reshape(-1).clone().view(prm.shape)      # default strides ([C,1,1,1], not [C,1,C,C])
    if fuse_bn and frontend == "reference":
        # the reference's model code calls batch-norm and activation itself: prepare the modules instead of the model
        import brevitas_b200
        brevitas_b200.fuse_batch_norm(model)
    if spec["loss"] == "ce":
        loss_fn = nn.CrossEntropyLoss()
    else:
        loss_fn = models.SqrHingeLoss()
    return model, loss_fn, spec


def make_optimizer(model, spec, capturable=False):
    if spec["opt"] == "adam":
        return torch.optim.Adam(model.parameters(), lr=spec["lr"], capturable=capturable)
    # graph mode: the fused multi-tensor SGD (one launch set per step instead of ~4 foreach launches per chunk)
    # (eager mode: fused=None lets torch pick its multi-tensor "foreach" path; an explicit False would select the
    # single-tensor loop -- 3 element-wise launches per parameter and step)
    return torch.optim.SGD(model.parameters(), lr=spec["lr"], momentum=0.9, weight_decay=1e-4,
                           fused=True if capturable else None)


class FlatGrads:
    """Every gradient is a view of ONE flat buffer, laid out in the order backward produces them (reverse parameter
    order) and cut into buckets.  There is no gather before the all-reduce and no copy back after it: each bucket is
    all-reduced IN PLACE (NCCL ``avg``) as soon as the last of its gradients has been accumulated, from a
    post-accumulate hook, so the exchange overlaps the rest of the backward pass -- inside a CUDA graph the hook-time
    fork / join on NCCL's stream becomes a parallel branch of the captured graph (SURVEY.md §5, §8e:
    ``gradient_as_bucket_view`` semantics, without DDP's per-step Python reducer).

    Bucket size, measured on 2 x B200 (tools/ddp_probe.sh, profiles/r02_ddp_probe.md): NCCL kernels that run BESIDE the
    backward kernels take SMs and HBM bandwidth from them, and these models' exchanges are short (0.07-0.12 ms for 17-47
    MB), so small buckets cost more than their overlap hides: ResNet-18 20.54 / 20.40 / 20.16 ms per step with 1 / 8 /
    64 MB buckets (19.95 ms without any exchange).  The default therefore keeps models of this size in ONE bucket
    (launched by the hook of the last gradient); larger models split and overlap."""

    def __init__(self, params, world=1, dist=None, bucket_bytes=64 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.world, self.dist = world, dist
        assert len({p.dtype for p in self.params}) == 1 and len({p.device for p in self.params}) == 1
        order = list(reversed(self.params))
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=order[0].dtype, device=order[0].device)
        self.buckets, self.bucket_of = [], {}          # [start, end, n_params]
        off, start, count = 0, 0, 0
        cap = max(1, bucket_bytes // self.flat.element_size())
        for p in order:
            n = p.numel()
            view = self.flat[off:off + n]
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                g = view.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)   # NHWC like p
            else:
                g = view.view(p.shape)
            p.grad = g
            self.bucket_of[p] = len(self.buckets)
            off += n
            count += 1
            if off - start >= cap:
                self.buckets.append([start, off, count])
                start, count = off, 0
        if count:
            self.buckets.append([start, off, count])
        self.pending = [b[2] for b in self.buckets]
        self.works = []
        self.reduced = [False] * len(self.buckets)
        self.allreduce_bytes = total * self.flat.element_size()
        # NCCL averages inside the collective; other backends (gloo in the CPU tests) sum, then one divide
        self.native_avg = world > 1 and dist.get_backend() == "nccl"
        if world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    def zero(self):
        self.flat.zero_()                               # one memset instead of one per parameter
        self.pending = [b[2] for b in self.buckets]
        self.reduced = [False] * len(self.buckets)
        self.works = []

    def _launch(self, i):
        s, e, _ = self.buckets[i]
        self.reduced[i] = True
        if os.environ.get("QAT_PROBE_NO_ALLREDUCE"):       # experiment knob: everything but the collective itself
            return
        op = self.dist.ReduceOp.AVG if self.native_avg else self.dist.ReduceOp.SUM
        self.works.append(self.dist.all_reduce(self.flat[s:e], op=op, async_op=True))

    def _hook(self, p):
        i = self.bucket_of[p]
        self.pending[i] -= 1
        if self.pending[i] == 0 and not self.reduced[i]:
            self._launch(i)

    def finish(self):
        """after backward: buckets whose parameters took no gradient this step, then join every exchange"""
        if self.world > 1:
            for i, done in enumerate(self.reduced):
                if not done:
                    self._launch(i)
            for w in self.works:
                w.wait()
            self.works = []
            if not self.native_avg:
                self.flat.div_(self.world)


def collecting_modules(model):
    """the scaling modules that are still in their statistics-collection phase (core/scaling/standalone.py:230-244)"""
    return [(name, m) for name, m in model.named_modules()
            if getattr(m, "collect_stats_steps", None) is not None and hasattr(m, "counter") and m.training
            and int(m.counter) <= int(m.collect_stats_steps)]


def assert_capturable(model, collecting=False):
    """A CUDA graph replays device work only: host-side state that a step is supposed to advance would freeze at its
    capture-time value.  Refuse to capture while any quantizer still counts steps on the host -- unless the
    caller captures the COLLECTION phase itself (``collecting``): between its first and its last step every collecting step
    issues the same device work (statistic, momentum update of the buffer, quantization with the batch statistic) as long as
    the running average uses a fixed momentum, and the caller advances the host counters at every replay."""
    for name, m in collecting_modules(model):
        steps = int(m.collect_stats_steps)
        if not collecting:
            raise RuntimeError(f"{name}: still collecting statistics ({int(m.counter)} of {steps} steps); run the "
                               "collection phase eagerly before capturing the step in a CUDA graph")
        if int(m.counter) < 1 or int(m.counter) >= steps or getattr(m, "momentum", None) is None:
            raise RuntimeError(f"{name}: the collection phase is capturable between its first and last step and with a fixed "
                               f"momentum only (counter {int(m.counter)} of {steps}, momentum {getattr(m, 'momentum', None)})")
    for name, m in model.named_modules():
        if getattr(m, "first_batch", False) and m.training:
            raise RuntimeError(f"{name}: running statistics not initialised yet (first_batch); run one eager step first")
        if int(getattr(m, "quant_delay_steps", 0) or 0) > 0:
            raise RuntimeError(f"{name}: quantization delay still counting down ({m.quant_delay_steps}); not capturable")


class GraphedStep:
    """The whole training step (forward, loss, backward, bucketed gradient all-reduce overlapped with backward,
    optimizer, weight clip) captured in ONE CUDA graph: small models such as TFC are launch-bound (~150 kernels of a few
    microseconds per step), and the C-ABI is capture-safe by construction (no allocation outside torch's caching
    allocator, no sync, current stream)."""

    def __init__(self, raw_model, loss_fn, opt, x, y, world=1, dist=None, bucket_bytes=64 << 20, collecting=False):
        bucket_bytes = int(float(os.environ.get("QAT_BUCKET_MB", bucket_bytes / (1 << 20))) * (1 << 20))
        self.raw, self.loss_fn, self.opt, self.world, self.dist = raw_model, loss_fn, opt, world, dist
        self.x, self.y = x.clone(), y.clone()
        assert_capturable(raw_model, collecting)
        # a graph of the collection phase: the host-side step counters are advanced here at every replay, and the graph
        # refuses to run the step that ends the phase (that one re-initialises the learned scale: run it eagerly, then
        # capture the steady-state graph)
        self.collectors = [m for _n, m in collecting_modules(raw_model)] if collecting else []
        self.grads = FlatGrads(raw_model.parameters(), world, dist, bucket_bytes)
        from brevitas_b200 import _kernels as K
        for _ in range(3):                       # warm-up on the (non-default) current stream, see run()
            self._eager_step()
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = K.launch_count
        with torch.cuda.graph(self.graph, stream=torch.cuda.current_stream()):
            self.loss = self._eager_step()
        self.captured_launches = K.launch_count - before
        for m in self.collectors:                # the capture pass ran the host code of a step but no device work
            m.counter = m.counter - 1

    def _eager_step(self):
        self.grads.zero()
        loss = self.loss_fn(self.raw(self.x), self.y)
        loss.backward()
        self.grads.finish()
        self.opt.step()
        if hasattr(self.raw, "clip_weights"):
            self.raw.clip_weights(-1, 1)
        return loss

    def __call__(self, x=None, y=None):
        if x is not None:
            self.x.copy_(x)
            self.y.copy_(y)
        for m in self.collectors:
            if int(m.counter) + 1 >= int(m.collect_stats_steps):
                raise RuntimeError("the statistics-collection phase ends with this step: run it (and the next one) eagerly, "
                                   "then capture the steady-state step")
            m.counter = m.counter + 1
        self.graph.replay()
        return self.loss


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


def run(name, batch, steps, warmup, collect_stats_steps=2, log=None, graph=False, channels_last=False, dtype="f32",
        frontend="mirror", brevitas_src=None, fuse_bn=False):
    """returns a dict with samples/s (all ranks), ms/step, kernel-launch count per step of OUR kernels"""
    import brevitas_b200  # noqa: F401
    if frontend == "reference":
        bind_reference(brevitas_src)
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
    if graph:
        # autograd's grad-accumulator nodes remember the stream they were first used on; a graph can only be captured
        # if that is not the legacy default stream, so the whole life of the model runs on a side stream
        torch.cuda.set_stream(torch.cuda.Stream(device))
    torch.manual_seed(1234)                               # identical initial weights on every rank
    raw, loss_fn, spec = build(name, device, collect_stats_steps, channels_last=channels_last, frontend=frontend,
                               fuse_bn=fuse_bn)
    tdt = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype]
    if tdt != torch.float32:
        raw = raw.to(tdt)               # parameters, buffers and activations in bf16: the packed 16-bit kernels
    model = raw
    if world > 1 and not graph:
        model = nn.parallel.DistributedDataParallel(raw, device_ids=[local], gradient_as_bucket_view=True,
                                                    broadcast_buffers=True)
    opt = make_optimizer(raw, spec, capturable=graph)
    model.train()
    batches = [make_batch(spec, batch, device, 100 + rank * 7 + i) for i in range(2)]
    if channels_last:       # NHWC end to end: cuDNN's native layout, and the fake-quant kernels take it in place
        batches = [(x.contiguous(memory_format=torch.channels_last), y) for x, y in batches]
    if tdt != torch.float32:
        batches = [(x.to(tdt), y if y.dtype == torch.int64 else y.to(tdt)) for x, y in batches]
    # warm-up runs past the statistics-collection phase of the activation quantizers (steady state)
    collecting = collect_stats_steps > 10000           # measure the statistics-collection phase itself
    for i in range(warmup if collecting else max(warmup, collect_stats_steps + 2)):
        loss = train_step(model, raw, *batches[i % 2], loss_fn, opt)
    torch.cuda.synchronize()
    if graph:
        gstep = GraphedStep(raw, loss_fn, opt, *batches[0], world=world, dist=dist, collecting=collecting)
        step_fn = lambda i: gstep(*batches[i % 2])
    else:
        step_fn = lambda i: train_step(model, raw, *batches[i % 2], loss_fn, opt)
    for i in range(10 if graph else 3):      # graph capture leaves the GPU idle for a while: replay past the clock ramp
        step_fn(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = K.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()             # ncu --profile-from-start off: only the timed steps
    e0.record()
    for i in range(steps):
        loss = step_fn(i)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1) / steps
    launches = gstep.captured_launches if graph else (K.launch_count - l0) / steps
    loss_val = float(loss.detach())
    allreduce = None
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if graph:
            # the gradient exchange on its own (all buckets back to back, nothing to overlap with): what the step would
            # pay if the exchange were exposed
            flat, g = gstep.grads.flat, gstep.grads
            dist.barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(5):
                for bs, be, _n in g.buckets:
                    dist.all_reduce(flat[bs:be], op=dist.ReduceOp.AVG)
            a1.record()
            torch.cuda.synchronize()
            ar_ms = a0.elapsed_time(a1) / 5
            allreduce = {"bytes": g.allreduce_bytes, "buckets": len(g.buckets), "ms_alone": round(ar_ms, 3),
                         "busbw_GBps": round(g.allreduce_bytes * 2 * (world - 1) / world / (ar_ms * 1e-3) / 1e9, 1),
                         "overlap": "bucket all-reduce launched from post-accumulate-grad hooks, joined before the optimizer; "
                                    "a parallel branch of the captured graph"}
    from qat import roofline as R
    counts = R.account(raw, batches[0][0], optimizer_words=7 if spec["opt"] == "adam" else 5, world=world,
                       dtype_bytes=2 if dtype == "bf16" else 4)
    roof = R.roofline(counts, ms, dtype)
    return {"model": name, "frontend": "unmodified brevitas.nn / brevitas_examples + brevitas_b200.install()"
            if frontend == "reference" else "brevitas_b200.nn mirror", "fuse_bn": bool(fuse_bn), "roofline": roof,
            "allreduce": allreduce, "per_gpu_batch": batch, "n_gpus": world, "ms_per_step": round(ms, 3),
            "samples_per_s": round(world * batch / (ms * 1e-3), 1), "fakequant_launches_per_step": launches,
            "final_loss": round(loss_val, 4), "dtype": dtype, "data": "synthetic",
            "memory_format": "channels_last" if channels_last else "contiguous",
            "phase": "collecting activation statistics (AbsPercentile every step)" if collecting
                     else f"steady state (after {collect_stats_steps} collect steps)",
            "step": "one CUDA graph (fwd+loss+bwd+bucketed all-reduce overlapped with bwd+optimizer)" if graph else "eager launches"
                    + (" + DDP bucketed NCCL all-reduce" if world > 1 else "")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet18", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--collect-stats-steps", type=int, default=2)
    ap.add_argument("--graph", action="store_true", help="capture the whole step in a CUDA graph")
    ap.add_argument("--channels-last", action="store_true", help="NHWC activations and conv weights")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--frontend", default="mirror", choices=["mirror", "reference"],
                    help="reference: the unmodified brevitas.nn / brevitas_examples models bound by brevitas_b200.install()")
    ap.add_argument("--fuse-bn", action="store_true", help="batch-norm + ReLU + activation quantizer in fused passes")
    ap.add_argument("--brevitas-src", default=None, help="source tree to import Brevitas from (default: installed / $BREVITAS_SRC)")
    a = ap.parse_args()
    os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
    os.environ.setdefault("NCCL_IB_DISABLE", "1")
    res = run(a.model, a.batch, a.steps, a.warmup, a.collect_stats_steps, graph=a.graph, channels_last=a.channels_last,
              dtype=a.dtype, frontend=a.frontend, brevitas_src=a.brevitas_src, fuse_bn=a.fuse_bn)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res))
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
