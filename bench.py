#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 fake-quant hot path (BASELINE.json: "fake-quant GB/s & % HBM peak").

A "step" is one pass of the hot path over one weight: Int8 per-output-channel weight fake-quant FORWARD (abs-max
per row -> scale -> quant-dequant) + STE BACKWARD (gradient incl. the path through the abs-max) on a 4096 x 11008
fp32 linear weight (BASELINE.json configs[1], SURVEY.md §8d "C2").  Algorithmic bytes per step:
fwd 1R+1W = 8 B/elem, bwd 2R+1W = 12 B/elem -> 20 B/elem x 45,088,768 = 901.8 MB.

  value     GB/s with inputs resident in HBM (C-ABI calls, device timed with CUDA events)
  e2e       the same metric through the public module API (RescalingIntQuant + autograd) with HOST pinned
            buffers: H2D of weight and incoming gradient, D2H of the weight gradient and scales, every step
  roofline  dominant kernel (backward) algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the UNMODIFIED reference (oracle/_ref, placed by oracle/make_ref.py): brevitas.nn.QuantLinear with
            Int8WeightPerChannelFloat, quant_weight() + autograd on the host cores, the same full-size weight

`--impl reference` times only that CPU path (kind "reference"), same config, every step the full weight.
N > 1 (torchrun): the path shards trivially (independent weights) -> every rank runs its own replica of the
workload, no data-path collective ("scaling": "weak"); value = units of all ranks / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS = 4096, 11008
FWD_B, BWD_B = 8, 12          # algorithmic bytes per element, fp32
NSETS = 4                     # rotating (W, G) input sets, > L2 in total
METRIC = "int8_per_channel_weight_fakequant_fwd_bwd_GBps"


def ncu_traffic(kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_extract.py); None if absent"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            for name, rec in json.load(f).items():
                if name.startswith(kernel_prefix):
                    return int(rec["traffic_bytes"])
    except (OSError, ValueError, KeyError):
        pass
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:                         # NVML directly: ~0.1 ms per sample instead of a 100+ ms nvidia-smi process
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            R = pynvml
            while not self.stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                flag = lambda bit: "Active" if rs & bit else "Not Active"
                self.rows.append([str(sm), str(mx), str(pw), flag(R.nvmlClocksEventReasonHwSlowdown),
                                  flag(R.nvmlClocksEventReasonHwThermalSlowdown),
                                  flag(R.nvmlClocksEventReasonSwThermalSlowdown),
                                  flag(R.nvmlClocksEventReasonSwPowerCap)])
                self.stop.wait(0.002)
            return
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.rows)}


def workload_config():
    """the ``config`` object, identical in both arms (the driver compares them)"""
    n = ROWS * COLS
    return {"workload": f"C2 (BASELINE.json configs[1]): int8 per-output-channel weight fake-quant fwd + STE bwd, "
                        f"{ROWS}x{COLS} fp32, one weight per rank (Int8WeightPerChannelFloat)",
            "algorithmic_bytes_per_step": n * (FWD_B + BWD_B),
            "l2": f"inputs rotate over {NSETS} (W,G) sets of 2x{n * 4 // 2**20} MiB (> 126 MB L2)"}


def load_reference():
    """Import the UNMODIFIED reference (oracle/_ref/src, placed by oracle/make_ref.py; /root/reference/src in the build
    container) on its own Python STE backend.  None of this repository's kernels or modules are involved: only the
    clean-room stand-in for the absent third-party `dependencies` package (wiring, no arithmetic) is put on the path
    when the real one is not installed."""
    os.environ.setdefault("BREVITAS_JIT", "0")
    for cand in (os.path.join(ROOT, "oracle", "_ref", "src"), "/root/reference/src"):
        if os.path.isdir(os.path.join(cand, "brevitas")):
            if cand not in sys.path:
                sys.path.insert(0, cand)
            break
    else:
        raise RuntimeError("reference not found: run `python oracle/make_ref.py` where /root/reference exists")
    try:
        import dependencies  # noqa: F401
    except ImportError:
        sys.path.append(os.path.join(ROOT, "brevitas_b200", "_compat"))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import brevitas.nn as qnn
        from brevitas.quant import Int8WeightPerChannelFloat
    import brevitas.function.ops_ste as ops_ste
    import brevitas
    assert ops_ste.fn_prefix is brevitas, "the reference arm must run the reference's own Python STE backend"
    return qnn, Int8WeightPerChannelFloat


class ReferenceC2:
    """BASELINE config 2 through the reference's public API: ``QuantLinear(11008, 4096, weight_quant=
    Int8WeightPerChannelFloat).quant_weight()`` (nn/mixin/parameter.py:54-55 -> proxy/parameter_quant.py:83-89 ->
    core/quant/int.py:156-163) + autograd backward, full 4096 x 11008 fp32 weight, on the host cores."""

    def __init__(self, torch, rows=ROWS, device="cpu"):
        qnn, quant = load_reference()
        torch.set_num_threads(os.cpu_count() or 1)
        self.torch, self.rows = torch, rows
        gen = torch.Generator().manual_seed(0)
        self.layer = qnn.QuantLinear(COLS, rows, False, weight_quant=quant)
        with torch.no_grad():
            self.layer.weight.copy_(torch.randn(rows, COLS, generator=gen))
        self.g = torch.randn(rows, COLS, generator=gen).to(device)
        self.layer.to(device)
        tq = self.layer.weight_quant.tensor_quant
        assert type(tq).__module__ == "brevitas.core.quant.int", type(tq)

    def step(self):
        self.layer.weight.grad = None
        qw = self.layer.quant_weight()
        qw.value.backward(self.g)
        return self.layer.weight.grad

    def describe(self, t):
        return (f"full {self.rows}x{COLS} fp32 weight per step, fwd+bwd, the reference's own brevitas.nn.QuantLinear."
                f"quant_weight() + autograd on {self.torch.get_num_threads()} host threads ({t * 1e3:.0f} ms/step)")


def cpu_baseline(torch, reps):
    """the reference itself on the host cores, same workload (bounded: reps x ~0.2-1 s)"""
    ref = ReferenceC2(torch)
    ref.step()
    ref.step()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        ref.step()
        ts.append(time.perf_counter() - t0)
    t = sorted(ts)[len(ts) // 2]
    gbps = ROWS * COLS * (FWD_B + BWD_B) / t / 1e9
    return {"value": round(gbps, 3), "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": ref.describe(t) + f", median of {reps}"}, t


def cpu_tfc_baseline(torch):
    """BASELINE.json configs[0]: bnn_pynq TFC 2W2A QAT fwd+bwd on a synthetic 28x28 batch of 256, on the host cores, the
    reference's own model (brevitas_examples/bnn_pynq/models/FC.py:19-69) and loss (models/losses.py:9-31)"""
    load_reference()
    from brevitas_examples.bnn_pynq.models import model_with_cfg
    from brevitas_examples.bnn_pynq.models.losses import SqrHingeLoss
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model, _ = model_with_cfg("tfc_2w2a", False)
    model.train()
    crit = SqrHingeLoss()
    g = torch.Generator().manual_seed(0)
    x = torch.rand(256, 1, 28, 28, generator=g)
    y = torch.full((256, 10), -1.0)
    y.scatter_(1, torch.randint(0, 10, (256, 1), generator=g), 1.0)
    ts = []
    for i in range(25):
        model.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        crit(model(x), y).backward()
        ts.append(time.perf_counter() - t0)
    t = sorted(ts[5:])[10]
    return {"samples_per_s": round(256 / t, 1), "ms_per_step": round(t * 1e3, 3), "cores": torch.get_num_threads(),
            "kind": "reference", "what": "TFC 2W2A fwd+bwd (no optimizer step), batch 256, brevitas_examples.bnn_pynq "
                                         "model_with_cfg('tfc_2w2a') unmodified, Python STE backend"}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = ReferenceC2(torch)
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    dt = (time.perf_counter() - t0) / args.steps
    step_bytes = ROWS * COLS * (FWD_B + BWD_B)
    gbps = step_bytes / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(gbps, 3), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": round(gbps, 3), "unit": "GB/s", "cores": torch.get_num_threads(),
                             "kind": "reference", "sample": ref.describe(dt)},
            "e2e": {"value": round(gbps, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    args.emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the bf16 / per-token / per-tensor extra kernels")
    ap.add_argument("--no-qat", action="store_true", help="skip the QAT-step workloads")
    ap.add_argument("--qat-all", action="store_true", help="also run MobileNetV1 at N=1")
    ap.add_argument("--qat-batch", type=int, default=256, help="per-GPU batch of the ResNet-18 QAT step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: libraries (NCCL's version banner, cuDNN warnings) write to fd 1 directly,
    # so fd 1 is pointed at stderr for the duration of the run and the line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    args.emit = emit
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import brevitas_b200  # noqa: F401  (fails loudly if the native library is missing)
    from brevitas_b200 import _kernels as K
    from brevitas_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = ROWS * COLS
    step_bytes = n * (FWD_B + BWD_B)

    # ---- inputs: NSETS rotating (W, G) pairs so no step finds its inputs in the 126 MB L2 ---------------------
    gen = torch.Generator(device=dev).manual_seed(rank)
    Ws = [torch.randn(ROWS, COLS, device=dev, generator=gen) for _ in range(NSETS)]
    Gs = [torch.randn(ROWS, COLS, device=dev, generator=gen) for _ in range(NSETS)]
    args_f = (1e-10, 127.0, 0.0, -127.0, 127.0, _lib.ROUND)

    # The timed loop calls the C-ABI directly (ctypes, pre-built argument tuples, pre-allocated outputs): a Python
    # wrapper that allocates outputs per call costs more host time than these 60-110 us kernels take, which would
    # make the measurement launch-bound instead of a measurement of the kernels.
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev).cuda_stream
    Y = torch.empty(ROWS, COLS, device=dev)
    GX = torch.empty(ROWS, COLS, device=dev)
    SC = torch.empty(ROWS, device=dev)
    fwd_args = [(w.data_ptr(), Y.data_ptr(), SC.data_ptr(), None, ROWS, COLS, 1e-10, 127.0, 0.0, -127.0, 127.0,
                 _lib.ROUND, _lib.F32, stream) for w in Ws]
    bwd_args = [(g.data_ptr(), w.data_ptr(), SC.data_ptr(), None, GX.data_ptr(), ROWS, COLS, 127.0, 0.0, -127.0, 127.0,
                 _lib.ROUND, _lib.CLAMP_STE, _lib.F32, stream) for w, g in zip(Ws, Gs)]
    c_fwd, c_bwd = lib.bvb_rows_absmax_int_quant_fwd, lib.bvb_rows_absmax_int_quant_bwd

    def step(i):
        rc = c_fwd(*fwd_args[i % NSETS]) | c_bwd(*bwd_args[i % NSETS])
        if rc:
            raise RuntimeError(_lib.last_error())

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    # per-kernel events (what the roofline uses) are recorded INSIDE the timed region, but only on every EV_EVERY-th
    # step: an event record between two kernels keeps the next launch from being staged behind the running one and
    # costs 2-3 us of idle HBM each (3 per step = 5 % of a 160 us step), a measurement artefact, not workload
    EV_EVERY = 4
    ev = {i: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for i in range(0, args.steps, EV_EVERY)}
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with ClockSampler(local) as clocks:
        torch.cuda.synchronize()
        # timed region: EXACTLY K steps
        start.record()
        for i in range(args.steps):
            k = i % NSETS
            e = ev.get(i)
            if e is None:
                rc = c_fwd(*fwd_args[k]) | c_bwd(*bwd_args[k])
            else:
                e[0].record()
                rc = c_fwd(*fwd_args[k])
                e[1].record()
                rc |= c_bwd(*bwd_args[k])
                e[2].record()
            launches += 2            # rows_fwd_tma_kernel + rows_bwd_tma_kernel, one kernel per C-ABI call
            if rc:
                raise RuntimeError(_lib.last_error())
        end.record()
        torch.cuda.synchronize()
        time.sleep(0.3)              # let the sampler take a few more readings right after the burst
    total_ms = start.elapsed_time(end)
    if dist is not None:
        dist.barrier()
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev.values()) / len(ev)
    bwd_ms = sum(e[1].elapsed_time(e[2]) for e in ev.values()) / len(ev)
    value = world * step_bytes / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = peaks()
    bwd_gbps = n * BWD_B / (bwd_ms * 1e-3) / 1e9
    fwd_gbps = n * FWD_B / (fwd_ms * 1e-3) / 1e9

    # ---- e2e: public module API, host pinned buffers, H2D + D2H inside the timed region -------------------------
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import IntQuant, RescalingIntQuant
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import IntScaling, StatsFromParameterScaling
    from brevitas_b200.core.stats import AbsMax
    from brevitas_b200.core.zero_point import ZeroZeroPoint
    w_param = torch.nn.Parameter(torch.empty(ROWS, COLS, device=dev))
    tq = RescalingIntQuant(
        IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClampSte()),
        StatsFromParameterScaling(AbsMax(1), fw.OverOutputChannelView(None), 1, [w_param], FloatRestrictValue(),
                                  (ROWS, 1), False, 1e-10),
        IntScaling(True, True), ZeroZeroPoint(), BitWidthConst(8)).to(dev)
    hw = torch.randn(ROWS, COLS).pin_memory()
    hg = torch.randn(ROWS, COLS).pin_memory()
    h_gw = torch.empty(ROWS, COLS).pin_memory()
    h_scale = torch.empty(ROWS, 1).pin_memory()
    g_dev = torch.empty(ROWS, COLS, device=dev)
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_module_step():
        with torch.no_grad():
            w_param.copy_(hw, non_blocking=True)
        g_dev.copy_(hg, non_blocking=True)
        w_param.grad = None
        y, scale, zp, bw = tq(w_param)
        y.backward(g_dev)
        h_gw.copy_(w_param.grad, non_blocking=True)
        h_scale.copy_(scale.detach(), non_blocking=True)

    # the host-buffer entry point (C-ABI bvb_host_rows_fakequant_fwd_bwd): row chunks pipelined over three streams so
    # H2D, kernels and D2H overlap; same bytes moved, same results (tests/test_gpu_parity.py::test_host_pipeline_*)
    from brevitas_b200.host_pipeline import weight_fake_quant_fwd_bwd_host

    def e2e_host_step():
        weight_fake_quant_fwd_bwd_host(hw, hg, bit_width=8, signed=True, narrow_range=True, out_grad=h_gw,
                                       out_scale=h_scale, synchronize=False)

    def time_e2e(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(e2e_steps):
            fn()
        e2.record()
        torch.cuda.synchronize()
        ms = s2.elapsed_time(e2) / e2e_steps
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    e2e_module_ms = time_e2e(e2e_module_step)
    e2e_ms = time_e2e(e2e_host_step)
    e2e_val = world * step_bytes / (e2e_ms * 1e-3) / 1e9
    e2e_module_val = world * step_bytes / (e2e_module_ms * 1e-3) / 1e9
    del w_param, g_dev, tq
    torch.cuda.empty_cache()

    # ---- extras: the other hot-path kernels at BASELINE sizes (reported, not the headline) ----------------------
    extras = {}
    if not args.no_extras and rank == 0:
        def timeit(fn, reps=30, period=4, graph=True):
            """device time per launch: a CUDA graph of `period` launches (inputs rotating over > L2) replayed `reps`
            times -- the Python wrappers' allocation / marshalling costs more host time than these kernels take"""
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if not graph:
                a.record()
                for i in range(reps):
                    fn(i)
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / reps
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                keep = [fn(i) for i in range(period)]
            cg.replay()
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                cg.replay()
            b.record()
            torch.cuda.synchronize()
            del keep
            return a.elapsed_time(b) / (reps * period)

        def gb(nbytes, ms):
            return round(nbytes / (ms * 1e-3) / 1e9, 1)

        del Gs[2:]
        Wb = [w.to(torch.bfloat16) for w in Ws]
        Gb = [g.to(torch.bfloat16) for g in Gs]
        sb = K.rows_absmax_int_quant_fwd(Wb[0], ROWS, COLS, *args_f)[1]
        ms = timeit(lambda i: K.rows_absmax_int_quant_fwd(Wb[i % NSETS], ROWS, COLS, *args_f))
        extras["c2_bf16_fwd"] = {"ms": round(ms, 4), "GBps": gb(n * 4, ms)}
        ms = timeit(lambda i: K.rows_absmax_int_quant_bwd(Gb[i % 2], Wb[i % NSETS], sb, None, ROWS, COLS, 127.0, 0.0,
                                                          -127.0, 127.0, 0, 0))
        extras["c2_bf16_bwd"] = {"ms": round(ms, 4), "GBps": gb(n * 6, ms)}
        # per-tensor weight quantizer (Int8WeightPerTensorFloat): two-phase reduction + quant pass, 2R+1W
        ms = timeit(lambda i: K.tensor_absmax_int_quant_fwd(Ws[i % NSETS], torch.float32, 1e-10, 127.0, 0.0, -127.0,
                                                            127.0, 0))
        extras["c2_f32_per_tensor_fwd"] = {"ms": round(ms, 4), "GBps_algorithmic_12B": gb(n * 12, ms)}
        # C3: per-token dynamic int8 activation quant, [8, 2048, 4096] bf16 (rows = 16384 tokens of 4096)
        T, C = 8 * 2048, 4096
        X = [torch.randn(T, C, device=dev, generator=gen).to(torch.bfloat16) for _ in range(4)]
        GX = [torch.randn(T, C, device=dev, generator=gen).to(torch.bfloat16) for _ in range(2)]
        c3 = (1e-10, 128.0, 0.0, -128.0, 127.0, 0)
        sx = K.rows_absmax_int_quant_fwd(X[0], T, C, *c3)[1]
        ms = timeit(lambda i: K.rows_absmax_int_quant_fwd(X[i % 4], T, C, *c3))
        extras["c3_bf16_per_token_fwd"] = {"ms": round(ms, 4), "GBps": gb(T * C * 4, ms)}
        ms = timeit(lambda i: K.rows_absmax_int_quant_bwd(GX[i % 2], X[i % 4], sx, None, T, C, 128.0, 0.0, -128.0,
                                                          127.0, 0, 1))
        extras["c3_bf16_per_token_bwd_masked"] = {"ms": round(ms, 4), "GBps": gb(T * C * 6, ms)}
        # provided-scale activation quant (learned per-tensor scale), bf16, fwd and masked bwd with scale gradient
        s0 = torch.tensor(0.02, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda i: K.int_quant_fwd(X[i % 4], s0, 0.0, 0.0, 255.0, 0))
        extras["act_bf16_learned_scale_fwd"] = {"ms": round(ms, 4), "GBps": gb(T * C * 4, ms)}
        ms = timeit(lambda i: K.int_quant_bwd(GX[i % 2], X[i % 4], s0, 0.0, 0.0, 255.0, 0, 1, True))
        extras["act_bf16_learned_scale_bwd"] = {"ms": round(ms, 4), "GBps": gb(T * C * 6, ms)}

    # ---- QAT step (north star: data-parallel QAT across 1/2/4/8 GPUs, NCCL gradient all-reduce) ---------------------
    qat = {}
    if not args.no_qat:
        del Ws, Gs
        torch.cuda.empty_cache()
        from qat.train import run as qat_run
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        # BASELINE.json configs[3]; NHWC end to end (cuDNN's native layout; the fake-quant kernels take it in place), ReLU
        # folded into the quantizer kernels, the whole step (fwd + loss + bwd + gradient all-reduce + SGD) as ONE CUDA graph
        qat["resnet18_int8"] = qat_run("resnet18", args.qat_batch, 20, 3, channels_last=True, graph=True)
        qat["resnet18_int8_eager_ddp"] = qat_run("resnet18", args.qat_batch, 10, 3, channels_last=True)
        # the first 300 steps of the default activation quantizers collect an exact 99.999th percentile of every
        # activation tensor (AbsPercentile, radix select): step time while collecting (SURVEY.md §8d C4 "warm-up")
        qat["resnet18_int8_collecting_stats"] = qat_run("resnet18", args.qat_batch, 6, 3, collect_stats_steps=10 ** 6,
                                                        channels_last=True)
        # the same phase as ONE CUDA graph (identical device work from its second to its last-but-one step; the host-side
        # step counters are advanced at every replay, qat/train.py::GraphedStep)
        qat["resnet18_int8_collecting_stats_cuda_graph"] = qat_run("resnet18", args.qat_batch, 10, 3,
                                                                   collect_stats_steps=10 ** 6, channels_last=True, graph=True)
        qat["tfc_2w2a"] = qat_run("tfc", 256, 30, 5)                                # configs[0] shape, on the GPU
        qat["tfc_2w2a_cuda_graph"] = qat_run("tfc", 256, 200, 5, graph=True)        # same step as one CUDA graph
        # BASELINE.json configs[4]: ~1100 small launches per step, host-bound when launched eagerly
        qat["mobilenet_v1_4b"] = qat_run("mobilenet_v1", 128, 20, 3, channels_last=True, graph=True)
        if world > 1 or args.qat_all:
            qat["mobilenet_v1_4b_eager_ddp"] = qat_run("mobilenet_v1", 128, 10, 3, channels_last=True)
        # batch-norm + ReLU + activation quantizer in fused passes (csrc/bn_act_quant.cu): 8 instead of 13 passes over an
        # activation per step.  Quantizer arithmetic unchanged; the batch statistics are summed in another order than
        # cuDNN's, so results track the unfused step to fp32 summation accuracy, not bit for bit (tests/test_gpu_fused_bn.py)
        qat["resnet18_int8_fused_bn"] = qat_run("resnet18", args.qat_batch, 20, 3, channels_last=True, graph=True, fuse_bn=True)
        qat["mobilenet_v1_4b_fused_bn"] = qat_run("mobilenet_v1", 128, 20, 3, channels_last=True, graph=True, fuse_bn=True)
        # the same three workloads built by the REFERENCE's own, unmodified model code (brevitas.nn layers, injector,
        # proxies, brevitas_examples models) after brevitas_b200.install(): the drop-in a Brevitas user gets.  The host
        # framework is imported from the copy that travelled with the repository (a user has it pip-installed).
        ref_src = os.environ.get("BREVITAS_SRC") or os.path.join(ROOT, "oracle", "_ref", "src")
        if os.path.isdir(os.path.join(ref_src, "brevitas")):
            fe = dict(frontend="reference", brevitas_src=ref_src)
            qat["resnet18_int8_reference_frontend"] = qat_run("resnet18", args.qat_batch, 20, 3, channels_last=True, graph=True, **fe)
            qat["mobilenet_v1_4b_reference_frontend"] = qat_run("mobilenet_v1", 128, 20, 3, channels_last=True, graph=True, **fe)
            # ... and with brevitas_b200.fuse_batch_norm(model): the modules are prepared, the model code is not touched
            qat["resnet18_int8_reference_frontend_fused_bn"] = qat_run("resnet18", args.qat_batch, 20, 3, channels_last=True,
                                                                       graph=True, fuse_bn=True, **fe)
            qat["mobilenet_v1_4b_reference_frontend_fused_bn"] = qat_run("mobilenet_v1", 128, 20, 3, channels_last=True,
                                                                         graph=True, fuse_bn=True, **fe)
            # the reference's FC.forward builds a tensor from a Python list every call (FC.py:66: a host-to-device copy),
            # so its step cannot be captured in a CUDA graph: launched eagerly
            qat["tfc_2w2a_reference_frontend"] = qat_run("tfc", 256, 30, 5, **fe)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    from brevitas_b200.binding import uninstall
    uninstall()              # the CPU baselines run the pure reference (its own Python backend and core classes)
    cpu, _ = cpu_baseline(torch, 5)
    if not args.no_extras:
        # the reference itself on the SAME GPU (its Python STE backend on eager ATen): what a Brevitas user gets on a
        # B200 today; a baseline leg like cpu_baseline, device-timed
        gref = ReferenceC2(torch, device=dev)
        for _ in range(3):
            gref.step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gref.step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        extras["c2_f32_reference_eager_aten_same_gpu_fwd_bwd"] = {
            "ms": round(ms, 3), "GBps_algorithmic": round(step_bytes / (ms * 1e-3) / 1e9, 1),
            "what": "unmodified brevitas.nn.QuantLinear.quant_weight() + backward on cuda (Python STE backend, ATen kernels)"}
        del gref
    if not args.no_qat:
        qat["tfc_2w2a_cpu_port"] = cpu_tfc_baseline(torch)
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "hbm": {"percent_of_hbm_peak": round(100 * value / world / peak, 1), "hbm_peak_gbs": peak,
                "hbm_peak_source": peak_src, "percent_of_nominal_8TBps": round(100 * value / world / 8000, 1)},
        "roofline": {"bound": "hbm", "kernel": "rows_bwd_tma_kernel<float,ROUND|ZP0|STE> (STE backward + grad through abs-max)",
                     "achieved": round(bwd_gbps, 1), "peak": peak, "unit": "GB/s", "frac": round(bwd_gbps / peak, 4),
                     "traffic": ncu_traffic("rows_bwd_tma_kernel<float"),
                     "traffic_note": "ncu --set full, one isolated launch (profiles/): below the algorithmic bytes because "
                                     "part of the output is still dirty in L2 when the launch ends",
                     "algorithmic_bytes_per_launch": n * BWD_B, "avg_launch_ms": round(bwd_ms, 5),
                     "launches_timed": f"{len(ev)} of {args.steps} (CUDA events around every {EV_EVERY}th step's kernels, "
                                       "inside the timed region)",
                     "peak_source": peak_src,
                     "other_kernels": {"rows_fwd_tma_kernel<float,ROUND>": {
                         "achieved": round(fwd_gbps, 1), "frac": round(fwd_gbps / peak, 4),
                         "algorithmic_bytes_per_launch": n * FWD_B, "avg_launch_ms": round(fwd_ms, 5)}}},
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_val, 2), "unit": "GB/s", "h2d_bytes_per_step": 2 * n * 4,
                "d2h_bytes_per_step": n * 4 + ROWS * 4, "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
                "api": "brevitas_b200.host_pipeline.weight_fake_quant_fwd_bwd_host (C-ABI bvb_host_rows_fakequant_fwd_bwd): "
                       "pinned host W and G in, dW and scales out, 8 row chunks pipelined over 3 streams",
                "module_api": {"value": round(e2e_module_val, 2), "ms_per_step": round(e2e_module_ms, 3),
                               "api": "RescalingIntQuant(w) + autograd backward with whole-tensor H2D / D2H copies"}},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "qat_step": qat,
        "extras": extras,
    }
    if qat:
        # LAST key of the line (survives a truncated tail): the data-parallel QAT step at this N, compact
        def compact(r):
            ro = r["roofline"]
            out = {"samples_per_s": r["samples_per_s"], "ms_per_step": r["ms_per_step"], "per_gpu_batch": r["per_gpu_batch"],
                   "roofline_bound_ms": ro["bound_ms"], "frac_of_bound": ro["frac_of_bound"],
                   "serial_bound_ms": ro["serial_bound_ms"], "frac_of_serial_bound": ro["frac_of_serial_bound"],
                   "compute_ms": ro["compute_ms"], "hbm_ms": ro["hbm_ms"], "nvlink_ms": ro["nvlink_ms"]}
            if r.get("allreduce"):
                out["allreduce_bytes"] = r["allreduce"]["bytes"]
                out["allreduce_ms_alone"] = r["allreduce"]["ms_alone"]
            return out
        line["qat_scaling"] = {"n_gpus": world, "scaling": "weak (fixed per-GPU batch)",
                               "resnet18_int8": compact(qat["resnet18_int8"]),
                               "mobilenet_v1_4b": compact(qat["mobilenet_v1_4b"]),
                               "tfc_2w2a": compact(qat["tfc_2w2a_cuda_graph"])}
        for k in ("resnet18_int8", "mobilenet_v1_4b", "tfc_2w2a"):
            r = qat.get(k + "_reference_frontend")
            if r is not None:
                line["qat_scaling"][k]["unmodified_brevitas_nn_samples_per_s"] = r["samples_per_s"]
            r = qat.get(k + "_reference_frontend_fused_bn")
            if r is not None:
                line["qat_scaling"][k]["unmodified_brevitas_nn_fused_bn_samples_per_s"] = r["samples_per_s"]
            r = qat.get(k + "_fused_bn")
            if r is not None:
                line["qat_scaling"][k]["fused_bn_samples_per_s"] = r["samples_per_s"]
                line["qat_scaling"][k]["fused_bn_ms_per_step"] = r["ms_per_step"]
    args.emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
