"""GPU (-m gpu): BASELINE configs 1, 4 and 5 against the reference, layer by layer.

The same network is built three times from one seed and one state dict:
  (R) the UNMODIFIED reference -- brevitas_examples / brevitas.nn on its Python STE backend and ATen (fp32 IEEE);
  (F) the same reference model code after ``brevitas_b200.install()``: the injector builds fused ``tensor_quant`` trees;
  (M) this repository's mirror layers (``qat.models``), which must accept the reference's state dict unchanged.
Every quantizer proxy of the network (weight / activation / bias / truncation) is hooked; over several training-mode
steps (through the statistics-collection phase into the learned-scale phase) its quantized VALUE, SCALE, zero-point
and bit-width must be bit-identical between (R) and (F)/(M), and so must the logits and the batch-norm running
statistics.  Parameter gradients are compared with a tolerance: a quantizer's backward differs from autograd's in
summation order at the arg-max / k-th-value entries only, and that 1e-7-level difference propagates upstream.
"""
import numpy as np
import pytest
import torch

from ref_util import reference_src

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    src = reference_src()
    if src is None:
        pytest.skip("reference not available (run oracle/make_ref.py in the build container)")
    import brevitas_b200
    from brevitas_b200.binding import uninstall
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
           torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    yield src
    uninstall()
    (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic,
     torch.backends.cudnn.benchmark) = old


def hook_quantizers(model, log):
    """record (value, scale, zero_point, bit_width) of every call of every quantizer proxy"""
    handles = []
    for name, m in model.named_modules():
        leaf = name.rsplit(".", 1)[-1]
        if not leaf.endswith("_quant") or "tensor_quant" in name:
            continue

        def hook(mod, inp, out, name=name):
            val = getattr(out, "value", None)
            if val is None:
                return
            rec = [val.detach().clone()]
            for f in ("scale", "zero_point", "bit_width"):
                t = getattr(out, f, None)
                rec.append(None if t is None else t.detach().clone())
            log.append((name, rec))
        handles.append(m.register_forward_hook(hook))
    return handles


def same_bits(a, b):
    if a is None or b is None:
        return a is None and b is None
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return bool(((a.view(torch.int32) == b.view(torch.int32)) | (torch.isnan(a) & torch.isnan(b))).all())


def run_steps(model, x, steps, loss_kind, target):
    log, outs, grads = [], [], []
    handles = hook_quantizers(model, log)
    model.train()
    for step in range(steps):
        torch.manual_seed(1000 + step)                      # dropout masks
        model.zero_grad(set_to_none=True)
        out = model(x)
        if loss_kind == "ce":
            loss = torch.nn.functional.cross_entropy(out, target)
        else:
            loss = ((1.0 - out * target).clamp_min(0.0) ** 2).mean()
        loss.backward()
        outs.append(out.detach().clone())
        grads.append({n: (p.grad.detach().clone() if p.grad is not None else None) for n, p in model.named_parameters()})
    for h in handles:
        h.remove()
    return log, outs, grads


CASES = {
    # name: (input shape, steps, loss, kwargs for the two builders)
    "tfc": ((32, 1, 28, 28), 2, "hinge", {}),
    "resnet18": ((4, 3, 64, 64), 4, "ce", {"collect_stats_steps": 2}),
    "mobilenet_v1": ((2, 3, 224, 224), 2, "ce", {}),
}


@pytest.mark.parametrize("name", list(CASES))
def test_model_matches_reference_layer_by_layer(ref, name):
    import brevitas_b200
    from brevitas_b200 import _kernels
    from brevitas_b200.binding import uninstall
    shape, steps, loss_kind, kw = CASES[name]
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(shape, generator=g) if name == "tfc" else torch.randn(shape, generator=g)).cuda()
    if loss_kind == "ce":
        target = torch.randint(0, 1000, (shape[0],), generator=g).cuda()
    else:
        target = torch.full((shape[0], 10), -1.0)
        target.scatter_(1, torch.randint(0, 10, (shape[0], 1), generator=g), 1.0)
        target = target.cuda()

    # (R) the pure reference
    brevitas_b200.install(ref, fuse=False)
    uninstall()
    from qat import models, ref_models
    torch.manual_seed(0)
    model_r = getattr(ref_models, name)(**kw).cuda()
    assert "brevitas_b200" not in type(next(m for n, m in model_r.named_modules() if n.endswith("tensor_quant"))).__module__
    state = {k: v.detach().cpu().clone() for k, v in model_r.state_dict().items()}
    log_r, out_r, grad_r = run_steps(model_r, x, steps, loss_kind, target)
    bn_r = {k: v.clone() for k, v in model_r.state_dict().items() if "running_" in k}

    results = {}
    # (F) the same reference model code on the fused classes
    brevitas_b200.install(ref, fuse=True)
    torch.manual_seed(0)
    # Brevitas order of operations: load the state dict, THEN move to the device -- the proxies re-instantiate their
    # tensor_quant on every load (proxy/quant_proxy.py:124-140), on the CPU.  A checkpoint taken before the first
    # statistics-collection step holds no learned `value` yet (core/scaling/standalone.py:266-275); the reference
    # tolerates that only under brevitas.config.IGNORE_MISSING_KEYS, which the fused modules follow (brevitas_b200.config)
    import brevitas.config as ref_config
    ref_config.IGNORE_MISSING_KEYS = True
    model_f = getattr(ref_models, name)(**kw)
    assert type(next(m for n, m in model_f.named_modules() if n.endswith("tensor_quant"))).__module__.startswith("brevitas_b200")
    model_f.load_state_dict(state, strict=True)
    model_f.cuda()
    before = _kernels.launch_count
    results["reference front-end + install()"] = run_steps(model_f, x, steps, loss_kind, target) + (model_f,)
    assert _kernels.launch_count > before
    # (M) the mirror layers, same state dict
    torch.manual_seed(0)
    model_m = getattr(models, name)(**kw)
    model_m.load_state_dict(state, strict=True)
    model_m.cuda()
    results["mirror layers"] = run_steps(model_m, x, steps, loss_kind, target) + (model_m,)
    ref_config.IGNORE_MISSING_KEYS = False

    assert len(log_r) >= steps * 3
    for what, (log, outs, grads, model) in results.items():
        # per quantizer (by module path), the sequence of its calls; disabled quantizers (no scale: the reference's
        # pass-through input / output proxies) may be absent from the mirror, every ENABLED one must be there
        by_name, by_name_r = {}, {}
        for d, lg in ((by_name, log), (by_name_r, log_r)):
            for qn, rec in lg:
                if rec[1] is not None:
                    d.setdefault(qn, []).append(rec)
        if what.startswith("reference"):
            assert [n for n, _ in log] == [n for n, _ in log_r], f"{what}: quantizer call sequence differs"
        else:
            # the reference's pass-through input_quant / output_quant proxies forward the incoming QuantTensor (its scale
            # included); the mirror has no such pass-through modules.  Every real quantizer must be present.
            real = {n for n in by_name_r if n.rsplit(".", 1)[-1] in ("weight_quant", "act_quant", "bias_quant", "trunc_quant")}
            assert real <= set(by_name) <= set(by_name_r), sorted(real ^ set(by_name))[:6]
        for qn, recs in by_name.items():
            assert len(recs) == len(by_name_r[qn]), f"{what}: {qn} called {len(recs)} times, reference {len(by_name_r[qn])}"
            for rec, rec_r in zip(recs, by_name_r[qn]):
                for field, a, b in zip(("value", "scale", "zero_point", "bit_width"), rec, rec_r):
                    assert same_bits(a, b), f"{what}: {name}.{qn}.{field} differs from the reference"
        for step in range(steps):
            assert same_bits(outs[step], out_r[step]), f"{what}: logits of step {step} differ"
        for k, v in bn_r.items():
            assert same_bits(model.state_dict()[k], v), f"{what}: {k} differs"
        n_exact = n_all = 0
        for step in range(steps):
            for pn, gr in grad_r[step].items():
                gg = grads[step][pn]
                assert (gg is None) == (gr is None), f"{what}: {pn} gradient presence differs at step {step}"
                if gr is None:
                    continue
                scale = float(gr.abs().max()) + 1e-12
                assert torch.allclose(gg, gr, rtol=2e-3, atol=2e-5 * scale), \
                    f"{what}: d{pn} step {step}: max |diff| {float((gg - gr).abs().max())} of {scale}"
                n_exact += int((gg.view(torch.int32) == gr.view(torch.int32)).sum())
                n_all += gr.numel()
        print(f"{name} / {what}: {sum(len(v) for v in by_name.values())} enabled-quantizer calls bit-identical over {steps} steps; "
              f"{100.0 * n_exact / n_all:.2f}% of gradient elements bit-identical")
