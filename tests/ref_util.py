"""Locating and driving the UNMODIFIED reference (Giuseppe5/brevitas) from the tests.

``oracle/_ref`` (placed by ``oracle/make_ref.py``; travels to the GPU box) is preferred, ``/root/reference`` is the
fall-back in the build container.  ``FactoryToCuda`` lets code written for CPU tensors -- the golden generator
``tests/golden/make_golden.py`` and the reference's own test-suite -- run unchanged on the GPU: every tensor a
factory call creates from non-tensor arguments (``torch.tensor``, ``torch.randn(generator=cpu_gen)``, parameter
initialisation inside ``nn.Linear`` ...) is moved to ``cuda`` right after it was created with CPU semantics, so seeded
values are identical to the CPU run that produced the golden vectors.
"""
import os
import sys

import torch
from torch.overrides import TorchFunctionMode
from torch.utils._pytree import tree_flatten, tree_map, tree_unflatten

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for cand in (os.path.join(ROOT, "oracle", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "src", "brevitas")):
            return cand
    return None


def reference_src():
    root = reference_root()
    return None if root is None else os.path.join(root, "src")


def _has_tensor(args, kwargs):
    flat, _ = tree_flatten((args, kwargs))
    return any(isinstance(a, torch.Tensor) for a in flat)


class FactoryToCuda(TorchFunctionMode):
    def __init__(self, device="cuda"):
        super().__init__()
        self.device = torch.device(device)

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.Tensor.numpy and args and args[0].is_cuda:        # tests written for CPU call x.numpy()
            return args[0].detach().cpu().numpy()
        if _has_tensor(args, kwargs):
            # an op mixing a device tensor with a host tensor that escaped the factories (torch.from_numpy, the legacy
            # torch.Tensor(...) constructor): bring the host side over, as a test written for one device expects
            flat, spec = tree_flatten((args, kwargs))
            devs = {a.device.type for a in flat if isinstance(a, torch.Tensor)}
            if devs == {"cpu", "cuda"} and getattr(func, "__name__", "") not in ("to", "cpu", "cuda", "copy_", "numpy"):
                flat = [a.to(self.device) if isinstance(a, torch.Tensor) and a.device.type == "cpu" and a.dim() > 0 else a
                        for a in flat]
                args, kwargs = tree_unflatten(flat, spec)
            return func(*args, **kwargs)
        out = func(*args, **kwargs)
        return tree_map(lambda t: t.to(self.device) if isinstance(t, torch.Tensor) and t.device.type == "cpu" else t, out)


def import_make_golden():
    gdir = os.path.join(ROOT, "tests", "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_golden
    return make_golden
