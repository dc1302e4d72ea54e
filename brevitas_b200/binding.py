"""``brevitas_b200.install()`` -- make an UNMODIFIED Brevitas installation run on ``libbrevitas_b200.so``.

Two levels, matching the two drop-in boundaries of SURVEY.md §8b:

* **op level** (always): ``brevitas.function.ops_ste.fn_prefix = torch`` so that every STE wrapper of the reference
  (src/brevitas/function/ops_ste.py:38-43, 67, 115, 142, 171, ...) dispatches to ``torch.ops.autograd_ste_ops.*``,
  the namespace of the reference's own native plugin (src/brevitas/csrc/autograd_ste_ops.cpp:258-271), which
  ``brevitas_b200.ops`` defines on the sm_100a kernels.
* **module level** (``fuse=True``): every class of ``brevitas.core.{quant,scaling,stats,zero_point,bit_width,
  restrict_val,function_wrapper,utils}`` that ``brevitas_b200.core`` mirrors (same name, same constructor argument
  names -- the injector resolves them BY NAME, src/brevitas/inject/__init__.py:98-170 -- same sub-module names, hence
  the same state-dict keys) is swapped for the mirror wherever the reference holds a reference to it: module
  globals of every loaded ``brevitas*`` module (the solvers do ``from brevitas.core.scaling import *``,
  quant/solver/common.py:5-13), class-valued attributes of the enum-like containers
  (``StatsInputViewShapeImpl.OVER_TENSOR``, core/function_wrapper/shape.py:106-111) and the ``__dependencies__``
  specifications of every already-defined injector (``NarrowIntQuant.zero_point_impl = ZeroZeroPoint``, quant/base.py:115-123).
  ``brevitas.nn`` layers, proxies, the named quantizers (``Int8WeightPerTensorFloat`` ...) and ``QuantTensor`` are
  untouched and now build ``tensor_quant`` trees whose ``forward`` launches one fused kernel.
  ``brevitas.proxy.runtime_quant.FusedActivationQuantProxy`` (:73-84) is swapped too: that is where ``nn.ReLU`` is
  folded into the quantizer kernel.

The third-party ``dependencies`` package (requirements/requirements.txt:4) is used when importable; otherwise the
clean-room stand-in under ``brevitas_b200/_compat`` is put on ``sys.path``.

``uninstall()`` undoes everything (used by the tests to run the reference's own arithmetic beside the kernels in
one process).  Nothing here falls back to the CPU: after ``install()`` the reference raises on CPU tensors.
"""
import importlib
import inspect
import os
import sys
from typing import Dict, List, Optional, Tuple

import torch

_COMPAT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_compat")

# mirror module -> reference package whose classes it re-states
_MIRRORS = [
    ("brevitas_b200.core.quant", "brevitas.core.quant"),
    ("brevitas_b200.core.scaling", "brevitas.core.scaling"),
    ("brevitas_b200.core.stats", "brevitas.core.stats"),
    ("brevitas_b200.core.zero_point", "brevitas.core.zero_point"),
    ("brevitas_b200.core.bit_width", "brevitas.core.bit_width"),
    ("brevitas_b200.core.restrict_val", "brevitas.core.restrict_val"),
    ("brevitas_b200.core.function_wrapper", "brevitas.core.function_wrapper"),
    ("brevitas_b200.core.utils", "brevitas.core.utils"),
]
_EXTRA = [  # (mirror module, class name, reference module)
    ("brevitas_b200.nn", "FusedActivationQuantProxy", "brevitas.proxy.runtime_quant"),
]

_undo: List[Tuple[object, str, object, str]] = []     # (holder, key, old value, 'attr' | 'item')
_state = {"installed": False, "fused": False, "shim": False, "table": {}}


def ensure_dependencies() -> bool:
    """Make ``import dependencies`` work; True when the stand-in had to be used."""
    try:
        import dependencies  # noqa: F401
        import _dependencies.injector  # noqa: F401
        return "brevitas_b200.compat" in getattr(dependencies, "__version__", "")
    except ImportError:
        pass
    if _COMPAT_DIR not in sys.path:
        sys.path.append(_COMPAT_DIR)
    importlib.invalidate_caches()
    import dependencies  # noqa: F401
    return True


def _preempt_native_backend():
    """With BREVITAS_JIT=1 / BREVITAS_NATIVE_STE_BACKEND=1 the reference JIT-compiles its own C++ plugin at import
    (src/brevitas/__init__.py:60-71) and would define ``autograd_ste_ops`` a second time.  The namespace already
    exists (brevitas_b200.ops), so that one ``cpp_extension.load`` call is answered without compiling; the reference
    then sets ``NATIVE_STE_BACKEND_LOADED`` and scripts its wrappers against OUR dispatcher ops."""
    from torch.utils import cpp_extension
    if getattr(cpp_extension.load, "_bvb_preempt", False):
        return
    original = cpp_extension.load

    def load(name, *args, **kwargs):
        if name == "autograd_ste_ops":
            return None
        return original(name, *args, **kwargs)

    load._bvb_preempt = True
    load._bvb_original = original
    cpp_extension.load = load


def _set(holder, key, new, kind="attr"):
    if kind == "attr":
        old = holder.__dict__.get(key) if isinstance(holder, type) else getattr(holder, key)
        _undo.append((holder, key, old, kind))
        setattr(holder, key, new)
    else:
        _undo.append((holder, key, holder[key], kind))
        holder[key] = new


def _ref_modules():
    return [m for n, m in list(sys.modules.items())
            if m is not None and (n == "brevitas" or n.startswith("brevitas.") or n == "brevitas_examples"
                                  or n.startswith("brevitas_examples."))]


def _build_table() -> Dict[type, type]:
    table = {}
    for ours_name, ref_pkg in _MIRRORS:
        ours = importlib.import_module(ours_name)
        importlib.import_module(ref_pkg)
        ref_classes = {}
        for name, mod in list(sys.modules.items()):
            if mod is not None and (name == ref_pkg or name.startswith(ref_pkg + ".")):
                for k, v in vars(mod).items():
                    if inspect.isclass(v) and getattr(v, "__module__", "").startswith(ref_pkg):
                        ref_classes[v.__name__] = v
        for k, v in vars(ours).items():
            if inspect.isclass(v) and v.__module__ == ours_name and k in ref_classes:
                table[ref_classes[k]] = v
    for ours_name, cls_name, ref_mod in _EXTRA:
        ours = getattr(importlib.import_module(ours_name), cls_name)
        ref = getattr(importlib.import_module(ref_mod), cls_name)
        if ref is not ours:
            table[ref] = ours
    return table


def _all_subclasses(cls):
    seen, todo = set(), [cls]
    while todo:
        c = todo.pop()
        for s in type.__subclasses__(c):
            if s not in seen:
                seen.add(s)
                todo.append(s)
    return seen


def _swap_everywhere(table: Dict[type, type]):
    from _dependencies.injector import _InjectorType
    from _dependencies.spec import _make_init_spec

    def mapped(v):
        try:
            return table.get(v) if isinstance(v, type) else None
        except TypeError:
            return None

    mirror_values = set(table.values())
    for mod in _ref_modules():
        for k, v in list(vars(mod).items()):
            new = mapped(v)
            if new is not None:
                _set(mod, k, new)
            elif (isinstance(v, type) and not isinstance(v, _InjectorType) and v not in mirror_values
                  and getattr(v, "__module__", "").startswith("brevitas")):
                for ck, cv in list(vars(v).items()):          # enum-like containers of classes
                    cnew = mapped(cv)
                    if cnew is not None:
                        _set(v, ck, cnew)
    import brevitas.inject as inject
    for inj in _all_subclasses(inject.ExtendedInjector):
        deps = inj.__dict__.get("__dependencies__")
        if not deps:
            continue
        for name, spec in list(deps.items()):
            new = mapped(spec[1])
            if new is not None:
                _set(deps, name, _make_init_spec(new), "item")


def _scope_parameter_init():
    """``ParameterFromStatsScalingInit.__call__`` (quant/solver/parameter.py:39-45) evaluates a statistic of the layer's
    weight while the layer is being constructed, i.e. on the host: run exactly that call inside
    ``ops.parameter_init_on_host`` (the one scope in which the ops accept host tensors)."""
    from brevitas.quant.solver import parameter as solver_parameter
    from .ops import parameter_init_on_host
    cls = solver_parameter.ParameterFromStatsScalingInit
    original = cls.__call__

    def __call__(self):
        with parameter_init_on_host():
            return original(self)

    __call__.__wrapped__ = original
    _set(cls, "__call__", __call__)


def _bind_integer_export():
    """``QuantTensor.int()`` (quant_tensor/__init__.py:174-187): after the reference's own validity check, the codes are
    written by ONE export kernel (1 read, 1 or 4 bytes written per element) instead of div, add, round and cast."""
    from brevitas.quant_tensor import QuantTensor
    original = QuantTensor.int

    def int_(self, float_datatype=False):
        value = self.value
        if (float_datatype or not isinstance(value, torch.Tensor) or not value.is_cuda or self.scale is None
                or self.zero_point is None or self.zero_point.numel() != 1
                or value.dtype not in (torch.float32, torch.bfloat16, torch.float16) or self.scale.dtype != value.dtype):
            return original(self, float_datatype)
        if not self.is_valid:
            raise RuntimeError("QuantTensor not valid.")
        narrow = bool(self.bit_width <= 8.)
        dtype = (torch.int8 if self.signed_t.item() else torch.uint8) if narrow else torch.int32
        try:
            return torch.ops.brevitas_b200.int_quant_to_int(value.detach(), self.scale.detach(), float(self.zero_point),
                                                            None, None, 0, dtype)
        except RuntimeError:                     # a scale layout the export kernel does not index
            return original(self, float_datatype)

    int_.__wrapped__ = original
    _set(QuantTensor, "int", int_)


def install(reference_path: Optional[str] = None, fuse: bool = True):
    """Bind Brevitas (already importable, or found under ``reference_path``) to the B200 kernels.  Idempotent.
    Returns the ``brevitas`` package."""
    import brevitas_b200  # noqa: F401  (loads the .so and defines torch.ops.autograd_ste_ops.* -- raises if absent)
    if reference_path is not None and reference_path not in sys.path:
        sys.path.insert(0, reference_path)
    _state["shim"] = ensure_dependencies()
    if "brevitas" not in sys.modules:
        _preempt_native_backend()
    import brevitas
    import brevitas.config as ref_config
    import brevitas.function.ops_ste as ops_ste
    from . import config
    if not _state["installed"]:
        _set(ops_ste, "fn_prefix", torch)
        _set(brevitas, "NATIVE_STE_BACKEND_LOADED", True)       # a native STE backend IS loaded: this library
        config.bind(ref_config)
        _scope_parameter_init()
        _bind_integer_export()
        _state["installed"] = True
    if fuse and not _state["fused"]:
        # everything that holds references to the core classes must be loaded before the sweep
        for name in ("brevitas.core.quant", "brevitas.core.scaling", "brevitas.core.stats", "brevitas.core.zero_point",
                     "brevitas.core.bit_width", "brevitas.core.restrict_val", "brevitas.core.function_wrapper",
                     "brevitas.quant", "brevitas.quant.solver", "brevitas.proxy", "brevitas.nn"):
            importlib.import_module(name)
        table = _build_table()
        _swap_everywhere(table)
        _state["table"] = table
        _state["fused"] = True
    return brevitas


def uninstall():
    """Restore the reference exactly as imported (Python STE backend, its own core classes)."""
    while _undo:
        holder, key, old, kind = _undo.pop()
        if kind == "attr":
            setattr(holder, key, old)
        else:
            holder[key] = old
    from . import config
    config.bind(None)
    _state.update(installed=False, fused=False, table={})


def status() -> dict:
    return {"installed": _state["installed"], "fused": _state["fused"], "dependencies_stand_in": _state["shim"],
            "classes_swapped": sorted(c.__name__ for c in _state["table"])}
