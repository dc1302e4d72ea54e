"""CPU: the clean-room ``dependencies`` stand-in (brevitas_b200/_compat), the UNMODIFIED reference's ``brevitas.inject``
/ ``brevitas.quant`` / ``brevitas.nn`` on top of it, and ``brevitas_b200.install()`` / ``uninstall()``.

The stand-in's contract is what src/brevitas/inject/__init__.py:7-19, 98-170 consumes; its behaviour is pinned here by
(a) direct semantics tests, (b) the reference's own nn / proxy tests (tests/brevitas/nn/test_{linear,conv2d,act}.py,
tests/brevitas/proxy/*) run on it in a subprocess, (c) the resolved ``tensor_quant`` trees of the named quantizers
(SURVEY.md Appendix B)."""
import inspect
import os
import subprocess
import sys

import pytest
import torch

from ref_util import ROOT, reference_root, reference_src

COMPAT = os.path.join(ROOT, "brevitas_b200", "_compat")
needs_ref = pytest.mark.skipif(reference_src() is None, reason="reference not available")


@pytest.fixture(scope="module")
def dep():
    import brevitas_b200.binding as inst
    inst.ensure_dependencies()
    import dependencies
    return dependencies


def test_injection_by_name_defaults_and_let(dep):
    class Amp:
        pass

    class Servo:
        def __init__(self, amplifier):
            self.amplifier = amplifier

    class Robot:
        def __init__(self, servo, controller, settings=None, speed=3, *args, **kwargs):
            self.servo, self.controller, self.settings, self.speed = servo, controller, settings, speed

    class C(dep.Injector):
        robot, servo, amplifier = Robot, Servo, Amp
        controller = "ctl"
        alias = dep.this.controller
        thing_class = Amp                      # *_class names are injected as they are

        @dep.value
        def double(speed):
            return speed * 2
        speed = 5

    r = C.robot
    assert isinstance(r.servo.amplifier, Amp) and r.controller == "ctl" and r.settings is None and r.speed == 5
    assert C.robot is not C.robot                                  # nothing is cached between accesses
    assert C.alias == "ctl" and C.double == 10 and C.thing_class is Amp
    assert "robot" in C and "nope" not in C
    D = C.let(speed=7, controller=dep.this.servo)
    assert D.robot.speed == 7 and isinstance(D.robot.controller, Servo) and C.robot.speed == 5
    assert issubclass(D, C) and D.__name__ == "C"
    assert {"robot", "servo", "let"} <= set(dir(C))
    with pytest.raises(dep.DependencyError, match="can not resolve attribute 'nothing'"):
        C.nothing
    with pytest.raises(dep.DependencyError, match="while building 'servo'"):
        C.let(amplifier=dep.this.missing_one)
        type(C)("X", (dep.Injector,), {"servo": Servo}).servo
    with pytest.raises(dep.DependencyError):
        C()
    with pytest.raises(dep.DependencyError):
        C.speed = 9


def test_definition_checks(dep):
    with pytest.raises(dep.DependencyError, match="circle link"):
        class A(dep.Injector):
            a = dep.this.b
            b = dep.this.a
    with pytest.raises(dep.DependencyError, match="circle dependency"):
        class F:
            def __init__(self, g):
                pass

        class G:
            def __init__(self, f):
                pass

        class B(dep.Injector):
            f, g = F, G
    with pytest.raises(dep.DependencyError, match="Magic methods"):
        class Cc(dep.Injector):
            __eq__ = 1
    with pytest.raises(dep.DependencyError, match="'let' redefinition"):
        class Dd(dep.Injector):
            let = 1
    with pytest.raises(dep.DependencyError, match="Injector subclasses only"):
        class Ee(dep.Injector, dict):
            pass
    with pytest.raises(dep.DependencyError):
        class Ff(dep.Injector):
            x = dep.this


def test_nested_injectors_operation_package_and(dep):
    class Outer(dep.Injector):
        x = 1

        class inner(dep.Injector):
            y = (dep.this << 1).x
            z = dep.this.y
    assert Outer.inner.y == 1 and Outer.inner.z == 1

    class Op(dep.Injector):
        @dep.operation
        def go(a, b=2):
            return a + b
        a = 1
    assert Op.go() == 3 and (Outer & Op).go() == 3 and (Outer & Op).x == 1

    class P(dep.Injector):
        od = dep.Package("collections").OrderedDict
        sep = dep.Package("os").path.sep
    assert type(P.od).__name__ == "OrderedDict" and P.sep == os.sep


@needs_ref
def test_reference_nn_and_examples_run_unmodified_on_the_stand_in():
    """In a subprocess (no brevitas_b200 op registration, pure reference on its Python backend, CPU): the reference's
    own nn / proxy tests, then the example models of BASELINE configs 1 and 5."""
    code = f"""
import sys, unittest.mock
sys.modules['mock'] = unittest.mock
sys.path[:0] = [{COMPAT!r}, {reference_src()!r}, {reference_root()!r}]
import pytest
rc = pytest.main(['-p', 'no:cacheprovider', '--rootdir={reference_root()}', '-c', '/dev/null', '-q', '-x', '--no-header',
                  'tests/brevitas/nn/test_linear.py', 'tests/brevitas/nn/test_conv2d.py', 'tests/brevitas/nn/test_act.py',
                  'tests/brevitas/proxy'])
assert rc == 0, rc
import torch
from brevitas_examples.bnn_pynq.models import model_with_cfg
m, _ = model_with_cfg('tfc_2w2a', False)
assert m(torch.rand(3, 1, 28, 28)).shape == (3, 10)
from brevitas_examples.imagenet_classification.models import model_with_cfg as im
m, _ = im('quant_mobilenet_v1_4b', False)
m.train()
out = m(torch.rand(2, 3, 224, 224)); out.sum().backward()
print('REFERENCE_OK', tuple(out.shape))
"""
    r = subprocess.run([sys.executable, "-c", code], cwd=reference_root(), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "REFERENCE_OK (2, 1000)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@needs_ref
def test_install_swaps_classes_and_uninstall_restores():
    import brevitas_b200
    from brevitas_b200.binding import status, uninstall
    uninstall()
    brevitas_b200.install(reference_src(), fuse=False)
    import brevitas
    import brevitas.core.quant as rq
    import brevitas.function.ops_ste as ops_ste
    assert ops_ste.fn_prefix is torch and rq.RescalingIntQuant.__module__ == "brevitas.core.quant.int"
    brevitas_b200.install(reference_src(), fuse=True)
    st = status()
    assert st["fused"] and len(st["classes_swapped"]) > 60
    import brevitas.nn as qnn
    import brevitas_b200.core as bc
    from brevitas.quant import (Int8ActPerTensorFloat, Int8Bias, Int8WeightPerChannelFloat, Int8WeightPerTensorFloat,
                                ShiftedUint8WeightPerTensorFloat, Uint8ActPerTensorFloat)
    from brevitas.quant.base import NarrowIntQuant
    from brevitas.core.function_wrapper.shape import StatsInputViewShapeImpl
    assert NarrowIntQuant.__dependencies__["zero_point_impl"][1] is bc.zero_point.ZeroZeroPoint
    assert StatsInputViewShapeImpl is bc.function_wrapper.StatsInputViewShapeImpl
    # SURVEY Appendix B rows 1-2, resolved by the reference's REAL injector into the mirror classes
    lin = qnn.QuantLinear(16, 8, bias=True, weight_quant=Int8WeightPerChannelFloat, bias_quant=Int8Bias)
    tq = lin.weight_quant.tensor_quant
    assert type(tq) is bc.quant.RescalingIntQuant and type(tq.int_quant) is bc.quant.IntQuant
    assert tq.int_quant.narrow_range and tq.int_quant.signed
    assert type(tq.int_quant.float_to_int_impl) is bc.function_wrapper.RoundSte
    assert type(tq.int_quant.tensor_clamp_impl) is bc.function_wrapper.TensorClampSte
    assert type(tq.scaling_impl) is bc.scaling.StatsFromParameterScaling
    assert type(tq.scaling_impl.parameter_list_stats.stats.stats_impl) is bc.stats.AbsMax
    assert tq.scaling_impl.parameter_list_stats.stats.stats_impl.stats_reduce_dim == 1
    assert type(tq.int_scaling_impl) is bc.scaling.IntScaling and type(tq.zero_point_impl) is bc.zero_point.ZeroZeroPoint
    assert tq.msb_clamp_bit_width_impl.bit_width_value == 8
    assert sorted(lin.state_dict()) == ["bias", "weight"]
    tqt = qnn.QuantLinear(16, 8, True, weight_quant=Int8WeightPerTensorFloat).weight_quant.tensor_quant
    assert tqt.scaling_impl.parameter_list_stats.stats.stats_impl.stats_reduce_dim is None
    assert type(lin.bias_quant.tensor_quant) is bc.quant.PrescaledRestrictIntQuant
    relu = qnn.QuantReLU(act_quant=Uint8ActPerTensorFloat)
    fq = relu.act_quant.fused_activation_quant_proxy
    assert type(fq).__module__ == "brevitas_b200.nn" and type(fq.activation_impl) is torch.nn.ReLU
    assert type(fq.tensor_quant.scaling_impl) is bc.scaling.ParameterFromRuntimeStatsScaling
    assert type(fq.tensor_quant.scaling_impl.stats.stats_impl) is bc.stats.AbsPercentile
    assert fq.tensor_quant.int_quant.signed is False
    assert type(fq.tensor_quant.int_quant.tensor_clamp_impl) is bc.function_wrapper.TensorClamp
    ident = qnn.QuantIdentity(act_quant=Int8ActPerTensorFloat)
    assert ident.act_quant.fused_activation_quant_proxy.tensor_quant.int_quant.signed is True
    sh = qnn.QuantLinear(16, 8, True, weight_quant=ShiftedUint8WeightPerTensorFloat).weight_quant.tensor_quant
    assert type(sh.zero_point_impl) is bc.zero_point.StatsFromParameterZeroPoint
    # no CPU fallback: the reference now raises on CPU tensors
    with pytest.raises(RuntimeError, match="CPU"):
        lin(torch.randn(2, 16))
    uninstall()
    assert ops_ste.fn_prefix is brevitas and NarrowIntQuant.__dependencies__["zero_point_impl"][1].__module__ == "brevitas.core.zero_point"
    assert type(qnn.QuantLinear(4, 4, True).weight_quant.tensor_quant).__module__ == "brevitas.core.quant.int"
    assert qnn.QuantLinear(4, 4, True)(torch.randn(2, 4)).shape == (2, 4)             # reference's own CPU path is back


@needs_ref
def test_mirror_constructors_match_the_reference_by_name():
    """The injector resolves constructor arguments BY NAME (inject/__init__.py:129-140): every mirrored class must take
    the reference's argument names, with defaults on the same ones (an argument that is optional in only one of the two
    would be resolved differently)."""
    import brevitas_b200
    from brevitas_b200.binding import _build_table, ensure_dependencies, uninstall
    uninstall()
    ensure_dependencies()
    if reference_src() not in sys.path:
        sys.path.insert(0, reference_src())
    table = _build_table()
    assert len(table) > 60
    bad = []
    for ref_cls, ours in table.items():
        def names(c):
            if c.__init__ is torch.nn.Module.__init__ or c.__init__ is object.__init__:
                return []
            ps = list(inspect.signature(c.__init__).parameters.values())[1:]
            return [(p.name, p.default is not p.empty) for p in ps if p.kind not in (p.VAR_POSITIONAL, p.VAR_KEYWORD)]
        a, b = names(ref_cls), names(ours)
        if ours.__name__ == "OverOutputChannelView":        # mirror also accepts no argument (= None)
            a = [(n, True) for n, _ in a]
        if a != b:
            bad.append((ref_cls.__name__, a, b))
    assert not bad, bad


@needs_ref
def test_config_flags_follow_the_reference_after_install():
    """one config module, read at call time by every module kind, following brevitas.config once bound"""
    import brevitas_b200
    from brevitas_b200 import config
    from brevitas_b200.binding import uninstall
    uninstall()
    assert config.IGNORE_MISSING_KEYS is False
    brevitas_b200.install(reference_src(), fuse=False)
    import brevitas.config as ref_config
    old = ref_config.IGNORE_MISSING_KEYS
    try:
        ref_config.IGNORE_MISSING_KEYS = True
        assert config.IGNORE_MISSING_KEYS is True
        from brevitas_b200.core import bit_width, scaling, stats, zero_point
        assert all(m.config is config for m in (bit_width, scaling, stats, zero_point))
    finally:
        ref_config.IGNORE_MISSING_KEYS = old
        uninstall()
    assert config.IGNORE_MISSING_KEYS is False


@needs_ref
@pytest.mark.parametrize("fuse", [False, True])
def test_construction_time_scale_init_needs_the_gpu_and_nothing_runs_on_the_cpu(fuse):
    """quant/solver/parameter.py:39-45: a learned scale initialised from the weight statistic is evaluated while the layer
    is constructed (weights still on the host).  That one call is scoped (ops.parameter_init_on_host): the host tensor is
    staged to the GPU and the sm_100a kernels compute the value -- so WITHOUT a GPU (this container) construction fails
    loudly, like every other use of the ops on host tensors.  The GPU side of this is in tests/test_gpu_named_quantizers.py
    (same values as the pure reference, bit for bit)."""
    import brevitas_b200
    from brevitas_b200.binding import uninstall
    if torch.cuda.is_available():
        pytest.skip("checks the behaviour of a machine without a GPU")
    uninstall()
    try:
        brevitas_b200.install(reference_src(), fuse=fuse)
        import brevitas.nn as qnn
        import brevitas.quant as Q
        if fuse:            # the fused statistics classes: their kernels would run on a staged copy -- no GPU here
            with pytest.raises(Exception, match="needs a CUDA device"):
                qnn.QuantLinear(16, 8, False, weight_quant=Q.Int8WeightPerChannelFloatDecoupled)
        else:               # op-level binding only: the reference's own statistics classes (ATen) build the initial value
            qnn.QuantLinear(16, 8, False, weight_quant=Q.Int8WeightPerChannelFloatDecoupled)
        with pytest.raises(RuntimeError, match="CPU tensors are not supported"):
            torch.ops.brevitas_b200.absmax_tensor(torch.randn(4))               # outside the scope: raises
        layer = qnn.QuantLinear(16, 8, False, weight_quant=Q.Int8WeightPerTensorFloat)    # nothing evaluated at construction
        with pytest.raises(RuntimeError, match="CPU tensors are not supported"):
            layer(torch.randn(2, 16))
    finally:
        uninstall()
