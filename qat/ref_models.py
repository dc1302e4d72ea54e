"""The BASELINE.json workloads built by the REFERENCE's own, unmodified model code on top of ``brevitas.nn`` --
``brevitas_examples.bnn_pynq`` TFC (config 1), ``brevitas_examples.imagenet_classification`` MobileNetV1 4b (config 5) and
a torchvision-topology ResNet-18 out of ``brevitas.nn.QuantConv2d / QuantReLU / QuantLinear`` with their default
quantizers (config 4; the reference ships no ResNet-18) -- after ``brevitas_b200.install()`` has bound
that Brevitas installation to the kernels.  Nothing here re-states a layer or a quantizer: the injector
(brevitas.inject), proxies, ``QuantTensor`` and layer drivers that run are the reference's.

Brevitas must be importable (``pip install brevitas``, or ``brevitas_b200.install(reference_path=...)`` with a source
tree); ``qat.train --frontend reference`` takes the path from ``--brevitas-src`` / ``$BREVITAS_SRC``.
"""
import torch
from torch import nn


def _layers():
    import brevitas.nn as qnn

    class _ReferenceLayers:
        QuantConv2d, QuantReLU, QuantLinear = qnn.QuantConv2d, qnn.QuantReLU, qnn.QuantLinear
    return _ReferenceLayers


def tfc():
    """cfg/tfc_2w2a.ini through brevitas_examples/bnn_pynq/models/__init__.py:33 (FC.py:19-69)"""
    from brevitas_examples.bnn_pynq.models import model_with_cfg
    model, _ = model_with_cfg("tfc_2w2a", False)
    return model


def mobilenet_v1():
    """cfg/quant_mobilenet_v1_4b.ini (imagenet_classification/models/mobilenetv1.py:169-184)"""
    from brevitas_examples.imagenet_classification.models import model_with_cfg
    model, _ = model_with_cfg("quant_mobilenet_v1_4b", False)
    return model


def resnet18(num_classes=1000, collect_stats_steps=300):
    from qat.models import ResNet18
    return ResNet18(num_classes, act_kw={"collect_stats_steps": collect_stats_steps}, L=_layers())


def sqr_hinge_loss():
    from brevitas_examples.bnn_pynq.models.losses import SqrHingeLoss
    return SqrHingeLoss()
