"""Public face of the stand-in: the names Brevitas imports from ``dependencies`` (quant/base.py:5,
quant/solver/parameter.py:7, utils/jit_utils.py:14, bnn_pynq/models/common.py:5)."""
from _dependencies.exceptions import DependencyError
from _dependencies.injector import Injector
from _dependencies.operation import operation
from _dependencies.package import Package
from _dependencies.this import this
from _dependencies.value import value

__all__ = ["Injector", "Package", "DependencyError", "operation", "this", "value"]
__version__ = "2.0.1+brevitas_b200.compat"
