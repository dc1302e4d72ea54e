import inspect

from _dependencies import markers
from _dependencies.exceptions import DependencyError
from _dependencies.signature import names_and_defaults


class Operation:
    """``@operation``: the dependency is a zero-argument callable that runs the function with injected arguments."""

    def __init__(self, function):
        if inspect.isclass(function):
            raise DependencyError("'operation' decorator can not be used on classes")
        self.__function__ = function
        self.__name__ = getattr(function, "__name__", "operation")


operation = Operation


class _Bound:
    def __init__(self, function, kwargs):
        self.__function__, self.__kwargs__ = function, kwargs

    def __call__(self):
        return self.__function__(**self.__kwargs__)


def _make_operation_spec(dependency):
    function = dependency.__function__
    args, have_defaults = names_and_defaults(function, function.__name__, skip_first=False)
    if "self" in args:
        raise DependencyError("'operation' decorator can not be used on methods")
    return markers.operation, (lambda **kwargs: _Bound(function, kwargs)), args, have_defaults
