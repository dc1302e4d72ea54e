import inspect

from _dependencies import markers
from _dependencies.exceptions import DependencyError
from _dependencies.signature import names_and_defaults


class Value:
    """``@value``: the decorated function is called with its arguments resolved from the injector and its RESULT is
    the dependency."""

    def __init__(self, function):
        if inspect.isclass(function):
            raise DependencyError("'value' decorator can not be used on classes")
        self.__function__ = function
        self.__doc__ = getattr(function, "__doc__", None)
        self.__name__ = getattr(function, "__name__", "value")

    def __repr__(self):
        return "<value {}>".format(self.__name__)


value = Value


def _make_value_spec(dependency):
    function = dependency.__function__
    args, have_defaults = names_and_defaults(function, function.__name__, skip_first=False)
    if "self" in args:
        raise DependencyError("'value' decorator can not be used on methods")
    return markers.value, function, args, have_defaults
