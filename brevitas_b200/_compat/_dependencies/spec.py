import inspect

from _dependencies import markers
from _dependencies.signature import names_and_defaults
from _dependencies.this import This, _make_this_spec
from _dependencies.value import Value, _make_value_spec


def _make_init_spec(dependency):
    """A class is built by calling it with its ``__init__`` arguments resolved by name."""
    init = dependency.__init__
    if init is object.__init__:
        return markers.init, dependency, [], 1
    args, have_defaults = names_and_defaults(init, dependency.__name__, skip_first=True)
    return markers.init, dependency, args, have_defaults


class _RawFactory:
    def __init__(self, dependency):
        self.dependency = dependency

    def __call__(self):
        return self.dependency


def _make_raw_spec(dependency):
    return markers.raw, _RawFactory(dependency), [], 1


class _NestedFactory:
    def __init__(self, injector):
        self.injector = injector

    def __call__(self, __self__):
        from _dependencies.injector import _with_parent
        return _with_parent(self.injector, __self__)


def _make_nested_injector_spec(dependency):
    return markers.nested_injector, _NestedFactory(dependency), ["__self__"], 0


def _make_dependency_spec(name, dependency):
    from _dependencies.injector import _InjectorType
    from _dependencies.operation import Operation, _make_operation_spec
    from _dependencies.package import Package, _make_package_spec
    if isinstance(dependency, This):
        return _make_this_spec(dependency)
    if isinstance(dependency, Value):
        return _make_value_spec(dependency)
    if isinstance(dependency, Operation):
        return _make_operation_spec(dependency)
    if isinstance(dependency, Package):
        return _make_package_spec(dependency)
    if inspect.isclass(dependency) and not name.endswith("_class"):
        if isinstance(dependency, _InjectorType):
            return _make_nested_injector_spec(dependency)
        return _make_init_spec(dependency)
    return _make_raw_spec(dependency)
