import importlib
import inspect

from _dependencies import markers
from _dependencies.attributes import _Replace


class Package:
    """Lazy import: ``Package('pkg').mod.Name`` is imported when the attribute is resolved; a class found that way is
    injected like a class written in the injector body (by swapping the specification, see ``_Replace``)."""

    def __init__(self, name):
        object.__setattr__(self, "__name__", name)
        object.__setattr__(self, "__attrs__", ())

    def __getattr__(self, attrname):
        if attrname.startswith("__") and attrname.endswith("__"):
            raise AttributeError(attrname)
        result = Package(self.__name__)
        object.__setattr__(result, "__attrs__", self.__attrs__ + (attrname,))
        return result


class _PackageFactory:
    def __init__(self, package):
        self.package = package

    def __call__(self):
        name, attrs = self.package.__name__, list(self.package.__attrs__)
        module = importlib.import_module(name)
        while attrs:
            try:
                module = importlib.import_module(module.__name__ + "." + attrs[0])
                attrs.pop(0)
            except ImportError:
                break
        result = module
        for attr in attrs:
            result = getattr(result, attr)
        if inspect.isclass(result):
            raise _Replace(result)
        return result


def _make_package_spec(dependency):
    return markers.package, _PackageFactory(dependency), [], 1
