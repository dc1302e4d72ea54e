from _dependencies import markers
from _dependencies.exceptions import DependencyError


class This:
    """A deferred reference into the injector being resolved: ``this.a.b``, ``this['k']``, ``(this << 1).a`` (one
    nested-injector level up).  Building expressions only records them; ``_ThisFactory`` evaluates."""

    def __init__(self, expression=()):
        object.__setattr__(self, "__expression__", tuple(expression))

    def __getattr__(self, attrname):
        if attrname.startswith("__") and attrname.endswith("__"):      # copy / pickle / inspect probes
            raise AttributeError(attrname)
        return This(self.__expression__ + ((".", attrname),))

    def __getitem__(self, item):
        return This(self.__expression__ + (("[]", item),))

    def __lshift__(self, num):
        if not isinstance(num, int) or num <= 0:
            raise ValueError("Positive integer argument is required")
        return This(self.__expression__ + (("<<", num),))

    def __setattr__(self, name, val):
        raise DependencyError("'this' expressions are read-only")

    def __repr__(self):
        return "this" + "".join("." + str(a) if k == "." else "[{!r}]".format(a) if k == "[]" else " << {}".format(a)
                                for k, a in self.__expression__)


this = This()


class _ThisFactory:
    def __init__(self, expression):
        self.expression = expression

    def __call__(self, __self__):
        result = __self__
        for kind, arg in self.expression:
            if kind == "<<":
                for _ in range(arg):
                    try:
                        result = result.__parent__
                    except DependencyError:
                        raise DependencyError("You tried to shift this more times than Injector has levels")
            elif kind == ".":
                result = getattr(result, arg)
            else:
                result = result[arg]
        return result


def _make_this_spec(dependency):
    expression = dependency.__expression__
    if not any(kind == "." for kind, _ in expression):
        raise DependencyError("You can not use 'this' directly in the 'Injector'")
    if expression and expression[-1][0] == "<<":
        raise DependencyError("You can not use 'this' directly in the 'Injector'")
    return markers.this, _ThisFactory(expression), ["__self__"], 0
