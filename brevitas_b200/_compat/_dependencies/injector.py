from _dependencies import markers
from _dependencies.attributes import _Replace
from _dependencies.checks.circles import _check_circles
from _dependencies.checks.injector import _check_attrs_redefinition, _check_dunder_name, _check_inheritance
from _dependencies.checks.loops import _check_loops
from _dependencies.exceptions import DependencyError
from _dependencies.spec import _make_dependency_spec, _make_raw_spec

_KEPT = ("__module__", "__doc__", "__weakref__", "__qualname__")


class _InjectorType(type):
    """Metaclass of ``Injector``: the class body becomes a table of specifications; reading an attribute of the
    CLASS builds that dependency, resolving constructor / function arguments by name, depth first."""

    def __new__(cls, class_name, bases, namespace):
        if not bases:
            namespace["__dependencies__"] = {}
            namespace["__wrapped__"] = None      # doctest
            namespace["_subs_tree"] = None       # typing
            return type.__new__(cls, class_name, bases, namespace)
        _check_inheritance(bases, Injector)
        ns = {}
        for attr in _KEPT:
            if attr in namespace:
                ns[attr] = namespace.pop(attr)
        for name in namespace:
            _check_dunder_name(name)
            _check_attrs_redefinition(name)
        dependencies = {}
        for base in reversed(bases):
            dependencies.update(base.__dependencies__)
        for name, dep in namespace.items():
            dependencies[name] = _make_dependency_spec(name, dep)
        _check_loops(class_name, dependencies)
        _check_circles(dependencies)
        ns["__dependencies__"] = dependencies
        return type.__new__(cls, class_name, bases, ns)

    def __getattr__(cls, attrname):
        __tracebackhide__ = True
        cache, cached = {"__self__": cls}, {"__self__"}
        current, stack, optional = attrname, [attrname], False
        while attrname not in cache:
            spec = cls.__dependencies__.get(current)
            if spec is None:
                if optional:                       # left to the default of whoever asked for it
                    cached.add(current)
                    current, optional = stack.pop(), False
                    continue
                if len(stack) > 1:
                    raise DependencyError("{!r} can not resolve attribute {!r} while building {!r}".format(
                        cls.__name__, current, stack.pop()))
                raise DependencyError("{!r} can not resolve attribute {!r}".format(cls.__name__, current))
            _, factory, args, have_defaults = spec
            missing = next(((n, a) for n, a in enumerate(args, 1) if a not in cached), None)
            if missing is not None:
                stack.append(current)
                current, optional = missing[1], missing[0] >= have_defaults
                continue
            try:
                cache[current] = factory(**{k: cache[k] for k in args if k in cache})
            except _Replace as replace:
                from _dependencies.replace import _deep_replace_dependency
                _deep_replace_dependency(cls, current, replace)
                _check_loops(cls.__name__, cls.__dependencies__)
                _check_circles(cls.__dependencies__)
                continue
            cached.add(current)
            current, optional = stack.pop(), False
        return cache[attrname]

    def __setattr__(cls, attrname, value):
        raise DependencyError("'Injector' modification is not allowed")

    def __delattr__(cls, attrname):
        raise DependencyError("'Injector' modification is not allowed")

    def __contains__(cls, attrname):
        return attrname in cls.__dependencies__

    def __and__(cls, other):
        return type(cls)(cls.__name__, (cls, other), {})

    def __dir__(cls):
        parent = set(dir(cls.__base__))
        own = set(cls.__dict__) - set(_KEPT) - {"__dependencies__"}
        return sorted((parent | own | set(cls.__dependencies__)) - {"__parent__"})


def __init__(self, *args, **kwargs):
    raise DependencyError("Do not instantiate Injector")


def let(cls, **kwargs):
    """A subclass of ``cls`` with ``kwargs`` added to (or replacing entries of) its dependencies."""
    return type(cls)(cls.__name__, (cls,), kwargs)


injector_doc = """Default dependencies specification DSL.

Classes inherited from this class may inject dependencies into classes specified in it namespace.
"""


def _with_parent(injector, parent):
    """``injector`` re-created one level below ``parent`` (so that ``this << 1`` finds it); bypasses the
    magic-name check on purpose, ``__parent__`` is not a user-visible dependency."""
    dependencies = dict(injector.__dependencies__)
    dependencies["__parent__"] = _make_raw_spec(parent)
    ns = {"__dependencies__": dependencies, "__module__": injector.__module__, "__doc__": injector.__doc__}
    return type.__new__(type(injector), injector.__name__, (injector,), ns)


Injector = _InjectorType("Injector", (), {"__init__": __init__, "__doc__": injector_doc, "let": classmethod(let)})
