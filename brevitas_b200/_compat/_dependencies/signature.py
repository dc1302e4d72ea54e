import inspect

from _dependencies.exceptions import DependencyError

_EMPTY = inspect.Parameter.empty


def names_and_defaults(func, owner_name, skip_first):
    """(argument names with the required ones first, 1-based position of the first defaulted one).

    ``*args`` / ``**kwargs`` cannot be injected by name; they are left out (the callable then sees them empty)."""
    try:
        params = list(inspect.signature(func).parameters.values())
    except (TypeError, ValueError):
        return [], 1
    if skip_first and params:
        params = params[1:]
    required, defaulted = [], []
    for p in params:
        if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        if p.kind is p.POSITIONAL_ONLY:
            raise DependencyError("{!r} has a positional-only argument {!r}".format(owner_name, p.name))
        if p.default is _EMPTY:
            required.append(p.name)
        else:
            if p.name.endswith("_class") and not inspect.isclass(p.default):
                raise DependencyError("{!r} default value should be a class".format(p.name))
            defaulted.append(p.name)
    return required + defaulted, len(required) + 1
