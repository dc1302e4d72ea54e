from _dependencies.exceptions import DependencyError


def _check_inheritance(bases, injector):
    """Every base of an injector must itself be an injector (``injector``: a class or a tuple of classes)."""
    for base in bases:
        if not (isinstance(base, type) and issubclass(base, injector)):
            raise DependencyError("Multiple inheritance is allowed for Injector subclasses only")


def _check_dunder_name(name):
    if name.startswith("__") and name.endswith("__"):
        raise DependencyError("Magic methods are not allowed")


def _check_attrs_redefinition(name):
    if name == "let":
        raise DependencyError("'let' redefinition is not allowed")
