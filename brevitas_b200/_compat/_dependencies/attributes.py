class _Replace(Exception):
    """Raised by a factory that found out, while resolving, that its specification has to be swapped for another
    dependency (a lazily imported class): ``dependency`` is the new object, ``attrs`` the attribute path from the
    injector being resolved down to the (possibly nested) injector that holds it."""

    def __init__(self, dependency, attrs=()):
        super().__init__(dependency, attrs)
        self.dependency = dependency
        self.attrs = tuple(attrs)
