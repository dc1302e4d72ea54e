class DependencyError(Exception):
    """Wrong injector definition or an attribute that cannot be resolved."""
