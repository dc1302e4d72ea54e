from _dependencies import markers
from _dependencies.spec import _make_dependency_spec


def _deep_replace_dependency(injector, current_attr, replace):
    """Store the specification of ``replace.dependency`` under the last name of ``replace.attrs`` (default:
    ``current_attr``), descending through nested injectors for the names before it."""
    path = list(replace.attrs) or [current_attr]
    holder = injector
    for name in path[:-1]:
        marker, factory = holder.__dependencies__[name][:2]
        if marker != markers.nested_injector:
            break
        holder = factory.injector
    holder.__dependencies__[path[-1]] = _make_dependency_spec(path[-1], replace.dependency)
