from _dependencies import markers
from _dependencies.exceptions import DependencyError


def _check_loops(class_name, dependencies):
    """A chain of same-level ``this`` links (``a = this.b``, ``b = this.a``) must end somewhere."""
    for origin, spec in dependencies.items():
        if spec[0] != markers.this:
            continue
        seen, name = {origin}, _first_hop(spec)
        while name is not None:
            if name in seen:
                raise DependencyError("{!r} is a circle link in the {!r} injector".format(origin, class_name))
            seen.add(name)
            nxt = dependencies.get(name)
            name = _first_hop(nxt) if nxt is not None and nxt[0] == markers.this else None


def _first_hop(spec):
    expression = getattr(spec[1], "expression", None)
    if not expression or expression[0][0] != ".":
        return None            # starts by leaving this level (``this << n``): not a same-level link
    if len(expression) > 1:
        return None            # ``this.a.b`` reads INTO a resolved object, it cannot loop back by name
    return expression[0][1]
