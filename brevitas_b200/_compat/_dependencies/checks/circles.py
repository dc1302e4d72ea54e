from _dependencies import markers
from _dependencies.exceptions import DependencyError


def _check_circles(dependencies):
    """No dependency may need itself, directly or through the arguments of what it needs.  Iterative three-colour
    depth-first search over the argument-name graph; nested injectors resolve in their own scope and are skipped."""
    done = set()
    for origin in dependencies:
        if origin in done:
            continue
        path, on_path = [origin], {origin}
        stack = [iter(_edges(dependencies, origin))]
        while stack:
            for name in stack[-1]:
                if name in on_path:
                    owner = dependencies[path[-1]][1]
                    raise DependencyError("{!r} is a circle dependency in the {!r} constructor".format(
                        name, getattr(owner, "__name__", path[-1])))
                if name in done or name not in dependencies:
                    continue
                path.append(name)
                on_path.add(name)
                stack.append(iter(_edges(dependencies, name)))
                break
            else:
                stack.pop()
                name = path.pop()
                on_path.discard(name)
                done.add(name)


def _edges(dependencies, name):
    spec = dependencies.get(name)
    if spec is None or spec[0] == markers.nested_injector:
        return ()
    return [a for a in spec[2] if a != "__self__"]
