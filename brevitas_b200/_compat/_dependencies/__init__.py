"""Internals of the ``dependencies`` stand-in (see ../README.md).

A dependency specification is the 4-tuple the reference's metaclass unpacks
(src/brevitas/inject/__init__.py:129): ``(marker, factory, args, have_defaults)`` --

* ``marker``: a string naming the kind ("init", "value", "this", "raw", "operation", "package",
  "nested_injector"; the reference tests ``'nested' not in marker``);
* ``factory``: called with the resolved arguments as keywords, returns the dependency;
* ``args``: the names to resolve first, required ones before defaulted ones;
* ``have_defaults``: 1-based position of the first argument that has a default (len(args) + 1 if none);
  an unresolvable argument at or past that position is left to the factory's default.
"""
