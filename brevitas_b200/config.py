"""Process-wide flags of the host side, read from the environment once at import like ``brevitas.config``
(src/brevitas/config.py:11-21).  Every module kind (scaling, statistics, zero-point, bit-width) reads them HERE, at call
time.  Once ``brevitas_b200.install()`` has bound a Brevitas installation the values ARE the reference's
(``brevitas.config.IGNORE_MISSING_KEYS = True`` at run time, as tests/brevitas/proxy/test_weight_scaling.py:13 does,
takes effect in the fused modules too)."""
import os

_FLAGS = ("IGNORE_MISSING_KEYS", "REINIT_ON_STATE_DICT_LOAD")
_bound = None          # the reference's ``brevitas.config`` module after install()


def env_to_bool(name: str, default: bool) -> bool:
    v = os.environ.get(name)
    if v is None:
        return default
    v = v.strip().lower()
    if v in ("y", "yes", "t", "true", "on", "1"):
        return True
    if v in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError(f"invalid truth value {v!r} for {name}")


_local = {"IGNORE_MISSING_KEYS": env_to_bool('BREVITAS_IGNORE_MISSING_KEYS', False),
          "REINIT_ON_STATE_DICT_LOAD": env_to_bool('BREVITAS_REINIT_ON_STATE_DICT_LOAD', True)}


def __getattr__(name):          # module-level: only reached for names that are not real attributes (the flags)
    if name in _FLAGS:
        return getattr(_bound, name) if _bound is not None else _local[name]
    raise AttributeError(name)


def bind(reference_config):
    """follow (or, with None, stop following) the reference's config module"""
    global _bound
    _bound = reference_config
    for name in _FLAGS:          # a value assigned on this module earlier would shadow __getattr__
        globals().pop(name, None)
