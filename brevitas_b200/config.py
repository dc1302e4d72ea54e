"""Process-wide flags of the host side, read from the environment once at import like ``brevitas.config``
(src/brevitas/config.py:11-21).  After ``brevitas_b200.install()`` the values follow the reference's module
(``install`` copies ``brevitas.config.IGNORE_MISSING_KEYS`` here), so toggling either has the same effect on
every module kind (scaling, statistics, zero-point, bit-width)."""
import os


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


IGNORE_MISSING_KEYS = env_to_bool('BREVITAS_IGNORE_MISSING_KEYS', False)
REINIT_ON_STATE_DICT_LOAD = env_to_bool('BREVITAS_REINIT_ON_STATE_DICT_LOAD', True)
