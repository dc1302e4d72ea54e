"""Loader for tests/golden/*.npz (vectors produced by the real reference, see tests/golden/make_golden.py)."""
import os
from functools import lru_cache

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DTYPES = ("f32", "bf16", "f16")


@lru_cache(maxsize=None)
def load(group):
    with np.load(os.path.join(GOLDEN_DIR, f"{group}.npz")) as z:
        return {k: z[k] for k in z.files}


def case(group, prefix):
    """dict of the arrays below ``prefix`` with the prefix stripped"""
    d = load(group)
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def bits_equal(a, b):
    """bit-exact comparison of float32 arrays where all NaNs are considered equal"""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    nan = np.isnan(a) & np.isnan(b)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | nan))


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    nan = np.isnan(a) & np.isnan(b)
    bad = ~((a.view(np.uint32) == b.view(np.uint32)) | nan)
    if bad.any():
        idx = np.argwhere(bad)[:5]
        msg = ", ".join(f"{tuple(i)}: {a[tuple(i)]!r} vs {b[tuple(i)]!r}" for i in idx)
        raise AssertionError(f"{what}: {int(bad.sum())} / {a.size} elements differ bitwise; first: {msg}")


def ulp(dtype):
    return {"f32": 2.0 ** -23, "bf16": 2.0 ** -7, "f16": 2.0 ** -10}[dtype]
