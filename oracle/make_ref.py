"""TEST INFRASTRUCTURE -- recipe that places the UNMODIFIED reference (Giuseppe5/brevitas, pure Python) under
``oracle/_ref/`` so that it travels to the GPU box (``oracle/_ref/`` is git-ignored, not gpurun-ignored).

    python oracle/make_ref.py            # run where /root/reference exists (the build container); idempotent

Copies, byte for byte, ``src/brevitas`` and ``src/brevitas_examples`` (the packages) plus the part of the reference's
own test-suite that pins the hot path (``tests/brevitas/{function,core,nn,proxy}`` and their helpers).  Nothing here
is product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs import from ``oracle/_ref``.  Reference sources are never committed to this repository.
"""
import filecmp
import os
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

TREES = [
    ("src/brevitas", "src/brevitas"),
    ("src/brevitas_examples", "src/brevitas_examples"),
    ("tests/brevitas/function", "tests/brevitas/function"),
    ("tests/brevitas/core", "tests/brevitas/core"),
    ("tests/brevitas/nn", "tests/brevitas/nn"),
    ("tests/brevitas/proxy", "tests/brevitas/proxy"),
]
FILES = ["tests/__init__.py", "tests/conftest.py", "tests/brevitas/__init__.py", "tests/brevitas/common.py",
         "tests/brevitas/hyp_helper.py", "LICENSE"]
SKIP = shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.onnx", ".hypothesis")


def available() -> bool:
    return os.path.isdir(os.path.join(SRC, "src", "brevitas"))


def make(verbose: bool = True) -> str:
    if not available():
        if os.path.isdir(os.path.join(DST, "src", "brevitas")):
            return DST                                   # GPU box: use what travelled
        raise RuntimeError(f"{SRC} not present and {DST} not populated")
    for rel_src, rel_dst in TREES:
        s, d = os.path.join(SRC, rel_src), os.path.join(DST, rel_dst)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(s, d, ignore=SKIP)
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if os.path.exists(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    # the copy must be the reference, unmodified
    cmp = filecmp.dircmp(os.path.join(SRC, "src", "brevitas"), os.path.join(DST, "src", "brevitas"), ignore=["__pycache__"])
    assert not cmp.diff_files and not cmp.left_only, (cmp.diff_files, cmp.left_only)
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"oracle/_ref: {n} files copied unmodified from {SRC}")
    return DST


if __name__ == "__main__":
    make()
    sys.exit(0)
