"""
Loader of the UNMODIFIED reference (LCAV/LenslessPiCam v1.0.8) for the baseline legs of bench.py and the drop-in tests.

``baseline/_ref`` is produced by ``__graft_entry__.build()`` in the build container with

    pip install --no-index --no-build-isolation --no-deps --ignore-requires-python --target baseline/_ref <copy of /root/reference>

(git-ignored, NOT gpurun-ignored: it travels to the GPU box with the snapshot; nothing of it is in this repository's
history).  The reference imports ``matplotlib`` and ``rawpy`` at module top level (recon.py:166-167, utils/io.py:93)
without using them on this path; neither is installed in the image, so empty stub modules are put on ``sys.path``
when -- and only when -- the real ones are missing.  ``load()`` returns the ``lensless`` module or ``None``.
"""
import importlib
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_cached = None


def _stub_missing():
    missing = []
    for mod in ("matplotlib", "rawpy"):
        try:
            importlib.import_module(mod)
        except Exception:
            missing.append(mod)
    if not missing:
        return
    stub = tempfile.mkdtemp(prefix="lpc_ref_stubs_")
    for mod in missing:
        os.makedirs(os.path.join(stub, mod), exist_ok=True)
        open(os.path.join(stub, mod, "__init__.py"), "w").close()
    if "matplotlib" in missing:
        for sub in ("pyplot", "cm"):
            open(os.path.join(stub, "matplotlib", sub + ".py"), "w").close()
    sys.path.insert(0, stub)


def available():
    return os.path.isdir(os.path.join(REF_DIR, "lensless"))


def load():
    """the reference's ``lensless`` package from baseline/_ref (or None when it was not installed)"""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    _stub_missing()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        _cached = importlib.import_module("lensless")
    except Exception as e:  # pragma: no cover - reported by the caller
        print(f"baseline/_ref present but not importable: {e!r}", file=sys.stderr)
        return None
    return _cached
