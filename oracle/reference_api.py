"""Import the UNMODIFIED reference modules (TEST INFRASTRUCTURE).

Two places can hold them:
  * ``/root/reference/modular/source`` -- the read-only sources, present only in the build container
    (``make_golden.py``, CPU tests that re-check the restatement against the live reference);
  * ``oracle/_ref/*.pyc.bin`` -- the same modules byte-compiled by ``oracle/build_ref.py`` (git-ignored binaries that
    travel to the GPU box), used by ``bench.py``'s reference arm / ``cpu_baseline`` leg there.
Nothing in the product path imports this.
"""
import importlib
import importlib.util
import os
import sys

REFERENCE_SRC = "/root/reference/modular/source"
REFERENCE_BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    """The reference SOURCES are present (build container)."""
    return os.path.isfile(os.path.join(REFERENCE_SRC, "inference_runner.py"))


def compiled_available() -> bool:
    return all(os.path.isfile(os.path.join(REFERENCE_BIN, m + ".pyc.bin")) for m in ("inference_runner", "model_merger"))


def _load_compiled(name: str):
    """Execute a byte-compiled module (a .pyc: 16-byte header + marshalled code object) under its own name."""
    import marshal
    import types
    path = os.path.join(REFERENCE_BIN, name + ".pyc.bin")
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != importlib.util.MAGIC_NUMBER:
        raise RuntimeError(f"{path} was compiled by another Python version; re-run `python -m oracle.build_ref`")
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(marshal.loads(data[16:]), mod.__dict__)
    return mod


def load(allow_compiled: bool = False):
    """Return ``(inference_runner, model_merger)`` reference modules, imported under their own names with the ``timm``
    shim in place; the product's same-named modules live inside the package directory and are never on ``sys.path`` as
    top-level names, so there is no clash.  ``allow_compiled``: fall back to ``oracle/_ref/*.pyc.bin``."""
    if available():
        where = REFERENCE_SRC
    elif allow_compiled and compiled_available():
        where = REFERENCE_BIN
    else:
        raise RuntimeError("the reference is not present on this machine")
    from . import timm_shim
    timm_shim.install()
    if where == REFERENCE_BIN:
        return _load_compiled("inference_runner"), _load_compiled("model_merger")
    if where not in sys.path:
        sys.path.insert(0, where)
    ir = importlib.import_module("inference_runner")
    mm = importlib.import_module("model_merger")
    for mod in (ir, mm):
        if not os.path.abspath(mod.__file__).startswith(where):
            raise RuntimeError(f"{mod.__name__} resolved to {mod.__file__}, not the reference")
    return ir, mm
