"""Import the UNMODIFIED reference modules (build container only).

``/root/reference`` is read-only, exists only in the build container and never on
the GPU box.  Nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call
this at run time; it is used by ``make_golden.py`` and by the CPU tests that
re-check the restatement against the live reference when it is present.
"""
import importlib
import os
import sys

REFERENCE_SRC = "/root/reference/modular/source"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "inference_runner.py"))


def load():
    """Return ``(inference_runner, model_merger)`` reference modules.

    They are imported under their own names from ``REFERENCE_SRC`` with the
    ``timm`` shim in place; the product's same-named modules live inside the
    package directory and are never on ``sys.path`` as top-level names, so there
    is no clash.
    """
    if not available():
        raise RuntimeError("reference sources are not present on this machine")
    from . import timm_shim
    timm_shim.install()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    ir = importlib.import_module("inference_runner")
    mm = importlib.import_module("model_merger")
    for mod in (ir, mm):
        if not os.path.abspath(mod.__file__).startswith(REFERENCE_SRC):
            raise RuntimeError(f"{mod.__name__} resolved to {mod.__file__}, not the reference")
    return ir, mm
