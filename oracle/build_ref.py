"""Compile the reference's two hot-path modules into oracle/_ref/ (TEST INFRASTRUCTURE; build container only).

    python -m oracle.build_ref

The reference is pure Python; "building" it means byte-compiling
``/root/reference/modular/source/{inference_runner,model_merger}.py`` from where they lie into 
``oracle/_ref/*.pyc.bin`` (a .pyc under a neutral extension: snapshot tools commonly drop ``*.pyc``) -- the Python analogue of compiling a C reference into ``oracle/_ref/*.so``.  No reference
source enters the repository: ``oracle/_ref/`` is git-ignored (binaries only) but travels to the GPU box with the
snapshot, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg import it through ``oracle.timm_shim`` so
the CPU arm times the reference's own functions (same interpreter: the box runs this image).
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/modular/source"
OUT = os.path.join(HERE, "_ref")
MODULES = ("inference_runner", "model_merger")


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref holds both compiled modules (freshly built or already there)."""
    if not os.path.isdir(SRC):
        ok = all(os.path.exists(os.path.join(OUT, m + ".pyc.bin")) for m in MODULES)
        if verbose:
            print("oracle/_ref: reference sources absent;", "using the prebuilt files" if ok else "nothing to build")
        return ok
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(SRC, m + ".py"), cfile=os.path.join(OUT, m + ".pyc.bin"), dfile=m + ".py",
                           doraise=True, optimize=0)
    if verbose:
        print("oracle/_ref: compiled", ", ".join(m + ".pyc.bin" for m in MODULES), "with", sys.version.split()[0])
    return True


if __name__ == "__main__":
    build()
