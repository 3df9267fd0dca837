"""Build libsad_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with the repo).

Each .cu is compiled to its own object (in parallel, re-done only when the file or a header changed) and linked."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsad_b200.so")
LIB_F16 = os.path.join(HERE, "libsad_b200_f16.so")   # same sources with -DSAD_ACT_F16 (csrc/act.cuh)
SOURCES = ["api.cu", "conv_umma.cu", "conv_umma2.cu", "conv_rows.cu", "conv_rows2.cu", "block_rows.cu", "stem_fused.cu", "frontend.cu",
           "ingest.cu", "head.cu", "synth.cu"]
HEADERS = ["conv_umma.h", "frontend.h", "ingest.h", "ingest_taps.h", "head.h", "stem_fused.h", "ptx.cuh", "fft2048.cuh",
           "synth.h", "act.cuh", "fft2048r16.cuh", os.path.join("..", "..", "include", "sad_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC"]


def _newest_header() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)


def _stale(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Both builds: bf16 activations (libsad_b200.so, returned) and fp16 (libsad_b200_f16.so)."""
    _build_one(LIB_F16, os.path.join(HERE, "build_f16"), ["-DSAD_ACT_F16"], force, verbose)
    return _build_one(LIB, OBJ, [], force, verbose)


def _build_one(lib: str, obj_dir: str, defines, force: bool, verbose: bool) -> str:
    """Compile what is newer than its object, link if anything changed; returns the library path."""
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = _newest_header()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(obj_dir, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append([nvcc] + NVCC_FLAGS + defines + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", lib] +
        [os.path.join(obj_dir, s[:-3] + ".o") for s in SOURCES])
    return lib


def build_fft_host_check() -> str:
    """CPU test helper that exercises the FFT pass bodies (see csrc/fft_host_check.cpp)."""
    out = os.path.join(HERE, "fft_host_check.bin")
    src = os.path.join(CSRC, "fft_host_check.cpp")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out) or \
            os.path.getmtime(os.path.join(CSRC, "fft2048.cuh")) > os.path.getmtime(out):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, src], check=True)
    return out


def build_fft16_host_check() -> str:
    """CPU test helper for the radix 16-16-8 FFT of the one-launch front end (see csrc/fft16_host_check.cpp)."""
    out = os.path.join(HERE, "fft16_host_check.bin")
    srcs = [os.path.join(CSRC, f) for f in ("fft16_host_check.cpp", "fft2048r16.cuh", "fft2048.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(f) > os.path.getmtime(out) for f in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, srcs[0]], check=True)
    return out


def build_ingest_host_check() -> str:
    """CPU test helper for the ingest stage's host arithmetic (see csrc/ingest_host_check.cpp)."""
    out = os.path.join(HERE, "ingest_host_check.bin")
    srcs = [os.path.join(CSRC, "ingest_host_check.cpp"), os.path.join(CSRC, "ingest_taps.h")]
    if not os.path.exists(out) or any(os.path.getmtime(f) > os.path.getmtime(out) for f in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", out, srcs[0]], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
