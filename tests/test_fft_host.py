"""CPU exercise of the front end's FFT pass bodies (csrc/fft2048.cuh compiled for the host): correctness against a
float64 DFT of two packed real frames and a replay of the shared-memory bank mapping (padding is conflict free)."""
import importlib.util
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fft_pass_bodies_on_host():
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "synthetic-audio-detection_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    exe = b.build_fft_host_check()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "max_bank_conflict 1" in out.stdout


def test_radix16_fft_pass_bodies_on_host():
    """csrc/fft2048r16.cuh (the one-launch front end's 128-thread radix 16-16-8 transform): power spectra of two packed
    real frames against a float64 DFT, and every shared-memory access pattern at most 2-way conflicted (only the
    mirrored-bin read of the power stage is; the FFT passes themselves are conflict free)."""
    spec = importlib.util.spec_from_file_location("_b", os.path.join(ROOT, "synthetic-audio-detection_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    exe = b.build_fft16_host_check()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    err = float(out.stdout.split("max_rel_power_err")[1].split()[0])
    assert err < 1e-4
