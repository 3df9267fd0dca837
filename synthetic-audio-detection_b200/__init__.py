"""B200-native inference hot path of Synthetic-Audio-Detection (drop-in for modular/source/inference_runner.py
and model_merger.py of the reference).  Host code is Python/PyTorch (device memory, streams, torch.distributed);
all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI in include/sad_b200.h.

There is NO CPU fallback: importing works anywhere, but every compute entry point raises if the CUDA library is
missing or no B200-class device is present.
"""
__version__ = "0.1.0"
