"""CPU restatement of the reference inference hot path (TEST INFRASTRUCTURE ONLY).

Plain torch-fp32 / numpy on the CPU, written out op by op so that every step the
CUDA path fuses can be checked in isolation.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Reference = ``/root/reference/modular/source/inference_runner.py`` (IR) and
``model_merger.py`` (MM).  The arithmetic lives in un-vendored dependencies whose
published algorithm is restated here:

* torchaudio (requirements.txt:6 ``torchaudio>=0.10.0``; validated against 2.11.0):
  ``transforms.MelSpectrogram`` / ``AmplitudeToDB`` -> ``functional.spectrogram``
  (functional.py:112-145), ``melscale_fbanks`` (functional.py:518-589),
  ``amplitude_to_DB`` (functional.py:356-404).
* torchvision (requirements.txt:7 ``torchvision>=0.11.0``; validated against 0.26.0):
  ``transforms.Resize`` on a tensor == ``F.interpolate(bilinear,
  align_corners=False, antialias=True)``; the separable anti-aliased kernel of
  ATen ``UpSampleKernel.cpp`` is restated in ``resize_weights``.
* timm (requirements.txt:8 ``timm>=0.4.12``; NOT installed): ResNet-18
  ``forward_features`` = conv7x7/2+BN+ReLU, maxpool3x3/2, 8 BasicBlocks
  (torchvision/models/resnet.py:59-105, 266-276 uses the same graph and key names).

Parity pinning: the reference has no tests or golden vectors.
This restatement is pinned against outputs of the reference itself, produced in
the build container by ``oracle/make_golden.py`` and committed under
``tests/golden/`` (checked by ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Constants fixed by IR:258-259 (main() overrides the dataclass defaults)
# --------------------------------------------------------------------------------------
SAMPLE_RATE = 32000
WINDOW_SAMPLES = 128000          # int(4.0 * 32000), IR:180
N_FFT = 2048
HOP = 512
N_FREQS = N_FFT // 2 + 1         # 1025
N_FRAMES = WINDOW_SAMPLES // HOP + 1   # 251 (center=True)
N_MELS = 128
F_MIN = 20.0
F_MAX = 12000.0
TOP_DB = 80.0
AMIN = 1e-10
IMG = 512
BN_EPS = 1e-5


# --------------------------------------------------------------------------------------
# a3  slice_waveform  (IR:176-190)  -- integer indexing, must be bit exact
# --------------------------------------------------------------------------------------
def window_and_hop(sr: int, window_size: float, overlap: float) -> Tuple[int, int]:
    """IR:180-181.  Python float arithmetic, truncated with int()."""
    window = int(window_size * sr)
    hop = int((1 - overlap) * window)
    return window, hop


def slice_starts(n_samples: int, window: int, hop: int) -> List[int]:
    """IR:184: ``range(0, T - window + 1, hop)`` (the ragged tail is dropped)."""
    return list(range(0, n_samples - window + 1, hop))


def slice_waveform(wf: torch.Tensor, sr: int, window_size: float, overlap: float,
                   silence_threshold: float) -> Tuple[List[int], List[bool]]:
    """IR:176-190.  Returns (all candidate start indices, kept mask).

    A window is dropped when ``piece.abs().max() < silence_threshold`` (IR:186);
    kept windows are compacted but keep their original start for the timestamp.
    """
    window, hop = window_and_hop(sr, window_size, overlap)
    starts = slice_starts(int(wf.shape[0]), window, hop)
    kept = []
    for s in starts:
        piece = wf[s:s + window]
        kept.append(not bool(piece.abs().max() < silence_threshold))
    return starts, kept


def pad_to_window(wf: torch.Tensor, sr: int = SAMPLE_RATE, window_size: float = 4.0) -> torch.Tensor:
    """IR:150-154: zero-pad clips shorter than one window."""
    needed = int(window_size * sr)
    if wf.shape[0] < needed:
        out = torch.zeros(needed, dtype=wf.dtype)
        out[:wf.shape[0]] = wf
        return out
    return wf


# --------------------------------------------------------------------------------------
# f1  preprocess_waveform after the container is parsed  (IR:144-155)
#     torchaudio.load -> float32 in [-1,1) (int16 / 32768), mean over channels, transforms.Resample
#     (torchaudio/functional/functional.py:1452-1577: sinc_interp_hann, lowpass_filter_width 6,
#     rolloff 0.99, kernel built in float64 then cast to float32, conv1d with stride orig_freq),
#     zero-pad to one window.
# --------------------------------------------------------------------------------------
RESAMPLE_WIDTH = 6
RESAMPLE_ROLLOFF = 0.99


def resample_kernel(orig_freq: int, new_freq: int) -> Tuple[np.ndarray, int, int, int]:
    """(kernel [new, 2*width+orig] float32, width, orig, new) with orig/new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base = min(orig, new) * RESAMPLE_ROLLOFF
    width = math.ceil(RESAMPLE_WIDTH * orig / base)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    # torch.arange(0, -new, -1) is int64 and `/ new_freq` yields the DEFAULT dtype: the phase -p/new is rounded to
    # float32 before it meets the float64 index grid (exact only when new is a power of two)
    phase = (np.arange(0, -new, -1).astype(np.float32) / np.float32(new)).astype(np.float64)
    t = phase[:, None] + idx
    t = t * base
    t = np.clip(t, -RESAMPLE_WIDTH, RESAMPLE_WIDTH)
    window = np.cos(t * math.pi / RESAMPLE_WIDTH / 2) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * (window * (base / orig))
    return k.astype(np.float32), width, orig, new


def resample_length(length: int, orig: int, new: int) -> int:
    """target_length of _apply_sinc_resample_kernel: ceil evaluated on a float32 tensor (torch.as_tensor(float))."""
    return int(np.ceil(np.float32(new * length / orig)))


def resample(x: torch.Tensor, orig_freq: int, new_freq: int = SAMPLE_RATE) -> torch.Tensor:
    """y[m*new + p] = sum_k kernel[p][k] * xpad[m*orig + k], xpad = pad(x, width, width + orig); fp32."""
    if int(orig_freq) == int(new_freq):
        return x
    k, width, orig, new = resample_kernel(orig_freq, new_freq)
    n = x.shape[0]
    xp = F.pad(x.float()[None, None], (width, width + orig))
    y = F.conv1d(xp, torch.from_numpy(k)[:, None, :], stride=orig)         # [1, new, frames]
    y = y.transpose(1, 2).reshape(-1)
    return y[:resample_length(n, orig, new)]


def ingest(samples: np.ndarray, sr_in: int) -> torch.Tensor:
    """Interleaved [frames, channels] int16 (or float32) -> mono float32 at 32 kHz, zero-padded to >= one window."""
    if samples.dtype == np.int16:
        wf = torch.from_numpy(samples.astype(np.float32) / 32768.0)
    else:
        wf = torch.from_numpy(np.asarray(samples, dtype=np.float32))
    wf = wf.T.contiguous().mean(dim=0)                                      # IR:146
    wf = resample(wf, sr_in, SAMPLE_RATE)                                   # IR:147-149
    return pad_to_window(wf)                                                # IR:150-154


# --------------------------------------------------------------------------------------
# a4  waveform_to_spectrogram  (IR:157-174)
# --------------------------------------------------------------------------------------
def hann_window() -> torch.Tensor:
    """torch.hann_window(2048) periodic: 0.5 - 0.5 cos(2 pi n / N) (Spectrogram default)."""
    return torch.hann_window(N_FFT, periodic=True, dtype=torch.float32)


def hz_to_mel_htk(f: float) -> float:
    """torchaudio functional.py `_hz_to_mel`, htk branch (python float math)."""
    return 2595.0 * math.log10(1.0 + (f / 700.0))


def mel_filterbank() -> torch.Tensor:
    """[1025,128] fp32 slaney-normalised HTK triangles (torchaudio `melscale_fbanks`)."""
    all_freqs = torch.linspace(0, SAMPLE_RATE // 2, N_FREQS)
    m_pts = torch.linspace(hz_to_mel_htk(F_MIN), hz_to_mel_htk(F_MAX), N_MELS + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:N_MELS + 2] - f_pts[:N_MELS])
    return fb * enorm.unsqueeze(0)


def reflect_pad_index(j: int, n: int = WINDOW_SAMPLES, pad: int = N_FFT // 2) -> int:
    """Source sample for padded position j (torch.stft center=True, pad_mode='reflect')."""
    i = j - pad
    if i < 0:
        return -i
    if i >= n:
        return 2 * (n - 1) - i
    return i


def frames(x: torch.Tensor) -> torch.Tensor:
    """[B,128000] -> [B,251,2048] frames after reflect padding (torch.stft framing)."""
    xp = F.pad(x.unsqueeze(1), (N_FFT // 2, N_FFT // 2), mode="reflect").squeeze(1)
    return xp.unfold(-1, N_FFT, HOP)


def power_spectrogram(x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """[B,128000] -> [B,1025,251] |STFT|^2 (functional.spectrogram, power=2)."""
    fr = frames(x.to(dtype)) * hann_window().to(dtype)
    spec = torch.fft.rfft(fr, dim=-1)
    return spec.abs().pow(2.0).transpose(-1, -2)


def mel_power(x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """[B,128000] -> [B,128,251] mel power (MelScale.forward: spec^T @ fb, transposed back)."""
    p = power_spectrogram(x, dtype)
    fb = mel_filterbank().to(dtype)
    return torch.matmul(p.transpose(-1, -2), fb).transpose(-1, -2)


def amplitude_to_db(mel: torch.Tensor) -> torch.Tensor:
    """Per-segment 10*log10(clamp(x,1e-10)) then max(x, segmax-80) (amplitude_to_DB, 4-D batch rule)."""
    db = 10.0 * torch.log10(torch.clamp(mel, min=AMIN))
    floor = db.amax(dim=(-2, -1), keepdim=True) - TOP_DB
    return torch.max(db, floor)


def logmel_db(x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """[B,128000] fp32 PCM -> [B,128,251] log-mel dB (IR:158-170)."""
    return amplitude_to_db(mel_power(x, dtype))


def standardise(db: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """IR:171: (x - mean) / (std + 1e-6), UNBIASED std over all 32128 cells of a segment."""
    mu = db.mean(dim=(-2, -1), keepdim=True)
    sd = db.std(dim=(-2, -1), keepdim=True)
    return (db - mu) / (sd + 1e-6), mu.flatten(), sd.flatten()


def resize_weights(in_size: int, out_size: int = IMG) -> Tuple[np.ndarray, np.ndarray]:
    """Anti-aliased bilinear taps of ATen `_compute_indices_min_size_weights_aa`.

    Returns (xmin[out] int32, w[out,2] float32).  For up-sampling the support is 1,
    so every output sample has at most two taps; weights are normalised by their sum.
    (torchvision Resize -> F.interpolate(..., align_corners=False, antialias=True), IR:172.)
    """
    scale = np.float32(in_size) / np.float32(out_size)
    assert scale <= 1.0, "only up-sampling is used on this path"
    support = np.float32(1.0)
    xmin = np.zeros(out_size, np.int32)
    w = np.zeros((out_size, 2), np.float32)
    for i in range(out_size):
        center = np.float32(scale * np.float32(i + 0.5))
        lo = max(int(np.float32(center - support + np.float32(0.5))), 0)
        hi = min(int(np.float32(center + support + np.float32(0.5))), in_size)
        size = hi - lo
        assert 1 <= size <= 2
        ws = []
        for j in range(size):
            t = np.float32(np.float32(j + lo) - center + np.float32(0.5))
            t = abs(t)
            ws.append(np.float32(1.0) - t if t < 1.0 else np.float32(0.0))
        tot = np.float32(sum(ws))
        for j in range(size):
            w[i, j] = np.float32(ws[j] / tot)
        xmin[i] = lo
    return xmin, w


def resize_512(n: torch.Tensor) -> torch.Tensor:
    """[B,128,251] -> [B,512,512]; horizontal pass first, then vertical (ATen separable order)."""
    B, H, W = n.shape
    xw, ww = resize_weights(W)
    xh, wh = resize_weights(H)
    xw_t = torch.from_numpy(xw).long()
    xh_t = torch.from_numpy(xh).long()
    ww_t = torch.from_numpy(ww)
    wh_t = torch.from_numpy(wh)
    x1 = torch.clamp(xw_t + 1, max=W - 1)
    # ATen accumulates the taps with a fused multiply-add: fma(w1, b, w0*a)  (bit-identical here)
    tmp = torch.addcmul(n[:, :, xw_t] * ww_t[:, 0], n[:, :, x1], ww_t[:, 1])
    y1 = torch.clamp(xh_t + 1, max=H - 1)
    return torch.addcmul(tmp[:, xh_t, :] * wh_t[:, 0].unsqueeze(-1), tmp[:, y1, :], wh_t[:, 1].unsqueeze(-1))


def waveform_to_image(x: torch.Tensor) -> torch.Tensor:
    """[B,128000] -> [B,512,512] fp32 (one channel; the reference repeats it 3x, IR:173)."""
    n, _, _ = standardise(logmel_db(x))
    return resize_512(n)


# --------------------------------------------------------------------------------------
# a5/a6  BinaryClassifier / ModularMultiHeadClassifier forward from a merged state_dict
# --------------------------------------------------------------------------------------
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=BN_EPS)


def _basic_block(x, sd, p, stride):
    """torchvision/models/resnet.py:59-105 (same graph as timm BasicBlock)."""
    out = F.conv2d(x, sd[p + ".conv1.weight"], None, stride=stride, padding=1)
    out = F.relu(_bn(out, sd, p + ".bn1"))
    out = F.conv2d(out, sd[p + ".conv2.weight"], None, stride=1, padding=1)
    out = _bn(out, sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        idn = F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride=stride)
        idn = _bn(idn, sd, p + ".downsample.1")
    else:
        idn = x
    return F.relu(out + idn)


def _bottleneck_block(x, sd, p, stride):
    """torchvision/models/resnet.py:108-164 (v1.5: the stride sits on the 3x3 conv, as in timm's resnet50/101/152)."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], None), sd, p + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, sd[p + ".conv2.weight"], None, stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(F.conv2d(out, sd[p + ".conv3.weight"], None), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        idn = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride=stride), sd, p + ".downsample.1")
    else:
        idn = x
    return F.relu(out + idn)


def backbone_features(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """timm ResNet.forward_features (IR:50): [B,3,512,512] -> [B,512,16,16] for the BasicBlock nets (resnet18/34),
    [B,2048,16,16] for the Bottleneck nets (resnet50/101/152; a block with a conv3 is a Bottleneck)."""
    x = F.conv2d(x, sd[p + "conv1.weight"], None, stride=2, padding=3)
    x = F.relu(_bn(x, sd, p + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        b = 0
        while f"{p}layer{li}.{b}.conv1.weight" in sd:          # 2 blocks per layer for resnet18; 3,4,6,3 for resnet34
            blk = _bottleneck_block if f"{p}layer{li}.{b}.conv3.weight" in sd else _basic_block
            x = blk(x, sd, f"{p}layer{li}.{b}", stride if b == 0 else 1)
            b += 1
    return x


def head_mlp(feats: torch.Tensor, sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """IR:36-48 in eval mode: avgpool, flatten, Linear-BN1d-ReLU x2 (dropout = identity), Linear(256,2)."""
    v = feats.mean(dim=(2, 3))
    v = F.linear(v, sd[p + "2.weight"], sd[p + "2.bias"])
    v = F.relu(_bn(v, sd, p + "3"))
    v = F.linear(v, sd[p + "6.weight"], sd[p + "6.bias"])
    v = F.relu(_bn(v, sd, p + "7"))
    return F.linear(v, sd[p + "10.weight"], sd[p + "10.bias"])


def head_indices(sd: Dict[str, torch.Tensor]) -> List[int]:
    """IR:89-98: sorted set of <i> in keys 'sub_models.<i>.*'."""
    idx = set()
    for k in sd:
        parts = k.split(".")
        if len(parts) >= 3 and parts[0] == "sub_models":
            try:
                idx.add(int(parts[1]))
            except ValueError:
                pass
    return sorted(idx)


def per_head_logits(x: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """[B,3,512,512] -> [B,N,2] raw (Real, Synthetic) logits of each sub-model (IR:49-51)."""
    outs = []
    with torch.no_grad():
        for i in head_indices(sd):
            p = f"sub_models.{i}."
            outs.append(head_mlp(backbone_features(x, sd, p + "base."), sd, p + "head."))
    return torch.stack(outs, dim=1)


def merge_logits(per_head: torch.Tensor) -> torch.Tensor:
    """IR:62-73: [B,N,2] -> [B,N+1] = [syn_1..syn_N, mean_i(real_i)] (index 0 Real, 1 Synthetic)."""
    syn = per_head[:, :, 1]
    real_mean = per_head[:, :, 0].mean(dim=1, keepdim=True)
    return torch.cat([syn, real_mean], dim=1)


def ensemble_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    return merge_logits(per_head_logits(x, sd))


# --------------------------------------------------------------------------------------
# a8  interpret_multihead_logits (IR:194-214)  and  a9 clip aggregation (IR:328-343)
# --------------------------------------------------------------------------------------
def decide_from_probs(s: np.ndarray, threshold: float) -> int:
    """Label index: N for Real, else argmax over the N synthetic probabilities (IR:207-213)."""
    n = s.shape[0] - 1
    if s[-1] >= threshold and bool((s[:n] < threshold).all()):
        return n
    return int(np.argmax(s[:n]))


def interpret(logits: torch.Tensor, threshold: float = 0.5) -> Tuple[np.ndarray, np.ndarray]:
    """[B,N+1] logits -> (labels[B] int32 with N == Real, probs[B,N+1] fp32)."""
    s = torch.sigmoid(logits.float()).numpy()
    thr = np.float32(threshold)   # torch compares the f32 tensor against the scalar in f32
    labels = np.array([decide_from_probs(row, thr) for row in s], dtype=np.int32)
    return labels, s


def label_name(idx: int, n: int, synthetic_names: Optional[Sequence[str]], real_name: str) -> str:
    """IR:208-213 naming."""
    if idx == n:
        return real_name
    if synthetic_names and idx < len(synthetic_names):
        return synthetic_names[idx]
    return f"Synthetic_{idx + 1}"


def clip_aggregate(probs: np.ndarray, clip_ids: np.ndarray, n_clips: int, threshold: float = 0.5):
    """IR:328: per-clip mean over windows of the sigmoid outputs (x100 -> percentages).

    The reference emits no clip label; the build defines it as rule a8 applied to the
    clip-mean probabilities (same rule IR:312-323 applies to smoothed probabilities).
    Clips with no (non-silent) window get zeros and label -1.
    """
    n1 = probs.shape[1]
    out = np.zeros((n_clips, n1), np.float32)
    lab = np.full((n_clips,), -1, np.int32)
    for c in range(n_clips):
        rows = probs[clip_ids == c]
        if rows.shape[0] == 0:
            continue
        out[c] = np.mean(rows, axis=0)
        lab[c] = decide_from_probs(out[c], threshold)
    return out, lab


def smooth_probs(probs: np.ndarray, threshold: float = 0.5):
    """IR:301-325 --smooth: gaussian sigma=2 per column, renormalise rows, re-label."""
    from scipy.ndimage import gaussian_filter1d
    arr = np.array(probs, dtype=np.float32)
    for d in range(arr.shape[1]):
        arr[:, d] = gaussian_filter1d(arr[:, d], sigma=2)
    for i in range(arr.shape[0]):
        tot = arr[i].sum()
        if tot > 0:
            arr[i] /= tot
    labels = np.array([decide_from_probs(r, threshold) for r in arr], dtype=np.int32)
    return arr, labels
