#!/usr/bin/env python3
"""Drop-in for the reference's ``modular/source/inference_runner.py``: same names, signatures, CLI flags, JSON schema
and error behaviour; the arithmetic runs in hand-written sm_100a CUDA kernels behind include/sad_b200.h.

Reference lines cited as IR:<line> (= /root/reference/modular/source/inference_runner.py).  Differences, all forced by
the "no CPU fallback, no network" rules of this build:
  * constructing ``BinaryClassifier`` does not download ImageNet weights (IR:35 uses pretrained=True);
  * ``--device cpu`` / a machine without CUDA is an error instead of a silent CPU run (IR:243);
  * only the fixed geometry of IR:258-259 (32 kHz, 4 s, n_fft 2048, hop 512, 128 mels, 20-12000 Hz, top_db 80,
    slaney) has kernels; other SpectrogramConfig values raise NotImplementedError.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

if __package__ in (None, ""):                       # executed as a script: make the package importable as sad_b200
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sad_b200  # noqa: F401
    from sad_b200 import _lib
    from sad_b200.backbone import ResNetTrunk
    from sad_b200.engine import Engine
else:
    from . import _lib
    from .backbone import ResNetTrunk
    from .engine import Engine

SEGMENT = 128000


def _cuda_device(device) -> torch.device:
    d = torch.device(device)
    if d.type != "cuda" or not torch.cuda.is_available():
        raise _lib.SadError(f"device {d}: this build has no CPU fallback; a CUDA (B200, sm_100a) device is required")
    return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())


def _state_version(module: nn.Module):
    """Engine-cache key: identity, storage and in-place version of every parameter / buffer.  `p.data = ...` swaps the
    storage (data_ptr changes), in-place ops bump `_version`; writes through `.data` (`p.data.copy_()`) change neither --
    call `invalidate()` after those (load_state_dict does it through a hook)."""
    return tuple((id(t), t.data_ptr(), int(t._version)) for t in list(module.parameters()) + list(module.buffers()))


# ------------------------------------------------------------------------------------------------------------------
# 1. Multi-head model (IR:28-73)
# ------------------------------------------------------------------------------------------------------------------
class BinaryClassifier(nn.Module):
    """Sub-model with a backbone + 2-output head: index 0 => Real, index 1 => Synthetic (IR:28-51)."""

    def __init__(self, model_name: str = "resnet18"):
        super().__init__()
        self.base = ResNetTrunk(model_name)
        self.model_name = model_name
        self.head = nn.Sequential(
            nn.AdaptiveAvgPool2d(1), nn.Flatten(),
            nn.Linear(self.base.num_features, 512), nn.BatchNorm1d(512), nn.ReLU(), nn.Dropout(0.5),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(), nn.Dropout(0.3),
            nn.Linear(256, 2))
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self.base._features_fn = self._features
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self) -> None:
        """Drop the folded weights held on the device; the next forward re-reads the live parameters (needed after
        writes the cache key cannot see, e.g. `p.data.copy_(...)`)."""
        self._engine_key = None

    def _own_engine(self, device) -> Engine:
        key = (str(device), _state_version(self))
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(1, device, max_batch=16, backbone=self.model_name)
            self._engine.load_head(0, self.state_dict())
            self._engine_key = key
        return self._engine

    def _run(self, x: torch.Tensor):
        dev = _cuda_device(x.device)
        eng = self._own_engine(dev)
        x = x.detach().to(dtype=torch.float32).contiguous()
        outs, feats = [], []
        for i in range(0, x.shape[0], eng.max_batch):
            xb = x[i:i + eng.max_batch]
            eng.forward_images(xb)
            outs.append(eng.debug_read(3, (xb.shape[0], 2), torch.float32))
            feats.append(eng.debug_read(2, (xb.shape[0], 16, 16, self.base.num_features), eng.act_dtype))
        return torch.cat(outs), torch.cat(feats)

    def _features(self, x: torch.Tensor) -> torch.Tensor:          # timm forward_features: [B,512,16,16]
        return self._run(x)[1].permute(0, 3, 1, 2).float().contiguous()

    def forward(self, x: torch.Tensor) -> torch.Tensor:            # IR:49-51 -> [B,2] = [Real, Synthetic]
        if self.training:
            raise RuntimeError("BinaryClassifier kernels implement eval-mode inference only; call .eval()")
        return self._run(x)[0]


class ModularMultiHeadClassifier(nn.Module):
    """Merged model: averages the Real outputs and keeps the Synthetic ones => [B, N+1] (IR:53-73)."""

    def __init__(self, sub_models: List[nn.Module]):
        super().__init__()
        self.sub_models = nn.ModuleList(sub_models)
        self._engine: Optional[Engine] = None
        self._engine_key = None
        self.max_batch = 64
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self) -> None:
        """Drop the folded weights held on the device; the next forward re-reads the live parameters."""
        self._engine_key = None
        for m in self.sub_models:
            if hasattr(m, "invalidate"):
                m.invalidate()

    def engine(self, device) -> Engine:
        """The sad_ctx holding this ensemble's folded weights (rebuilt when parameters change)."""
        dev = _cuda_device(device)
        key = (str(dev), _state_version(self), self.max_batch)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            names = {getattr(m, "model_name", "resnet18") for m in self.sub_models}
            if len(names) != 1:
                raise ValueError(f"all sub-models must share one backbone, got {sorted(names)}")
            self._engine = Engine(len(self.sub_models), dev, max_batch=self.max_batch, backbone=names.pop())
            self._engine.load_merged_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("ModularMultiHeadClassifier kernels implement eval-mode inference only; call .eval()")
        eng = self.engine(x.device)
        logits, _, _ = eng.forward_images(x.detach().to(dtype=torch.float32).contiguous())
        return logits

    # fused entry used by main(): PCM in, (logits, probs, labels) out, never materialising the 3x512x512 images
    def forward_pcm(self, pcm: torch.Tensor, threshold: float = 0.5):
        return self.engine(pcm.device).forward_pcm(pcm, threshold)


# ------------------------------------------------------------------------------------------------------------------
# 2. Loading the merged model (IR:77-123)
# ------------------------------------------------------------------------------------------------------------------
def load_merged_model(merged_path: str, device: torch.device, backbone_name: str = "resnet18"):
    """Returns (final_model, metadata); metadata must contain "class_names" (IR:77-123)."""
    device = torch.device(device)
    state = torch.load(merged_path, map_location="cpu")
    sd = state["state_dict"]
    metadata = state.get("metadata", None)
    if not metadata or "class_names" not in metadata:
        raise ValueError("Merged model checkpoint does not contain metadata for class names!")

    submodel_indices = set()
    for k in sd.keys():
        parts = k.split(".")
        if len(parts) >= 3 and parts[0] == "sub_models":
            try:
                submodel_indices.add(int(parts[1]))
            except ValueError:
                pass
    submodel_indices = sorted(submodel_indices)
    print(f"Found {len(submodel_indices)} sub-model(s): {submodel_indices}")

    sub_models = []
    for idx in submodel_indices:
        sm = BinaryClassifier(model_name=backbone_name)
        own = sm.state_dict()
        local_sd = {}
        for param_key in own.keys():                     # missing keys keep the freshly built value (IR:105-110)
            big_key = f"sub_models.{idx}." + param_key
            local_sd[param_key] = sd[big_key] if big_key in sd else own[param_key]
        sm.load_state_dict(local_sd, strict=False)
        sm.eval()
        sub_models.append(sm)

    final_model = ModularMultiHeadClassifier(sub_models)
    final_model.eval()
    if device.type == "cuda" and torch.cuda.is_available():
        final_model.to(device)
        dummy_in = torch.randn(2, 3, 512, 512, device=device)                 # quick test (IR:120-122)
        dummy_out = final_model(dummy_in)
        print("Rebuilt merged model => dummy output shape:", dummy_out.shape)
    else:
        print("Rebuilt merged model (no CUDA device: forward is unavailable, there is no CPU fallback)")
    return final_model, metadata


# ------------------------------------------------------------------------------------------------------------------
# 3. Windows + spectrogram (IR:127-190)
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class AudioConfig:
    sample_rate: int = 32000
    window_size: float = 4.0      # seconds
    overlap: float = 0.85         # fraction overlap
    silence_threshold: float = 1e-4


@dataclass
class SpectrogramConfig:
    n_fft: int = 2048
    hop_length: int = 512
    n_mels: int = 128
    f_min: int = 20
    f_max: int = 12000
    top_db: int = 80
    norm: str = "slaney"


_FRONT_ENGINE: Dict[str, Engine] = {}


def _front_engine(device) -> Engine:
    """A head-less use of the library: the front-end kernels need no weights (1 head slot is allocated, unused)."""
    dev = _cuda_device(device)
    if str(dev) not in _FRONT_ENGINE:
        _FRONT_ENGINE[str(dev)] = Engine(1, dev, max_batch=8)
    return _FRONT_ENGINE[str(dev)]


def _read_wav_pcm(path: str) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE -> (interleaved samples [frames, channels], sample rate).  16-bit PCM stays int16 (the ingest kernel
    scales by 1/32768 like torchaudio.load); 8/24/32-bit integer PCM is scaled to float32 in [-1, 1) here; 32-bit float
    stays float32."""
    import struct
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:      # WAVE_FORMAT_EXTENSIBLE: real tag = first word of the GUID
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError(f"{path}: missing fmt/data chunk")
    tag, ch, sr, _, _, bits = fmt
    if tag == 3 and bits == 32:
        x = np.frombuffer(raw[:len(raw) // 4 * 4], dtype="<f4").astype(np.float32)
    elif tag == 1 and bits == 16:
        x = np.frombuffer(raw[:len(raw) // 2 * 2], dtype="<i2").astype(np.int16)
    elif tag == 1 and bits == 32:
        x = np.frombuffer(raw[:len(raw) // 4 * 4], dtype="<i4").astype(np.float32) / 2147483648.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(raw[:len(raw) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        x = (np.where(v >= 1 << 23, v - (1 << 24), v)).astype(np.float32) / 8388608.0
    elif tag == 1 and bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format tag {tag}, {bits} bits)")
    return np.ascontiguousarray(x[:len(x) // ch * ch].reshape(-1, ch)), int(sr)


def _read_wav(path: str) -> Tuple[torch.Tensor, int]:
    """What torchaudio.load returns for the file: ([channels, T] fp32 in [-1,1), sample rate)."""
    x, sr = _read_wav_pcm(path)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / 32768.0
    return torch.from_numpy(np.ascontiguousarray(x.T)), sr


def preprocess_waveform(path: str, cfg: AudioConfig, device=None):
    """IR:144-155: load, force mono, resample to cfg.sample_rate, zero-pad to at least one window.

    The container is parsed on the host (torchaudio.load needs torchcodec, absent in this image); the PCM goes to the GPU
    as it sits in the file (int16: half the bytes of float32) and channel mix, resampling and padding run there
    (``sad_ingest``).  The waveform is returned ON THE DEVICE -- slicing, gating and the model consume it there."""
    if cfg.sample_rate != 32000:
        raise NotImplementedError(f"the ingest kernel resamples to 32 kHz only (IR:258), got {cfg.sample_rate}")
    if int(cfg.window_size * cfg.sample_rate) != SEGMENT:
        raise NotImplementedError("ingest pads to one 4-s window of 128000 samples (IR:258); other windows have no kernels")
    pcm, sr = _read_wav_pcm(path)
    eng = _front_engine(device if device is not None else "cuda")
    wf = eng.ingest(torch.from_numpy(pcm).to(eng.device), sr)
    return wf, cfg.sample_rate


def _check_spec_cfg(sr: int, spec_cfg: SpectrogramConfig):
    want = SpectrogramConfig()
    if sr != 32000 or any(getattr(spec_cfg, k) != getattr(want, k) for k in want.__dataclass_fields__):
        raise NotImplementedError(
            "the sm_100a front end implements the reference's fixed geometry only (IR:258-259): 32 kHz, n_fft 2048, "
            f"hop 512, 128 mels, 20-12000 Hz, top_db 80, slaney; got sr={sr}, {spec_cfg}")


def waveform_to_spectrogram(waveform: torch.Tensor, sr: int, spec_cfg: SpectrogramConfig):
    """IR:157-174: [128000] -> [1,3,512,512] fp32, returned on the input's device."""
    _check_spec_cfg(sr, spec_cfg)
    if waveform.dim() != 1 or waveform.shape[0] != SEGMENT:
        raise ValueError(f"expected one 4-s 32 kHz segment of {SEGMENT} samples, got shape {tuple(waveform.shape)}")
    src = waveform.device
    dev = src if src.type == "cuda" else torch.device("cuda")
    eng = _front_engine(dev)
    img = eng.image(waveform.detach().to(eng.device, torch.float32).contiguous().unsqueeze(0))     # [1,512,512]
    spec3 = img.repeat(3, 1, 1)                                                                       # IR:173
    return spec3.unsqueeze(0).to(src)


def window_and_hop(sr: int, cfg: AudioConfig) -> Tuple[int, int]:
    """IR:180-181, python float arithmetic truncated with int()."""
    window_samples = int(cfg.window_size * sr)
    hop_samples = int((1 - cfg.overlap) * window_samples)
    if hop_samples == 0:
        raise ValueError("range() arg 3 must not be zero")        # what the reference's range(...) raises (IR:184)
    return window_samples, hop_samples


def slice_waveform(wf: torch.Tensor, sr: int, cfg: AudioConfig):
    """IR:176-190: returns (chunks, timestamps); windows whose max |x| < silence_threshold are dropped."""
    window_samples, hop_samples = window_and_hop(sr, cfg)
    dev = wf.device if wf.device.type == "cuda" else torch.device("cuda")
    eng = _front_engine(dev)
    wd = wf.detach().to(eng.device, torch.float32).contiguous()
    keep = eng.slice_gate(wd, window_samples, hop_samples, cfg.silence_threshold).cpu().numpy().astype(bool)
    chunks, timestamps = [], []
    for w in np.nonzero(keep)[0]:
        start_idx = int(w) * hop_samples
        chunks.append(wf[start_idx:start_idx + window_samples])
        timestamps.append(start_idx / sr)
    return chunks, timestamps


# ------------------------------------------------------------------------------------------------------------------
# 4. Probability interpretation (IR:194-214)
# ------------------------------------------------------------------------------------------------------------------
def label_from_index(idx: int, n: int, synthetic_names: Optional[List[str]], real_name: str) -> str:
    if idx == n:
        return real_name
    if synthetic_names and idx < len(synthetic_names):
        return synthetic_names[idx]
    return f"Synthetic_{idx + 1}"


def interpret_multihead_logits(logits: torch.Tensor, threshold=0.5, synthetic_names: List[str] = None,
                               real_name: str = "Real"):
    """One row of [N+1] logits -> (label, sigmoid probabilities as numpy).  Kept for API compatibility; main() gets
    labels and probabilities for the whole batch from the merge/decision kernel instead of calling this per row."""
    s = torch.sigmoid(logits)
    n = s.shape[0] - 1
    syn_probs, real_prob = s[:n], s[-1]
    if real_prob >= threshold and bool((syn_probs < threshold).all()):
        idx = n
    else:
        idx = int(torch.argmax(syn_probs).item())
    return label_from_index(idx, n, synthetic_names, real_name), s.detach().cpu().numpy()


def smooth_probabilities(raw_probs: np.ndarray, threshold: float):
    """IR:301-325 (--smooth): gaussian sigma=2 per column, renormalise rows, re-label.  Host side, as in the reference."""
    from scipy.ndimage import gaussian_filter1d
    arr = np.array(raw_probs)
    for dim in range(arr.shape[1]):
        arr[:, dim] = gaussian_filter1d(arr[:, dim], sigma=2)
    for i in range(arr.shape[0]):
        row_sum = arr[i].sum()
        if row_sum > 0:
            arr[i] /= row_sum
    n = arr.shape[1] - 1
    idx = [n if (row[-1] >= threshold and (row[:-1] < threshold).all()) else int(row[:-1].argmax()) for row in arr]
    return arr, idx


# ------------------------------------------------------------------------------------------------------------------
# 5. Main (IR:218-353)
# ------------------------------------------------------------------------------------------------------------------
def analyze_waveform(model: ModularMultiHeadClassifier, wf: torch.Tensor, sr: int, class_names: List[str],
                     audio_cfg: AudioConfig, threshold: float, smooth: bool, device) -> Dict:
    """Windows -> fused PCM->decision kernels -> per-window labels and clip percentages (IR:262-343)."""
    synthetic_names, real_name = class_names[:-1], class_names[-1]
    eng = model.engine(device)
    window_samples, hop_samples = window_and_hop(sr, audio_cfg)
    wd = wf.to(eng.device, torch.float32).contiguous()
    keep = eng.slice_gate(wd, window_samples, hop_samples, audio_cfg.silence_threshold)
    starts = torch.nonzero(keep).flatten().to(torch.int64) * hop_samples
    if starts.numel() == 0:
        return {"segments": [], "percentages": {}}
    if window_samples != SEGMENT:
        raise NotImplementedError("kernels are built for 4-s 32 kHz windows (IR:258)")
    pcm = eng.gather_windows(wd, starts, window_samples)
    _, probs, labels = eng.forward_pcm(pcm, threshold)
    clip_id = torch.zeros(pcm.shape[0], dtype=torch.int32, device=eng.device)
    clip_probs, _ = eng.clip_reduce(probs, clip_id, 1, threshold)                 # IR:328: mean over windows
    n = len(class_names) - 1
    idx = labels.cpu().tolist()
    final = clip_probs[0].cpu().numpy()
    if smooth:
        arr, idx = smooth_probabilities(probs.cpu().numpy(), threshold)
        final = np.mean(arr.tolist(), axis=0)                                     # IR:325-328 (python floats => f64)
    names = [label_from_index(i, n, synthetic_names, real_name) for i in idx]
    prob_dict = {}
    for i in range(n):
        prob_dict[synthetic_names[i] if i < len(synthetic_names) else f"Synthetic_{i + 1}"] = float(final[i] * 100)
    prob_dict[real_name] = float(final[-1] * 100)
    stamps = (starts.cpu().numpy() / sr).tolist()
    segments = [{"start_sec": t, "end_sec": t + audio_cfg.window_size, "label": lbl} for t, lbl in zip(stamps, names)]
    return {"segments": segments, "percentages": prob_dict}


def main(argv=None):
    parser = argparse.ArgumentParser(
        description="Multi-head inference with overlapping windows using metadata from the merged model.")
    parser.add_argument("--merged-model", type=str, required=True, help="Path to merged .pth")
    parser.add_argument("--audio", type=str, required=True, help="Path to WAV file")
    parser.add_argument("--threshold", type=float, default=0.5, help="Threshold for deciding Real vs Synthetic")
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--confidence-threshold", type=float, default=0.45, help="Confidence threshold for segments.")
    parser.add_argument("--smooth", action="store_true", help="Apply smoothing across windows.")
    parser.add_argument("--output-json", type=str, default="results.json")
    args = parser.parse_args(argv)

    seed = 9                                               # IR:232-241
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    device = _cuda_device(args.device)
    torch.cuda.manual_seed_all(seed)

    model, metadata = load_merged_model(args.merged_model, device)
    model.eval()
    class_names = metadata["class_names"]
    print("Using metadata names:")
    print("Synthetic names:", class_names[:-1])
    print("Real name:", class_names[-1])

    audio_cfg = AudioConfig(sample_rate=32000, window_size=4.0, overlap=0.0, silence_threshold=1e-3)   # IR:258
    wf, sr = preprocess_waveform(args.audio, audio_cfg)
    res = analyze_waveform(model, wf, sr, class_names, audio_cfg, args.threshold, args.smooth, device)
    if not res["segments"]:
        print("No valid audio chunks found (all below silence threshold). Exiting.")
    out_json = {"filename": args.audio, "segments": res["segments"], "percentages": res["percentages"]}
    with open(args.output_json, "w", encoding="utf-8") as f:
        json.dump(out_json, f, indent=4)
    if res["segments"]:
        print("Wrote results to", args.output_json)
        print(json.dumps(out_json, indent=4))


if __name__ == "__main__":
    main()
