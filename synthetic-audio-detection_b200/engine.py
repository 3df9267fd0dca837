"""Engine: one sad_ctx (one GPU) driven from PyTorch tensors.  Host-side plumbing only: every number is produced
by the CUDA kernels behind include/sad_b200.h."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

SEGMENT = 128000


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _frontend_constants() -> Tuple[torch.Tensor, torch.Tensor]:
    """Hann window and mel filterbank exactly as the reference's torchaudio.transforms.MelSpectrogram builds them
    (inference_runner.py:158-166): same formulas, fp32 torch arithmetic, so the uploaded constants are bit-identical."""
    import math
    window = torch.hann_window(2048, periodic=True, dtype=torch.float32)
    all_freqs = torch.linspace(0, 32000 // 2, 1025)
    m_min = 2595.0 * math.log10(1.0 + 20.0 / 700.0)
    m_max = 2595.0 * math.log10(1.0 + 12000.0 / 700.0)
    m_pts = torch.linspace(m_min, m_max, 130)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.max(torch.zeros(1), torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))
    fb = fb * (2.0 / (f_pts[2:130] - f_pts[:128])).unsqueeze(0)
    return window.contiguous(), fb.contiguous()


def default_dtype(backbone: str) -> str:
    return "fp16" if backbone in ("resnet50", "resnet101", "resnet152") else "bf16"


class Engine:
    """Owns a sad_ctx.  Tensors passed in must be CUDA fp32 contiguous on this engine's device."""

    def __init__(self, n_heads: int, device: Optional[torch.device] = None, max_batch: int = 64,
                 backbone: str = "resnet18", dtype: Optional[str] = None):
        """dtype: element type of activations / conv weights, "bf16" or "fp16" (two builds of the same kernels,
        csrc/act.cuh).  Default: bf16 -- the dtype BASELINE.json names -- for the BasicBlock nets (resnet18/34), fp16 for
        the Bottleneck nets (resnet50/101/152), whose 53+ convolutions do not meet the 2e-2 logit bound with bf16
        storage.  SAD_DTYPE in the environment overrides the default."""
        import os
        if dtype is None:
            dtype = os.environ.get("SAD_DTYPE") or default_dtype(backbone)
        self.dtype = dtype
        self.act_dtype = torch.float16 if dtype == "fp16" else torch.bfloat16
        self.lib = _lib.load(dtype)
        if not torch.cuda.is_available():
            raise _lib.SadError("no CUDA device: the sm_100a kernels cannot run and there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise _lib.SadError(f"device {self.device} is not CUDA; there is no CPU fallback")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.n_heads = int(n_heads)
        self.max_batch = int(max_batch)
        self.ctx = C.c_void_p(0)
        self.backbone = backbone
        if self.lib.sad_backbone_weight_count(backbone.encode()) < 0:
            raise NotImplementedError(f"backbone {backbone!r} has no sm_100a kernels (resnet18/34/50/101/152 only)")
        torch.cuda.init()
        code = self.lib.sad_create_ex(C.byref(self.ctx), idx, self.n_heads, self.max_batch, backbone.encode())
        if code != 0:
            msg = self.lib.sad_last_error(self.ctx).decode() if self.ctx else ""
            if self.ctx:
                self.lib.sad_destroy(self.ctx)
            self.ctx = C.c_void_p(0)
            raise _lib.SadError(f"sad_create failed ({code}): {msg}")
        w, fb = _frontend_constants()
        _lib.check(self.ctx, self.lib.sad_set_frontend_constants(self.ctx, _ptr(w), _ptr(fb)), "sad_set_frontend_constants")
        bb = backbone.encode()
        n = self.lib.sad_backbone_weight_count(bb)
        self._names = [self.lib.sad_backbone_weight_name(bb, i).decode() for i in range(n)]
        self._numel = [self.lib.sad_backbone_weight_numel(bb, i) for i in range(n)]

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sad_destroy(self.ctx)
            self.ctx = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_head(self, head: int, sd: Dict[str, torch.Tensor], prefix: str = ""):
        """Upload one BinaryClassifier's tensors: sd[prefix + name] for every name the library lists."""
        keep: List[torch.Tensor] = []
        arr = (C.c_void_p * len(self._names))()
        for i, (name, numel) in enumerate(zip(self._names, self._numel)):
            key = prefix + name
            if key not in sd:
                raise KeyError(f"missing tensor {key!r}")
            t = sd[key].detach().to(device="cpu", dtype=torch.float32).contiguous()
            if t.numel() != numel:
                raise ValueError(f"{key}: expected {numel} elements, got {t.numel()}")
            keep.append(t)
            arr[i] = t.data_ptr()
        _lib.check(self.ctx, self.lib.sad_load_weights(self.ctx, head, arr, len(self._names)), "sad_load_weights")

    def load_merged_state_dict(self, sd: Dict[str, torch.Tensor], indices: Optional[List[int]] = None):
        """Merged layout of model_merger.py:154-159: keys 'sub_models.<i>.<name>'; heads are packed densely in
        sorted index order (inference_runner.py:89-98)."""
        if indices is None:
            found = set()
            for k in sd:
                parts = k.split(".")
                if len(parts) >= 3 and parts[0] == "sub_models":
                    try:
                        found.add(int(parts[1]))
                    except ValueError:
                        pass
            indices = sorted(found)
        if len(indices) != self.n_heads:
            raise ValueError(f"checkpoint has {len(indices)} sub-models, engine was created for {self.n_heads}")
        for slot, idx in enumerate(indices):
            self.load_head(slot, sd, prefix=f"sub_models.{idx}.")

    # ------------------------------------------------------------------ compute
    def _check(self, x: torch.Tensor, shape_tail):
        if not x.is_cuda or x.device != self.device:
            raise _lib.SadError(f"input must live on {self.device} (got {x.device}); there is no CPU fallback")
        if x.dtype != torch.float32 or not x.is_contiguous():
            raise ValueError("input must be contiguous float32")
        if tuple(x.shape[1:]) != tuple(shape_tail):
            raise ValueError(f"expected shape [B,{','.join(map(str, shape_tail))}], got {tuple(x.shape)}")

    def logmel(self, pcm: torch.Tensor, want_stats: bool = True):
        self._check(pcm, (SEGMENT,))
        B = pcm.shape[0]
        db = torch.empty(B, 128, 251, device=self.device, dtype=torch.float32)
        ms = torch.empty(B, 2, device=self.device, dtype=torch.float32) if want_stats else None
        _lib.check(self.ctx, self.lib.sad_frontend_logmel(self.ctx, _ptr(pcm), B, _ptr(db), _ptr(ms), _stream(self.device)),
                   "sad_frontend_logmel")
        return db, ms

    def logmel_into(self, pcm: torch.Tensor, db: torch.Tensor, mu_sigma: Optional[torch.Tensor] = None):
        """As logmel(), into caller-provided CUDA buffers (db [B,128,251] fp32, mu_sigma [B,2] fp32 or None)."""
        self._check(pcm, (SEGMENT,))
        self._check(db, (128, 251))
        _lib.check(self.ctx, self.lib.sad_frontend_logmel(self.ctx, _ptr(pcm), pcm.shape[0], _ptr(db), _ptr(mu_sigma),
                                                          _stream(self.device)), "sad_frontend_logmel")

    def image(self, pcm: torch.Tensor) -> torch.Tensor:
        self._check(pcm, (SEGMENT,))
        B = pcm.shape[0]
        img = torch.empty(B, 512, 512, device=self.device, dtype=torch.float32)
        _lib.check(self.ctx, self.lib.sad_frontend_image(self.ctx, _ptr(pcm), B, _ptr(img), _stream(self.device)),
                   "sad_frontend_image")
        return img

    def _outputs(self, B):
        n1 = self.n_heads + 1
        return (torch.empty(B, n1, device=self.device, dtype=torch.float32),
                torch.empty(B, n1, device=self.device, dtype=torch.float32),
                torch.empty(B, device=self.device, dtype=torch.int32))

    def forward_pcm(self, pcm: torch.Tensor, threshold: float = 0.5):
        """[B,128000] fp32 CUDA -> (logits [B,N+1], probs [B,N+1], labels [B] int32; N == Real)."""
        self._check(pcm, (SEGMENT,))
        B = pcm.shape[0]
        lo, pr, la = self._outputs(B)
        _lib.check(self.ctx, self.lib.sad_forward(self.ctx, _ptr(pcm), B, threshold, _ptr(lo), _ptr(pr), _ptr(la),
                                                  _stream(self.device)), "sad_forward")
        return lo, pr, la

    def forward_images(self, x: torch.Tensor, threshold: float = 0.5):
        self._check(x, (3, 512, 512))
        B = x.shape[0]
        lo, pr, la = self._outputs(B)
        _lib.check(self.ctx, self.lib.sad_forward_images(self.ctx, _ptr(x), B, threshold, _ptr(lo), _ptr(pr), _ptr(la),
                                                         _stream(self.device)), "sad_forward_images")
        return lo, pr, la

    def forward_host(self, pcm: torch.Tensor, threshold: float = 0.5):
        """End-to-end with HOST tensors (pinned or pageable): H2D + compute + D2H inside the C call."""
        if pcm.is_cuda or pcm.dtype != torch.float32 or not pcm.is_contiguous() or pcm.shape[1] != SEGMENT:
            raise ValueError("forward_host wants a contiguous float32 CPU tensor [B,128000]")
        B = pcm.shape[0]
        n1 = self.n_heads + 1
        lo = torch.empty(B, n1, dtype=torch.float32)
        pr = torch.empty(B, n1, dtype=torch.float32)
        la = torch.empty(B, dtype=torch.int32)
        _lib.check(self.ctx, self.lib.sad_forward_host(self.ctx, _ptr(pcm), B, threshold, _ptr(lo), _ptr(pr), _ptr(la)),
                   "sad_forward_host")
        return lo, pr, la

    def clip_reduce(self, probs: torch.Tensor, clip_id: torch.Tensor, n_clips: int, threshold: float = 0.5):
        """probs [B,N+1] fp32, clip_id [B] int32 sorted ascending -> (clip_probs [n_clips,N+1], clip_label [n_clips])."""
        B = probs.shape[0]
        cp = torch.empty(n_clips, self.n_heads + 1, device=self.device, dtype=torch.float32)
        cl = torch.empty(n_clips, device=self.device, dtype=torch.int32)
        _lib.check(self.ctx, self.lib.sad_clip_reduce(self.ctx, _ptr(probs), _ptr(clip_id), B, n_clips, threshold, _ptr(cp),
                                                      _ptr(cl), _stream(self.device)), "sad_clip_reduce")
        return cp, cl

    def ingest(self, pcm: torch.Tensor, sr_in: int) -> torch.Tensor:
        """Interleaved PCM [frames, channels] (int16 or float32, CUDA) -> mono fp32 at 32 kHz, zero-padded to >= 128000
        samples: mean over channels, torchaudio-default sinc resampling, padding (reference IR:144-155)."""
        if not pcm.is_cuda or pcm.device != self.device:
            raise _lib.SadError(f"pcm must live on {self.device} (got {pcm.device}); there is no CPU fallback")
        if pcm.dim() != 2 or not pcm.is_contiguous() or pcm.dtype not in (torch.int16, torch.float32):
            raise ValueError("ingest wants a contiguous [frames, channels] int16 or float32 tensor")
        frames, ch = pcm.shape
        n = self.lib.sad_ingest_length(frames, int(sr_in))
        if n < 0:
            raise ValueError(f"bad ingest arguments: frames {frames}, sample rate {sr_in}")
        out = torch.empty(n, device=self.device, dtype=torch.float32)
        fmt = 0 if pcm.dtype == torch.int16 else 1
        _lib.check(self.ctx, self.lib.sad_ingest(self.ctx, _ptr(pcm), fmt, frames, ch, int(sr_in), _ptr(out),
                                                 _stream(self.device)), "sad_ingest")
        return out

    def slice_gate(self, wf: torch.Tensor, window: int, hop: int, silence_threshold: float) -> torch.Tensor:
        n = self.lib.sad_slice_count(wf.shape[0], window, hop)
        keep = torch.zeros(max(n, 0), device=self.device, dtype=torch.uint8)
        if n > 0:
            _lib.check(self.ctx, self.lib.sad_slice_gate(self.ctx, _ptr(wf), wf.shape[0], window, hop, silence_threshold,
                                                         _ptr(keep), _stream(self.device)), "sad_slice_gate")
        return keep

    def gather_windows(self, wf: torch.Tensor, starts: torch.Tensor, window: int) -> torch.Tensor:
        n = starts.shape[0]
        out = torch.empty(n, window, device=self.device, dtype=torch.float32)
        if n > 0:
            _lib.check(self.ctx, self.lib.sad_gather_windows(self.ctx, _ptr(wf), _ptr(starts), n, window, _ptr(out),
                                                             _stream(self.device)), "sad_gather_windows")
        return out

    def debug_conv(self, head: int, layer: int, x: torch.Tensor, residual: Optional[torch.Tensor], out_shape, relu: bool):
        out = torch.empty(out_shape, device=self.device, dtype=self.act_dtype)
        _lib.check(self.ctx, self.lib.sad_debug_conv(self.ctx, head, layer, _ptr(x), _ptr(residual), _ptr(out), x.shape[0],
                                                     int(relu), _stream(self.device)), "sad_debug_conv")
        return out

    def debug_block(self, head: int, layer: int, x: torch.Tensor) -> torch.Tensor:
        """One fused layer1 BasicBlock (convs `layer`, `layer`+1 + identity) on x [B,128,128,64] bf16."""
        out = torch.empty_like(x)
        _lib.check(self.ctx, self.lib.sad_debug_block(self.ctx, head, layer, _ptr(x), _ptr(out), x.shape[0],
                                                      _stream(self.device)), "sad_debug_block")
        return out

    def debug_stem(self, pcm: torch.Tensor) -> torch.Tensor:
        """Pooled stem output [H*B,128,128,64] bf16 for pcm [B,128000] (B <= max_batch)."""
        self._check(pcm, (SEGMENT,))
        B = pcm.shape[0]
        out = torch.empty(self.n_heads * B, 128, 128, 64, device=self.device, dtype=self.act_dtype)
        _lib.check(self.ctx, self.lib.sad_debug_stem(self.ctx, _ptr(pcm), B, _ptr(out), _stream(self.device)),
                   "sad_debug_stem")
        return out

    def debug_read(self, which: int, shape, dtype) -> torch.Tensor:
        out = torch.empty(shape, device=self.device, dtype=dtype)
        n = self.lib.sad_debug_read(self.ctx, which, _ptr(out), out.numel() * out.element_size(), _stream(self.device))
        _lib.check(self.ctx, int(n), "sad_debug_read")
        return out

    PROF_KINDS = 44
    PROF_FRONTEND, PROF_IMAGE, PROF_POOL, PROF_HEAD = 40, 41, 42, 43

    def profile_enable(self, on: bool = True):
        _lib.check(self.ctx, self.lib.sad_profile_enable(self.ctx, int(on)), "sad_profile_enable")

    def profile_read(self):
        """(ms_by_kind[44], launches_by_kind[44]); kinds 0..39 = convolutions in state_dict order (0 = stem), 40 front
        end, 41 image (+im2col), 42 max pool (3-channel path), 43 head/merge."""
        ms = (C.c_double * self.PROF_KINDS)()
        n = (C.c_longlong * self.PROF_KINDS)()
        _lib.check(self.ctx, self.lib.sad_profile_read(self.ctx, ms, n), "sad_profile_read")
        return list(ms), list(n)

    @property
    def launches(self) -> int:
        return int(self.lib.sad_launch_count(self.ctx))
