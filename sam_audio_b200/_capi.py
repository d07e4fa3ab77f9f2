"""ctypes binding of libsamaudio_b200.so (include/samaudio_b200.h).

PyTorch is used only as the owner of device memory and streams: every call
passes raw ``data_ptr()`` values and the current CUDA stream.  There is no
fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_LIB_NAME = "libsamaudio_b200.so"
_lib: Optional[ctypes.CDLL] = None


class SabConfig(ctypes.Structure):
    _fields_ = [
        ("dim", ctypes.c_int32), ("n_heads", ctypes.c_int32), ("n_layers", ctypes.c_int32),
        ("ffn_hidden", ctypes.c_int32), ("out_channels", ctypes.c_int32), ("in_channels", ctypes.c_int32),
        ("text_dim", ctypes.c_int32), ("vision_dim", ctypes.c_int32), ("n_anchor_tokens", ctypes.c_int32),
        ("anchor_dim", ctypes.c_int32), ("max_positions", ctypes.c_int32),
        ("rope_theta", ctypes.c_float), ("norm_eps", ctypes.c_float),
        ("codec_encoder_dim", ctypes.c_int32), ("codec_latent_dim", ctypes.c_int32),
        ("codec_decoder_dim", ctypes.c_int32), ("codec_codebook_dim", ctypes.c_int32),
        ("codec_n_rates", ctypes.c_int32),
        ("codec_encoder_rates", ctypes.c_int32 * 8), ("codec_decoder_rates", ctypes.c_int32 * 8),
    ]


class SabT5Config(ctypes.Structure):
    _fields_ = [("vocab_size", ctypes.c_int32), ("d_model", ctypes.c_int32), ("d_kv", ctypes.c_int32),
                ("d_ff", ctypes.c_int32), ("n_layers", ctypes.c_int32), ("n_heads", ctypes.c_int32),
                ("n_buckets", ctypes.c_int32), ("eps", ctypes.c_float)]


EXPORTS = [
    "sab_last_error", "sab_version", "sab_create", "sab_destroy", "sab_load_weight", "sab_finalize_weights",
    "sab_encode", "sab_prepare", "sab_dit_forward", "sab_solve", "sab_decode", "sab_launch_count",
    "sab_workspace_bytes", "sab_profile", "sab_profile_report", "sab_test_gemm", "sab_test_attention", "sab_test_attention_tc", "sab_test_attention_tc2", "sab_preprocess_frames", "sab_test_aa_taps", "sab_test_solver_grid",
    "sab_t5_create", "sab_t5_destroy", "sab_t5_load_weight", "sab_t5_finalize", "sab_t5_forward", "sab_t5_launch_count",
]


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{_LIB_NAME} not built ({path}); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "The B200 separation path has no CPU/PyTorch fallback.")
        L = ctypes.CDLL(path)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        L.sab_last_error.restype = ctypes.c_char_p
        L.sab_create.argtypes = [ctypes.POINTER(SabConfig), i32, ctypes.POINTER(vp)]
        L.sab_destroy.argtypes = [vp]
        L.sab_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32, i32, vp]
        L.sab_finalize_weights.argtypes = [vp, i32, ctypes.c_char_p, i64, vp]
        L.sab_encode.argtypes = [vp, vp, i32, i64, vp, vp]
        L.sab_prepare.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, i32, vp]
        L.sab_dit_forward.argtypes = [vp, vp, vp, vp, vp]
        L.sab_solve.argtypes = [vp, vp, i32, i32, vp, vp]
        L.sab_decode.argtypes = [vp, vp, i32, i32, vp, vp]
        L.sab_launch_count.argtypes = [vp, i32]
        L.sab_launch_count.restype = i64
        L.sab_workspace_bytes.argtypes = [vp]
        L.sab_workspace_bytes.restype = i64
        L.sab_profile.argtypes = [vp, i32, vp]
        L.sab_profile_report.argtypes = [vp, ctypes.c_char_p, i64, vp]
        L.sab_test_gemm.argtypes = [i32, i32, i32, vp, vp, vp, i32, i32, i32, vp]
        L.sab_test_attention.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        L.sab_test_attention_tc.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, vp]
        L.sab_test_attention_tc2.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, ctypes.c_float, i32, vp, vp]
        L.sab_preprocess_frames.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp]
        L.sab_test_aa_taps.argtypes = [i32, i32, i32, vp, vp, vp, vp]
        L.sab_test_solver_grid.argtypes = [i32, i32, i32, vp, vp]
        L.sab_t5_create.argtypes = [ctypes.POINTER(SabT5Config), i32, ctypes.POINTER(vp)]
        L.sab_t5_destroy.argtypes = [vp]
        L.sab_t5_load_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32, i32, vp]
        L.sab_t5_finalize.argtypes = [vp, vp]
        L.sab_t5_forward.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
        L.sab_t5_launch_count.argtypes = [vp, i32]
        L.sab_t5_launch_count.restype = i64
        _lib = L
    return _lib


PREP_NO_VIDEO_TERM, PREP_NO_ANCHORS, PREP_NO_TEXT = 1, 2, 4   # include/samaudio_b200.h SAB_PREP_*


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("samaudio_b200: " + lib().sab_last_error().decode("utf-8", "replace"))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def make_config(cfg) -> SabConfig:
    """cfg: sam_audio_b200.config.SAMAudioConfig"""
    tc, cc = cfg.transformer, cfg.audio_codec
    c = SabConfig()
    c.dim, c.n_heads, c.n_layers, c.ffn_hidden = tc.dim, tc.n_heads, tc.n_layers, tc.ffn_hidden
    c.out_channels, c.in_channels = tc.out_channels, cfg.in_channels
    c.text_dim, c.vision_dim = cfg.text_encoder.dim, cfg.vision_encoder.dim
    c.n_anchor_tokens, c.anchor_dim = cfg.num_anchors + 1, cfg.anchor_embedding_dim
    c.max_positions, c.rope_theta, c.norm_eps = tc.max_positions, tc.rope_theta, tc.norm_eps
    c.codec_encoder_dim, c.codec_latent_dim = cc.encoder_dim, cc.latent_dim
    c.codec_decoder_dim, c.codec_codebook_dim = cc.decoder_dim, cc.codebook_dim
    c.codec_n_rates = len(cc.encoder_rates)
    assert len(cc.decoder_rates) == len(cc.encoder_rates) <= 8
    for i, r in enumerate(cc.encoder_rates):
        c.codec_encoder_rates[i] = r
    for i, r in enumerate(cc.decoder_rates):
        c.codec_decoder_rates[i] = r
    return c


def preprocess_frames(frames: torch.Tensor, out_size: int) -> torch.Tensor:
    """uint8 [N, 3, H, W] (cuda) -> fp32 [N, 3, S, S]: antialiased bicubic resize, uint8 rounding, (x/255 - .5)/.5."""
    assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[1] == 3
    frames = frames.contiguous()
    n, _, h, w = frames.shape
    out = torch.empty(n, 3, out_size, out_size, device=frames.device, dtype=torch.float32)
    ws = torch.empty(n, 3, h, out_size, device=frames.device, dtype=torch.float32)
    with torch.cuda.device(frames.device):
        check(lib().sab_preprocess_frames(frames.data_ptr(), n, h, w, out_size, ws.data_ptr(), out.data_ptr(), stream_ptr()))
    return out


class Engine:
    """Owns one sab_engine handle."""

    def __init__(self, cfg, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("samaudio_b200: CUDA device required (no CPU fallback)")
        self._h = ctypes.c_void_p()
        self._cfg = make_config(cfg)
        self.device = device
        check(lib().sab_create(ctypes.byref(self._cfg), device, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().sab_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weight(self, name: str, t: torch.Tensor):
        t = t.detach().to(torch.float32).contiguous()
        shape = (ctypes.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
        check(lib().sab_load_weight(self._h, name.encode(), t.data_ptr(), shape, max(t.dim(), 1),
                                    1 if t.is_cuda else 0, stream_ptr()))

    def finalize(self, allow_missing: bool = False):
        """Returns the list of missing keys (empty unless allow_missing)."""
        import ctypes as C
        buf = C.create_string_buffer(1 << 16)
        check(lib().sab_finalize_weights(self._h, 1 if allow_missing else 0, buf, len(buf), stream_ptr()))
        return [k for k in buf.value.decode().split("\n") if k]

    def encode(self, wav: torch.Tensor, features: torch.Tensor):
        B, S = wav.shape
        check(lib().sab_encode(self._h, wav.data_ptr(), B, S, features.data_ptr(), stream_ptr()))

    def prepare(self, B, candidates, T, L, features, text, text_mask, video, anchor_ids, anchor_alignment, pad_mask,
                flags=0):
        check(lib().sab_prepare(self._h, B, candidates, T, L, features.data_ptr(), ptr(text), text_mask.data_ptr(),
                                ptr(video), ptr(anchor_ids), 0 if anchor_ids is None else anchor_ids.shape[1],
                                ptr(anchor_alignment), pad_mask.data_ptr(), flags, stream_ptr()))

    def dit_forward(self, noisy, time, out):
        check(lib().sab_dit_forward(self._h, noisy.data_ptr(), time.data_ptr(), out.data_ptr(), stream_ptr()))

    ODE_METHODS = {"midpoint": 0, "euler": 1, "rk4": 2}    # include/samaudio_b200.h SAB_ODE_*

    def solve(self, noise, n_steps, out, method: str = "midpoint"):
        check(lib().sab_solve(self._h, noise.data_ptr(), n_steps, self.ODE_METHODS[method], out.data_ptr(), stream_ptr()))

    def decode(self, latent, Bc, T, wav):
        check(lib().sab_decode(self._h, latent.data_ptr(), Bc, T, wav.data_ptr(), stream_ptr()))

    def launch_count(self, reset=False) -> int:
        return int(lib().sab_launch_count(self._h, 1 if reset else 0))

    def profile(self, enable: bool):
        check(lib().sab_profile(self._h, 1 if enable else 0, stream_ptr()))

    def profile_report(self) -> dict:
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        check(lib().sab_profile_report(self._h, buf, len(buf), stream_ptr()))
        return json.loads(buf.value.decode())

    def workspace_bytes(self) -> int:
        return int(lib().sab_workspace_bytes(self._h))


class T5Engine:
    """Owns one sab_t5 handle (native T5 text encoder)."""

    def __init__(self, hf_config, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("samaudio_b200: CUDA device required (no CPU fallback)")
        c = SabT5Config()
        c.vocab_size, c.d_model, c.d_kv, c.d_ff = hf_config.vocab_size, hf_config.d_model, hf_config.d_kv, hf_config.d_ff
        c.n_layers, c.n_heads = hf_config.num_layers, hf_config.num_heads
        c.n_buckets, c.eps = hf_config.relative_attention_num_buckets, hf_config.layer_norm_epsilon
        self._cfg = c
        self._h = ctypes.c_void_p()
        check(lib().sab_t5_create(ctypes.byref(c), device, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().sab_t5_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, sd):
        for k, v in sd.items():
            t = v.detach().to(torch.float32).contiguous()
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            check(lib().sab_t5_load_weight(self._h, k.encode(), t.data_ptr(), shape, t.dim(), 1 if t.is_cuda else 0,
                                           stream_ptr()))
        check(lib().sab_t5_finalize(self._h, stream_ptr()))

    def forward(self, ids, mask_u8, rel_bucket, out):
        B, L = ids.shape
        check(lib().sab_t5_forward(self._h, ids.data_ptr(), mask_u8.data_ptr(), rel_bucket.data_ptr(), B, L,
                                   out.data_ptr(), stream_ptr()))

    def launch_count(self, reset=False) -> int:
        return int(lib().sab_t5_launch_count(self._h, 1 if reset else 0))
