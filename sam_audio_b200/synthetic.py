"""Seeded random-init weights and synthetic clips (no network, no checkpoints).

The state dict uses the reference's parameter names (measured
from an instantiation of sam_audio/model/model.py:79-102) so that the same
loader path serves real checkpoints and synthetic ones.  Codec names follow the
Descript-DAC module layout that ``dacvae`` derives from (source absent from the
reference tree; see DESIGN.md "parity unpinned").

Initialisation: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for every linear / conv
weight and bias (torch's default scale), norm weights 1 + 0.1 N(0,1), modulation
tables N(0,1)/sqrt(d), both tanh-gates 0.5 so the align / anchor branches are
exercised (they initialise to 0 in the reference: align.py:26-27, model.py:51).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .config import SAMAudioConfig

DESCRIPTIONS = ["man speaking", "dog barking", "car honking", "thunder"]


def _uniform(gen, shape, fan_in, device):
    b = 1.0 / math.sqrt(max(fan_in, 1))
    return (torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * b


def _randn(gen, shape, device, scale=1.0):
    return torch.randn(shape, generator=gen, device=device, dtype=torch.float32) * scale


def codec_param_shapes(cfg) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for the DAC-VAE encoder / bottleneck / decoder.
    kind in {"conv", "convT", "bias", "alpha"}."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(name, co, ci, k):
        out.append((f"{name}.weight", (co, ci, k), "conv"))
        out.append((f"{name}.bias", (co,), "bias"))

    def convT(name, ci, co, k):
        out.append((f"{name}.weight", (ci, co, k), "convT"))
        out.append((f"{name}.bias", (co,), "bias"))

    def snake(name, c):
        out.append((f"{name}.alpha", (1, c, 1), "alpha"))

    def res_unit(name, c):
        snake(f"{name}.block.0", c)
        conv(f"{name}.block.1", c, c, 7)
        snake(f"{name}.block.2", c)
        conv(f"{name}.block.3", c, c, 1)

    d = cfg.encoder_dim
    conv("encoder.block.0", d, 1, 7)
    for i, s in enumerate(cfg.encoder_rates):
        d *= 2
        p = f"encoder.block.{i + 1}"
        for j in range(3):
            res_unit(f"{p}.block.{j}", d // 2)
        snake(f"{p}.block.3", d // 2)
        conv(f"{p}.block.4", d, d // 2, 2 * s)
    snake(f"encoder.block.{len(cfg.encoder_rates) + 1}", d)
    conv(f"encoder.block.{len(cfg.encoder_rates) + 2}", cfg.latent_dim, d, 3)
    conv("quantizer.in_proj", 2 * cfg.codebook_dim, cfg.latent_dim, 1)
    conv("quantizer.out_proj", cfg.latent_dim, cfg.codebook_dim, 1)
    ch = cfg.decoder_dim
    conv("decoder.model.0", ch, cfg.latent_dim, 7)
    for i, s in enumerate(cfg.decoder_rates):
        ci, co = ch // 2 ** i, ch // 2 ** (i + 1)
        p = f"decoder.model.{i + 1}"
        snake(f"{p}.block.0", ci)
        convT(f"{p}.block.1", ci, co, 2 * s)
        for j in range(3):
            res_unit(f"{p}.block.{j + 2}", co)
    n = len(cfg.decoder_rates)
    snake(f"decoder.model.{n + 1}", ch // 2 ** n)
    conv(f"decoder.model.{n + 2}", 1, ch // 2 ** n, 7)
    return out


def make_state_dict(cfg: SAMAudioConfig, seed: int = 0, device="cpu",
                    include_codec: bool = True) -> Dict[str, torch.Tensor]:
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    tc = cfg.transformer
    d, hd, hid = tc.dim, tc.head_dim, tc.ffn_hidden
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias=False):
        sd[f"{name}.weight"] = _uniform(gen, (out_f, in_f), in_f, device)
        if bias:
            sd[f"{name}.bias"] = _uniform(gen, (out_f,), in_f, device)

    def norm(name, n):
        sd[name] = 1.0 + _randn(gen, (n,), device, 0.1)

    # --- top-level conditioning (reference: model.py:85-93) ---
    lin("proj", d, cfg.in_channels, bias=True)
    lin("memory_proj", d, cfg.text_encoder.dim, bias=True)
    vd = cfg.vision_encoder.dim
    sd["align_masked_video.conv.weight"] = _uniform(gen, (d, vd, 1), vd, device)
    sd["align_masked_video.conv.bias"] = _uniform(gen, (d,), vd, device)
    norm("align_masked_video.layer_norm.weight", d)
    sd["align_masked_video.layer_norm.bias"] = _randn(gen, (d,), device, 0.1)
    sd["align_masked_video.gate"] = torch.full((1,), 0.5, device=device)
    emb = _randn(gen, (cfg.num_anchors + 1, cfg.anchor_embedding_dim), device)
    emb[cfg.num_anchors] = 0.0  # padding_idx row
    sd["embed_anchors.embed.weight"] = emb
    lin("embed_anchors.proj", d, cfg.anchor_embedding_dim)
    sd["embed_anchors.gate"] = torch.full((1,), 0.5, device=device)

    # --- DiT (reference: transformer.py:394-471) ---
    T = "transformer"
    for blk in ("block1", "block2"):
        p = f"{T}.x_embedder.block.{blk}"
        norm(f"{p}.groupnorm.weight", d)
        sd[f"{p}.groupnorm.bias"] = _randn(gen, (d,), device, 0.1)
        sd[f"{p}.project.weight"] = _uniform(gen, (d, d, 3), 3 * d, device)
        sd[f"{p}.project.bias"] = _uniform(gen, (d,), 3 * d, device)
    fe = tc.frequency_embedding_dim
    lin(f"{T}.t_embedder.projection.w1", d, fe)
    lin(f"{T}.t_embedder.projection.w2", d, d)
    lin(f"{T}.t_embedder.projection.w3", d, fe)
    lin(f"{T}.t_block", 6 * d, d, bias=True)
    for w in ("w1", "w2", "w3"):
        lin(f"{T}.y_embedder.projection.{w}", d, tc.context_dim if w != "w2" else d)
    for i in range(tc.n_layers):
        L = f"{T}.layers.{i}"
        for att in ("attention", "cross_attention"):
            for w in ("wq", "wk", "wv", "wo"):
                lin(f"{L}.{att}.{w}", d, d)
            norm(f"{L}.{att}.q_norm.weight", hd)
            norm(f"{L}.{att}.k_norm.weight", hd)
        lin(f"{L}.feed_forward.w1", hid, d)
        lin(f"{L}.feed_forward.w2", d, hid)
        lin(f"{L}.feed_forward.w3", hid, d)
        norm(f"{L}.attention_norm.weight", d)
        norm(f"{L}.ffn_norm.weight", d)
        sd[f"{L}.scale_shift_table"] = _randn(gen, (6, d), device, d ** -0.5)
    norm(f"{T}.norm.weight", d)
    lin(f"{T}.output", tc.out_channels, d)
    sd[f"{T}.final_layer_scale_shift_table"] = _randn(gen, (2, d), device, d ** -0.5)

    # --- DAC-VAE codec ---
    if include_codec:
        for name, shape, kind in codec_param_shapes(cfg.audio_codec):
            full = f"audio_codec.{name}"
            if kind == "conv":
                sd[full] = _uniform(gen, shape, shape[1] * shape[2], device)
            elif kind == "convT":
                # two taps of a 2s-tap transposed conv reach each output sample
                sd[full] = _uniform(gen, shape, shape[0] * 2, device)
            elif kind == "bias":
                sd[full] = _randn(gen, shape, device, 0.02)
            else:  # snake alpha, strictly positive
                sd[full] = (1.0 + _randn(gen, shape, device, 0.1)).abs() + 0.05
    return sd


def synthetic_clip(i: int, n_samples: int = 480_000, sr: int = 48_000) -> torch.Tensor:
    """0.1 N(0,1) + 3 random sines (100-8000 Hz, amp 0.2), clipped."""
    g = torch.Generator().manual_seed(1234 + i)
    t = torch.arange(n_samples, dtype=torch.float32) / sr
    wav = 0.1 * torch.randn(n_samples, generator=g)
    for _ in range(3):
        f = 100.0 + 7900.0 * torch.rand((), generator=g).item()
        wav = wav + 0.2 * torch.sin(2 * math.pi * f * t)
    return wav.clamp_(-1, 1).unsqueeze(0)


def synthetic_descriptions(n: int) -> List[str]:
    return [DESCRIPTIONS[i % len(DESCRIPTIONS)] for i in range(n)]


def synthetic_noise(bc: int, t: int, ch: int = 256, seed: int = 4321) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(bc, t, ch, generator=g)


def synthetic_text_features(descriptions: List[str], dim: int = 768, seed: int = 99):
    """Stand-in for T5-base hidden states when no t5-base checkpoint/tokenizer is on
    disk: one deterministic N(0,1) vector per whitespace token (+1 end token),
    padded to the longest description, with the reference's bool mask convention
    (True = real token; text_encoder.py:37)."""
    toks = [d.split() + ["</s>"] for d in descriptions]
    L = max(len(t) for t in toks)
    feats = torch.zeros(len(toks), L, dim)
    mask = torch.zeros(len(toks), L, dtype=torch.bool)
    for b, ts in enumerate(toks):
        for j, tok in enumerate(ts):
            h = seed
            for c in tok:
                h = (h * 131 + ord(c)) % (2 ** 31 - 1)
            g = torch.Generator().manual_seed(h + 7919 * j)
            feats[b, j] = torch.randn(dim, generator=g)
            mask[b, j] = True
    return feats, mask
