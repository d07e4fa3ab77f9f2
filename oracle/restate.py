"""TEST INFRASTRUCTURE — fp32 CPU restatement of the SAMAudio.separate() hot path.

Functional code over a state dict with the reference's parameter names.  Every
function cites the reference file:line it follows (paths relative to
/root/reference).  Written from the reference's behaviour, not copied from it;
structure differs deliberately (functional, explicit head indexing, explicit
midpoint loop) so that it is an independent check.

Pinned against the reference's own modules by oracle/make_golden.py (DiT,
SAMAudio.forward/separate control flow, processor, anchors).  The DAC-VAE codec
part restates the Descript-DAC layout that ``dacvae`` derives from —
PARITY UNPINNED (dacvae source is absent from the reference tree).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# processor (integer work; must be bit-exact)
# ----------------------------------------------------------------------------
def batch_audio(audios: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """processor.py:23-36 — mono mix by channel mean, zero-pad to longest."""
    monos = [a.mean(0) for a in audios]
    n = max(m.numel() for m in monos)
    out = torch.zeros(len(monos), 1, n, dtype=monos[0].dtype)
    for i, m in enumerate(monos):
        out[i, 0, : m.numel()] = m
    return out, torch.tensor([m.numel() for m in monos])


def wav_to_feature_idx(wav_sizes: torch.Tensor, hop: int) -> torch.Tensor:
    """processor.py:190-195 — torch.ceil(int64 / int) is a float32 tensor."""
    return torch.ceil(wav_sizes / hop)


def mask_from_sizes(sizes: torch.Tensor) -> torch.Tensor:
    """processor.py:127-128."""
    return torch.arange(int(sizes.max()))[None, :] < sizes[:, None]


ANCHOR_VOCAB = {"<null>": 0, "+": 1, "-": 2, "<pad>": 3}


def process_anchors(anchors, pad_mask: torch.Tensor, hop: int, sr: int):
    """processor.py:78-124 — ids / alignment; python-double ceil; later anchors win."""
    B, T = pad_mask.shape
    align = torch.zeros(B, T, dtype=torch.long)
    align[~pad_mask] = 1
    if anchors is None:
        ids = torch.zeros(B, 2, dtype=torch.long)
        ids[:, 1] = ANCHOR_VOCAB["<pad>"]
        return ids, align
    rows = []
    for i, lst in enumerate(anchors):
        cur = [ANCHOR_VOCAB["<null>"], ANCHOR_VOCAB["<pad>"]]
        for tok, t0, t1 in lst:
            s = math.ceil(t0 * sr / hop)
            e = math.ceil(t1 * sr / hop)
            align[i, s:e] = len(cur)
            cur.append(ANCHOR_VOCAB[tok])
        rows.append(cur)
    K = max(len(r) for r in rows)
    ids = torch.full((B, K), ANCHOR_VOCAB["<pad>"], dtype=torch.long)
    for i, r in enumerate(rows):
        ids[i, : len(r)] = torch.tensor(r)
    return ids, align


# ----------------------------------------------------------------------------
# DiT building blocks
# ----------------------------------------------------------------------------
def rmsnorm(x, w, eps=1e-5):
    """transformer.py:42-47 (fp32)."""
    x = x.float()
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def swiglu_proj(sd: SD, p: str, x):
    """transformer.py:72-80 ProjectionLayer with swiglu: w2(silu(w1 x) * w3 x)."""
    return F.linear(F.silu(F.linear(x, sd[f"{p}.w1.weight"])) * F.linear(x, sd[f"{p}.w3.weight"]),
                    sd[f"{p}.w2.weight"])


def timestep_embedding(t, dim=256, max_period=10000.0):
    """transformer.py:228-253 — cat(cos, sin) of t * exp(-ln(P) i/half), raw t."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    a = t[:, None].float() * freqs[None]
    return torch.cat([a.cos(), a.sin()], -1)


def sinusoidal_embedding(t, dim, theta=10000.0):
    """model.py:25-42 with pos=time (model.py:170)."""
    half = dim // 2
    inv = torch.exp(-math.log(theta) * torch.arange(half).float() / half)
    e = t[:, None] * inv[None]
    return torch.cat([e.cos(), e.sin()], -1)


def rope_angles(T: int, hd: int, theta: float):
    """rope.py:116-145 — angle[pos, i] = pos * theta^(-2i/hd)."""
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2)[: hd // 2].float() / hd))
    ang = torch.outer(torch.arange(T), inv).float()
    return ang.cos(), ang.sin()


def apply_rope(x, cos, sin):
    """rope.py:147-155 — rotate adjacent pairs (2i,2i+1); x is [B,H,T,hd]."""
    x0, x1 = x[..., 0::2], x[..., 1::2]
    o0 = x0 * cos - x1 * sin
    o1 = x0 * sin + x1 * cos
    return torch.stack([o0, o1], -1).flatten(-2)


def split_heads_interleaved(x, H: int):
    """transformer.py:121-126 — channel c = d*H + h belongs to head h."""
    B, T, C = x.shape
    return x.reshape(B, T, C // H, H).permute(0, 3, 1, 2)


def attention(sd: SD, p: str, x, H: int, theta: float, cross_x=None, key_mask=None,
              use_rope=False, eps=1e-5):
    """transformer.py:128-161.  key_mask: bool [B,S], True = attend."""
    q = F.linear(x, sd[f"{p}.wq.weight"])
    src = x if cross_x is None else cross_x
    k = F.linear(src, sd[f"{p}.wk.weight"])
    v = F.linear(src, sd[f"{p}.wv.weight"])
    q, k, v = (split_heads_interleaved(z, H) for z in (q, k, v))
    q = rmsnorm(q, sd[f"{p}.q_norm.weight"], eps)
    k = rmsnorm(k, sd[f"{p}.k_norm.weight"], eps)
    hd = q.shape[-1]
    if use_rope:
        cos, sin = rope_angles(q.shape[2], hd, theta)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
    o = torch.softmax(s, -1) @ v                       # [B,H,T,hd]
    o = o.permute(0, 2, 1, 3).reshape(x.shape[0], x.shape[1], -1)  # (h d)
    return F.linear(o, sd[f"{p}.wo.weight"])


def dit_block(sd: SD, p: str, x, y, t0, H, theta, pad_mask, mem_mask, eps=1e-5):
    """transformer.py:354-391."""
    B, _, d = x.shape
    mod = sd[f"{p}.scale_shift_table"][None] + t0.reshape(B, 6, d)
    sh1, sc1, g1, sh2, sc2, g2 = (mod[:, i:i + 1] for i in range(6))
    a = attention(sd, f"{p}.attention",
                  rmsnorm(x, sd[f"{p}.attention_norm.weight"], eps) * (1 + sc1) + sh1,
                  H, theta, key_mask=pad_mask, use_rope=True, eps=eps)
    h = x + a * g1
    h = h + attention(sd, f"{p}.cross_attention", h, H, theta, cross_x=y, key_mask=mem_mask, eps=eps)
    z = rmsnorm(h, sd[f"{p}.ffn_norm.weight"], eps) * (1 + sc2) + sh2
    ff = F.linear(F.silu(F.linear(z, sd[f"{p}.feed_forward.w1.weight"])) *
                  F.linear(z, sd[f"{p}.feed_forward.w3.weight"]), sd[f"{p}.feed_forward.w2.weight"])
    return h + ff * g2


def x_embedder(sd: SD, p: str, x_btc):
    """patcher.py:138-164 with patch_size=1, num_groups=1: GN(1)->SiLU->conv k3 (zero pad 1,1), twice, + x."""
    x = x_btc.transpose(1, 2)  # [B,C,T]
    h = x
    for blk in ("block1", "block2"):
        q = f"{p}.block.{blk}"
        h = F.group_norm(h, 1, sd[f"{q}.groupnorm.weight"], sd[f"{q}.groupnorm.bias"], eps=1e-5)
        h = F.conv1d(F.pad(F.silu(h), (1, 1)), sd[f"{q}.project.weight"], sd[f"{q}.project.bias"])
    return (h + x).transpose(1, 2)


def dit_forward(sd: SD, cfg, x, time, pad_mask, memory, mem_mask, return_layers=False):
    """transformer.py:473-524.  x [B,T,d], time [B], memory [B,L,d]."""
    p = "transformer"
    H, theta, eps = cfg.n_heads, cfg.rope_theta, cfg.norm_eps
    h = x_embedder(sd, f"{p}.x_embedder", x)
    t = swiglu_proj(sd, f"{p}.t_embedder.projection", timestep_embedding(time, cfg.frequency_embedding_dim))
    t0 = F.linear(F.silu(t), sd[f"{p}.t_block.weight"], sd[f"{p}.t_block.bias"])
    y = swiglu_proj(sd, f"{p}.y_embedder.projection", memory)
    layers = [h]
    for i in range(cfg.n_layers):
        h = dit_block(sd, f"{p}.layers.{i}", h, y, t0, H, theta, pad_mask, mem_mask, eps)
        layers.append(h)
    fin = sd[f"{p}.final_layer_scale_shift_table"][None] + t[:, None]
    shift, scale = fin[:, 0:1], fin[:, 1:2]
    h = rmsnorm(h, sd[f"{p}.norm.weight"], eps) * (1 + scale) + shift
    out = F.linear(h, sd[f"{p}.output.weight"])
    return (out, layers) if return_layers else out


# ----------------------------------------------------------------------------
# SAMAudio.forward (one ODE function evaluation)
# ----------------------------------------------------------------------------
def conditioning(sd: SD, audio_features, video_features, anchor_ids, anchor_alignment):
    """Time-independent part of align_inputs (model.py:108-128, align.py:30-50,
    model.py:54-65): everything except the noisy-audio third of ``proj``."""
    W, b = sd["proj.weight"], sd["proj.bias"]
    c2 = audio_features.shape[-1]
    x = F.linear(audio_features, W[:, 2 * c2:3 * c2], b)      # middle third multiplies zeros
    if video_features is not None:                             # align.py:41-42: None -> input unchanged
        pc = F.conv1d(video_features, sd["align_masked_video.conv.weight"], sd["align_masked_video.conv.bias"])
        pc = F.layer_norm(pc.permute(0, 2, 1), (W.shape[0],), sd["align_masked_video.layer_norm.weight"],
                          sd["align_masked_video.layer_norm.bias"], eps=1e-5)
        x = x + torch.tanh(sd["align_masked_video.gate"]) * pc
    if anchor_ids is not None:                                 # model.py:57-58: None -> input unchanged
        emb = sd["embed_anchors.embed.weight"][anchor_ids.gather(1, anchor_alignment)]
        x = x + torch.tanh(sd["embed_anchors.gate"]) * F.linear(emb, sd["embed_anchors.proj.weight"])
    return x


def samaudio_forward(sd: SD, cfg, noisy_audio, audio_features, text_features, time,
                     masked_video_features, text_mask, anchor_ids, anchor_alignment, audio_pad_mask):
    """model.py:130-180."""
    c2 = audio_features.shape[-1]
    x = F.linear(noisy_audio, sd["proj.weight"][:, :c2]) + conditioning(
        sd, audio_features, masked_video_features, anchor_ids, anchor_alignment)
    temb = sinusoidal_embedding(time, cfg.transformer.dim)[:, None]
    memory = temb                                              # model.py:170-172: text None -> time-only memory
    if text_features is not None:
        memory = F.linear(text_features, sd["memory_proj.weight"], sd["memory_proj.bias"]) + temb
    return dit_forward(sd, cfg.transformer, x, time, audio_pad_mask, memory, text_mask)


def odeint_midpoint(f: Callable, y0, n_steps: int = 16):
    """model.py:22,285-290 + torchdiffeq fixed-grid midpoint: t_k = k/n, per step
    f0=f(t,y); y += dt * f(t+dt/2, y + f0*dt/2).  32 evaluations for n=16."""
    y, dt = y0, 1.0 / n_steps
    for k in range(n_steps):
        t0 = torch.tensor(k * dt)
        f0 = f(t0, y)
        y = y + dt * f(t0 + dt / 2, y + f0 * (dt / 2))
    return y


def odeint_fixed(f: Callable, y0, n_steps: int = 16, method: str = "midpoint"):
    """The fixed-grid solvers the reference can select through **ode_opt (model.py:285-290 -> torchdiffeq.odeint).
    torchdiffeq's source is absent here (PARITY UNPINNED: restated from its published fixed_grid.py / rk_common.py):
    euler  y += dt f(t, y);  midpoint as above;  rk4 = the 3/8 rule (rk4_alt_step_func)."""
    if method == "midpoint":
        return odeint_midpoint(f, y0, n_steps)
    y, dt = y0, 1.0 / n_steps
    for k in range(n_steps):
        t0 = torch.tensor(k * dt)
        if method == "euler":
            y = y + dt * f(t0, y)
        elif method == "rk4":
            k1 = f(t0, y)
            k2 = f(t0 + dt / 3, y + dt * k1 / 3)
            k3 = f(t0 + dt * 2 / 3, y + dt * (k2 - k1 / 3))
            k4 = f(t0 + dt, y + dt * (k1 - k2 + k3))
            y = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        else:
            raise ValueError(method)
    return y


# ----------------------------------------------------------------------------
# DAC-VAE codec (PARITY UNPINNED: restated from the Descript-DAC layout)
# ----------------------------------------------------------------------------
def snake(x, alpha):
    """Snake1d: x + sin^2(alpha x) / (alpha + 1e-9), alpha [1,C,1]."""
    return x + torch.sin(alpha * x).pow(2) / (alpha + 1e-9)


def _conv(sd, p, x, **kw):
    return F.conv1d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], **kw)


def _res_unit(sd, p, x, dil):
    y = _conv(sd, f"{p}.block.1", snake(x, sd[f"{p}.block.0.alpha"]), dilation=dil, padding=3 * dil)
    y = _conv(sd, f"{p}.block.3", snake(y, sd[f"{p}.block.2.alpha"]))
    return x + y


def codec_pad(wav, hop):
    """codec.py:72-78 — reflect-pad on the right to a multiple of hop."""
    n = wav.shape[-1]
    return F.pad(wav, (0, hop - n % hop), mode="reflect") if n % hop else wav


def codec_encode(sd: SD, ccfg, wav, prefix="audio_codec"):
    """codec.py:65-70: encoder -> quantizer.in_proj -> chunk(2) -> mean.  [B,1,S] -> [B,128,T]."""
    e = f"{prefix}.encoder"
    x = _conv(sd, f"{e}.block.0", codec_pad(wav, ccfg.hop_length), padding=3)
    for i, s in enumerate(ccfg.encoder_rates):
        b = f"{e}.block.{i + 1}"
        for j, dil in enumerate((1, 3, 9)):
            x = _res_unit(sd, f"{b}.block.{j}", x, dil)
        x = _conv(sd, f"{b}.block.4", snake(x, sd[f"{b}.block.3.alpha"]), stride=s, padding=math.ceil(s / 2))
    n = len(ccfg.encoder_rates)
    x = _conv(sd, f"{e}.block.{n + 2}", snake(x, sd[f"{e}.block.{n + 1}.alpha"]), padding=1)
    z = _conv(sd, f"{prefix}.quantizer.in_proj", x)
    return z[:, : ccfg.codebook_dim]


def codec_decode(sd: SD, ccfg, z, prefix="audio_codec"):
    """codec.py:86-89: quantizer.out_proj -> decoder.  [B,128,T] -> [B,1,T*hop]."""
    d = f"{prefix}.decoder"
    x = _conv(sd, f"{prefix}.quantizer.out_proj", z)
    x = _conv(sd, f"{d}.model.0", x, padding=3)
    for i, s in enumerate(ccfg.decoder_rates):
        b = f"{d}.model.{i + 1}"
        x = F.conv_transpose1d(snake(x, sd[f"{b}.block.0.alpha"]), sd[f"{b}.block.1.weight"],
                               sd[f"{b}.block.1.bias"], stride=s, padding=math.ceil(s / 2))
        for j, dil in enumerate((1, 3, 9)):
            x = _res_unit(sd, f"{b}.block.{j + 2}", x, dil)
    n = len(ccfg.decoder_rates)
    x = _conv(sd, f"{d}.model.{n + 2}", snake(x, sd[f"{d}.model.{n + 1}.alpha"]), padding=3)
    return torch.tanh(x)


# ----------------------------------------------------------------------------
# visual prompting: frame transform + chunked encoding (vision_encoder.py:47-69, 91-113)
# ----------------------------------------------------------------------------
def _aa_cubic_taps(in_size: int, out_size: int):
    """Tap window and normalised weights per output index of torch's antialiased bicubic (the kernel torchvision's
    Resize(..., BICUBIC, antialias=True) reaches through F.interpolate): PIL-style cubic a = -1/2, support and filter
    stretched by scale = in/out when down-sampling; all arithmetic in float32 (restated from the published ATen
    algorithm, aten/native/cpu/UpSampleKernel.cpp; pinned against torchvision by tests/test_oracle_vision.py)."""
    import numpy as np
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = f(2.0) * scale if scale >= 1 else f(2.0)
    invscale = f(1.0) / scale if scale >= 1 else f(1.0)
    a = f(-0.5)
    out = []
    for i in range(out_size):
        center = scale * f(i + 0.5)
        lo = max(int(center - support + f(0.5)), 0)
        n = min(int(center + support + f(0.5)), in_size) - lo
        ws, tot = [], f(0)
        for j in range(n):
            x = f(abs(f(j + lo) - center + f(0.5)) * invscale)
            if x < 1:
                w = ((a + f(2)) * x - (a + f(3))) * x * x + f(1)
            elif x < 2:
                w = (((x - f(5)) * x + f(8)) * x - f(4)) * a
            else:
                w = f(0)
            ws.append(f(w))
            tot = f(tot + f(w))
        out.append((lo, [f(w / tot) for w in ws]))
    return out


def _resize_axis(x: torch.Tensor, out_size: int, axis: int) -> torch.Tensor:
    x = x.movedim(axis, -1)
    res = torch.empty(*x.shape[:-1], out_size, dtype=torch.float32)
    for i, (lo, ws) in enumerate(_aa_cubic_taps(x.shape[-1], out_size)):
        t = x[..., lo] * float(ws[0])
        for j in range(1, len(ws)):
            t = t + x[..., lo + j] * float(ws[j])           # separate multiply / add, in tap order
        res[..., i] = t
    return res.movedim(-1, axis)


def frame_transform(video_u8: torch.Tensor, size: int = 336) -> torch.Tensor:
    """vision_encoder.py:91-113: Resize((size, size), BICUBIC) [antialias, uint8 in -> uint8 out], x/255,
    Normalize(.5, .5).  [T, 3, H, W] uint8 -> [T, 3, size, size] float32.  Width first, then height."""
    x = _resize_axis(_resize_axis(video_u8.float(), size, 3), size, 2)
    x = x.clamp(0, 255).round()                             # torchvision: clamp, round, cast to uint8
    return (x / 255.0 - 0.5) / 0.5


def vision_encode(videos, encode: Callable, size: int = 336, batch_size: int = 300, transform=frame_transform):
    """vision_encoder.py:47-69: per video transform, encode in chunks of batch_size frames, zero-pad to the longest."""
    feats = []
    for v in videos:
        x = transform(v, size)
        if batch_size > 0 and x.shape[0] > batch_size:
            feats.append(torch.cat([encode(x[i:i + batch_size]) for i in range(0, x.shape[0], batch_size)], 0))
        else:
            feats.append(encode(x))
    T = max(f.shape[0] for f in feats)
    out = feats[0].new_zeros(len(feats), T, feats[0].shape[-1])
    for i, f in enumerate(feats):
        out[i, : f.shape[0]] = f
    return out


# ----------------------------------------------------------------------------
# separate()
# ----------------------------------------------------------------------------
def separate(sd: SD, cfg, audios, pad_mask, sizes, text_features, text_mask, anchor_ids,
             anchor_alignment, noise, video_features=None, candidates: int = 1, n_steps: int = 16,
             return_latent=False, method: str = "midpoint", ranker_scores=None):
    """model.py:247-338 with rankers None (candidate 0; config.py:214-215, model.py:329-330).
    audios [B,1,S] fp32; returns (target list, residual list[, latent])."""
    cc = cfg.audio_codec
    feats = codec_encode(sd, cc, audios).transpose(1, 2)
    feats = torch.cat([feats, feats], 2)                              # model.py:183-184
    B, T, _ = feats.shape
    if video_features is None:
        video_features = feats.new_zeros(B, cfg.vision_encoder.dim, T)  # model.py:188-189

    def rep(x):                                                        # model.py:193-203
        return x if candidates == 1 else x.repeat_interleave(candidates, 0)

    fa = dict(audio_features=rep(feats), text_features=rep(text_features), text_mask=rep(text_mask),
              masked_video_features=rep(video_features), anchor_ids=rep(anchor_ids),
              anchor_alignment=rep(anchor_alignment), audio_pad_mask=rep(pad_mask))

    def field(t, y):
        return samaudio_forward(sd, cfg, y, time=t.expand(y.shape[0]), **fa)

    lat = odeint_fixed(field, noise, n_steps, method)
    Bc = lat.shape[0]
    wavs = codec_decode(sd, cc, lat.transpose(1, 2).reshape(2 * Bc, cc.codebook_dim, T)).view(Bc, 2, -1)
    n = (sizes * cc.hop_length).int()                                  # codec.py:91-97
    # model.py:306-330: arg-max over a ranker's [B, candidates] scores, candidate 0 without a ranker
    pick = [0] * B if ranker_scores is None else [int(i) for i in ranker_scores.argmax(dim=1)]
    tgt = [wavs[b * candidates + pick[b], 0, : int(n[b])] for b in range(B)]
    res = [wavs[b * candidates + pick[b], 1, : int(n[b])] for b in range(B)]
    return (tgt, res, lat) if return_latent else (tgt, res)
