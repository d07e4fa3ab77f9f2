"""Staged bring-up checks on a B200 (development tool; the real tests live in tests/).

    python tools/gpu_check.py --stage gemm|attn|dit|codec|separate|perf
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sam_audio_b200 import _capi  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def stage_gemm():
    L = _capi.lib()
    ok = True
    for (M, N, K, bn, bk, cg) in [(128, 256, 64, 256, 64, 1), (128, 128, 64, 128, 64, 1), (300, 256, 256, 256, 64, 1),
                                  (1000, 512, 2048, 256, 64, 1), (4096, 4096, 4096, 256, 64, 1),
                                  (300, 96, 96, 96, 32, 1), (777, 128, 160, 128, 32, 1), (500, 192, 192, 192, 64, 1),
                                  (333, 96, 192, 96, 64, 1), (129, 64, 128, 64, 64, 1),
                                  (256, 256, 64, 256, 64, 2), (300, 512, 256, 256, 64, 2), (1000, 512, 2048, 256, 64, 2),
                                  (4096, 4096, 4096, 256, 64, 2), (16000, 2816, 2816, 256, 64, 2)]:
        g = torch.Generator(device="cuda").manual_seed(M + N + K)
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
        c = torch.full((M, N), float("nan"), device="cuda")
        _capi.check(L.sab_test_gemm(M, N, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), bn, bk, cg, _capi.stream_ptr()))
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        e = rel(c, ref)
        good = e < 1e-3 and not torch.isnan(c).any()
        ok &= bool(good)
        print(f"gemm M={M} N={N} K={K} BN={bn} BK={bk} CG={cg}: rel={e:.3e} nan={int(torch.isnan(c).sum())} {'OK' if good else 'FAIL'}",
              flush=True)
        if not good:
            d = (c - ref).abs()
            print("   max abs err", float(d.max()), "at", divmod(int(d.argmax()), N), " c[0,:4]", c[0, :4].tolist(),
                  "ref", ref[0, :4].tolist(), flush=True)
    # timing of the big one
    M = N = K = 8192
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    c = torch.empty(M, N, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for cg in (1, 2):
        for _ in range(2):
            _capi.check(L.sab_test_gemm(M, N, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), 256, 64, cg, _capi.stream_ptr()))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            _capi.check(L.sab_test_gemm(M, N, K, a.data_ptr(), b.data_ptr(), c.data_ptr(), 256, 64, cg, _capi.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"gemm 8192^3 BN=256 CG={cg}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    for _ in range(2):
        torch.matmul(a, b.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        torch.matmul(a, b.t())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"cublas 8192^3: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok


def stage_attn():
    L = _capi.lib()
    ok = True
    for (items, heads, Tq, Tk, masked) in [(2, 2, 64, 64, False), (2, 3, 250, 250, True), (3, 2, 37, 5, True),
                                           (1, 2, 300, 130, True)]:
        g = torch.Generator(device="cuda").manual_seed(Tq + Tk)
        q = torch.randn(items * Tq, heads * 128, device="cuda", generator=g).bfloat16()
        k = torch.randn(items * Tk, heads * 128, device="cuda", generator=g).bfloat16()
        v = torch.randn(items * Tk, heads * 128, device="cuda", generator=g).bfloat16()
        mask = torch.ones(items, Tk, dtype=torch.uint8, device="cuda")
        if masked:
            for i in range(items):
                mask[i, max(1, Tk - 3 * (i + 1)):] = 0
        o = torch.zeros(items * Tq, heads * 128, device="cuda", dtype=torch.bfloat16)
        _capi.check(L.sab_test_attention(items, heads, Tq, Tk, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                         mask.data_ptr(), o.data_ptr(), _capi.stream_ptr()))
        torch.cuda.synchronize()
        qf = q.float().view(items, Tq, heads, 128).permute(0, 2, 1, 3)
        kf = k.float().view(items, Tk, heads, 128).permute(0, 2, 1, 3)
        vf = v.float().view(items, Tk, heads, 128).permute(0, 2, 1, 3)
        s = qf @ kf.transpose(-1, -2) / 128 ** 0.5
        s = s.masked_fill(~mask.bool()[:, None, None, :], float("-inf"))
        ref = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(items * Tq, heads * 128)
        e = rel(o.float(), ref)
        good = e < 1e-2
        ok &= good
        print(f"attn items={items} H={heads} Tq={Tq} Tk={Tk}: rel={e:.3e} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def stage_attn_tc():
    """tcgen05 self-attention vs fp32 torch; also probes the MN-major V descriptor strides."""
    L = _capi.lib()
    ok_default = True
    for (lbo, sbo) in [(0, 0), (1024, 32768), (32768, 128), (128, 1024)]:
        for (items, heads, T, masked) in [(1, 1, 256, False), (2, 3, 250, True), (3, 2, 37, True), (2, 2, 129, False)]:
            g = torch.Generator(device="cuda").manual_seed(T + items)
            q, k, v = (torch.randn(items * T, heads * 128, device="cuda", generator=g).bfloat16() for _ in range(3))
            mask = torch.ones(items, T, dtype=torch.uint8, device="cuda")
            if masked:
                for i in range(items):
                    mask[i, max(1, T - 3 * (i + 1)):] = 0
            o = torch.zeros(items * T, heads * 128, device="cuda", dtype=torch.bfloat16)
            _capi.check(L.sab_test_attention_tc(items, heads, T, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                                mask.data_ptr(), o.data_ptr(), lbo, sbo, _capi.stream_ptr()))
            torch.cuda.synchronize()
            qf, kf, vf = (x.float().view(items, T, heads, 128).permute(0, 2, 1, 3) for x in (q, k, v))
            s = (qf @ kf.transpose(-1, -2) / 128 ** 0.5).masked_fill(~mask.bool()[:, None, None, :], float("-inf"))
            ref = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(items * T, heads * 128)
            e = rel(o.float(), ref)
            good = e < 1e-2
            if (lbo, sbo) == (0, 0):
                ok_default &= good
            print(f"attn_tc lbo={lbo} sbo={sbo} items={items} H={heads} T={T}: rel={e:.3e} {'OK' if good else 'FAIL'}", flush=True)
            if not good and (lbo, sbo) == (0, 0):
                d = (o.float() - ref)
                print("   o[0,:8]", o[0, :8].float().tolist(), "\n   ref    ", ref[0, :8].tolist(), flush=True)
                print("   o[0,64:72]", o[0, 64:72].float().tolist(), "\n   ref      ", ref[0, 64:72].tolist(), flush=True)
    # timing at the bench shape
    items, heads, T = 64, 22, 250
    q, k, v = (torch.randn(items * T, heads * 128, device="cuda").bfloat16() for _ in range(3))
    o = torch.zeros_like(q)
    mask = torch.ones(items, T, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("tcgen05", lambda: L.sab_test_attention_tc(items, heads, T, q.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr(), o.data_ptr(), 0, 0, _capi.stream_ptr())),
                     ("mma.sync", lambda: L.sab_test_attention(items, heads, T, T, q.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr(), o.data_ptr(), _capi.stream_ptr()))):
        for _ in range(3):
            _capi.check(fn())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            _capi.check(fn())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"self-attention {name}: {ms * 1e3:.1f} us  {4 * items * heads * T * T * 128 / ms / 1e9:.1f} TFLOP/s", flush=True)
    return ok_default


def _tiny_model():
    from sam_audio_b200.model import build_synthetic_model
    return build_synthetic_model("sam-audio-tiny", seed=0)


def stage_dit():
    g = torch.load(os.path.join(ROOT, "tests/golden/samaudio_forward_tiny.pt"))
    m = _tiny_model()
    ok = True
    for tag, vid in (("video", g["video"]), ("novideo", None)):
        out = m.forward(g["noisy"].cuda(), g["feats"].cuda(), g["text"].cuda(), g["time"].cuda(),
                        masked_video_features=None if vid is None else vid.cuda(), text_mask=g["text_mask"].cuda(),
                        anchor_ids=g["anchor_ids"].cuda(), anchor_alignment=g["anchor_alignment"].cuda(),
                        audio_pad_mask=g["pad_mask"].cuda())
        torch.cuda.synchronize()
        e = rel(out.cpu(), g["out"][tag])
        good = e < 3e-2
        ok &= good
        print(f"SAMAudio.forward[{tag}] vs reference golden: rel={e:.3e} {'OK' if good else 'FAIL'}", flush=True)
    return ok


def stage_codec():
    from oracle import restate
    from sam_audio_b200.config import stand_in_config
    from sam_audio_b200.synthetic import make_state_dict, synthetic_clip
    cfg = stand_in_config("sam-audio-tiny")
    sd = make_state_dict(cfg, seed=0)
    m = _tiny_model()
    wav = torch.stack([synthetic_clip(i, 1920 * 13) for i in range(2)])  # [2,1,S]
    ref_feat = restate.codec_encode(sd, cfg.audio_codec, wav).transpose(1, 2)
    feats = m._get_audio_features(wav.cuda())
    torch.cuda.synchronize()
    e1 = rel(feats[:, :, :128].cpu(), ref_feat)
    e1b = rel(feats[:, :, 128:].cpu(), ref_feat)
    print(f"codec encode vs oracle: rel={e1:.3e} (dup half {e1b:.3e})", flush=True)
    lat = torch.randn(2, 13, 256, generator=torch.Generator().manual_seed(3))
    ref_wav = restate.codec_decode(sd, cfg.audio_codec, lat.transpose(1, 2).reshape(4, 128, 13)).view(2, 2, -1)
    out = torch.empty(2, 2, 13 * 1920, device="cuda")
    m._ensure_engine().decode(lat.cuda(), 2, 13, out)
    torch.cuda.synchronize()
    e2 = rel(out.cpu(), ref_wav)
    print(f"codec decode vs oracle: rel={e2:.3e}", flush=True)
    return e1 < 3e-2 and e2 < 5e-2


def stage_separate():
    from sam_audio_b200 import SAMAudioProcessor
    from sam_audio_b200.synthetic import synthetic_clip, synthetic_descriptions
    g = torch.load(os.path.join(ROOT, "tests/golden/separate_tiny.pt"))
    m = _tiny_model()
    proc = SAMAudioProcessor(1920, 48000)
    ok = True
    for cand in (1, 2):
        auds = [synthetic_clip(i, n) for i, n in enumerate(g["lens"])]
        batch = proc(descriptions=synthetic_descriptions(2), audios=auds).to("cuda")
        r = g["results"][cand]
        out = m.separate(batch, noise=r["noise"].cuda(), reranking_candidates=cand)
        torch.cuda.synchronize()
        for name, ours, ref in (("target", out.target, r["target"]), ("residual", out.residual, r["residual"])):
            for i, (a, b) in enumerate(zip(ours, ref)):
                assert a.shape == b.shape, (a.shape, b.shape)
                e = rel(a.cpu(), b)
                snr = -20 * torch.log10(torch.tensor(e)).item()
                good = snr > 20
                ok &= good
                print(f"separate cand={cand} {name}[{i}] len={a.numel()} rel={e:.3e} snr={snr:.1f} dB "
                      f"{'OK' if good else 'FAIL'}", flush=True)
    return ok


def stage_perf(name="sam-audio-base", B=8):
    from sam_audio_b200 import SAMAudioProcessor
    from sam_audio_b200.model import build_synthetic_model
    from sam_audio_b200.synthetic import synthetic_clip, synthetic_descriptions, synthetic_noise
    t0 = time.time()
    m = build_synthetic_model(name, seed=0, weights_device="cuda")
    print(f"model {name} built in {time.time() - t0:.1f}s", flush=True)
    proc = SAMAudioProcessor(1920, 48000)
    auds = [synthetic_clip(i) for i in range(B)]
    batch = proc(descriptions=synthetic_descriptions(B), audios=auds).to("cuda")
    noise = synthetic_noise(B, 250).cuda()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        out = m.separate(batch, noise=noise)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"separate B={B}: {dt * 1e3:.1f} ms  -> {B / dt:.2f} clips/s  ws={m._engine.workspace_bytes() / 2**30:.1f} GiB "
              f"finite={bool(torch.isfinite(out.target[0]).all())}", flush=True)
    # split timing
    eng = m._ensure_engine()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    feats = m._get_audio_features(batch.audios)
    ev[1].record()
    lat = torch.empty_like(noise)
    eng.solve(noise, 16, lat)
    ev[2].record()
    wav = torch.empty(B, 2, 480000, device="cuda")
    eng.decode(lat, B, 250, wav)
    ev[3].record()
    torch.cuda.synchronize()
    print(f"encode {ev[0].elapsed_time(ev[1]):.1f} ms | solve {ev[1].elapsed_time(ev[2]):.1f} ms | "
          f"decode {ev[2].elapsed_time(ev[3]):.1f} ms", flush=True)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True)
    ap.add_argument("--model", default="sam-audio-base")
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    fn = {"gemm": stage_gemm, "attn": stage_attn, "attn_tc": stage_attn_tc, "dit": stage_dit, "codec": stage_codec,
          "separate": stage_separate, "perf": lambda: stage_perf(a.model, a.batch)}[a.stage]
    ok = fn()
    print(f"STAGE {a.stage}: {'PASS' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)
