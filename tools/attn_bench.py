"""Isolated timing of the tcgen05 self-attention kernels at the benchmarked shape (64 clips x 22 heads x T = 250).

    python tools/attn_bench.py [--items 64 --heads 22 --T 250 --iters 30]

CUDA events around `iters` back-to-back launches (q, k, v = 3 x 90 MB > L2, so every launch streams from HBM).
Prints one line per variant: us/launch, TFLOP/s (4 T^2 128 per (item, head)), clocks per work item at the sampled
SM clock, and the error against fp32 torch.  Variants: v1 = attention_tc.cuh; v2 = attention_tc2.cuh, exact / single
pass, polynomial share 0..4 of every 8 exponential pairs."""
import argparse
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sm_clock():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True).stdout
        return float(out.strip().splitlines()[0])
    except Exception:
        return float("nan")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=64)
    ap.add_argument("--heads", type=int, default=22)
    ap.add_argument("--T", type=int, default=250)
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    from sam_audio_b200 import _capi
    lib = _capi.lib()
    gen = torch.Generator(device="cuda").manual_seed(0)
    n = a.items * a.T
    # unit-RMS rows like QK-normalised vectors
    q, k, v = (torch.randn(n, a.heads * 128, device="cuda", generator=gen).bfloat16() for _ in range(3))
    mask = torch.ones(a.items, a.T, dtype=torch.uint8, device="cuda")
    o = torch.zeros(n, a.heads * 128, device="cuda", dtype=torch.bfloat16)
    qf, kf, vf = (x.float().view(a.items, a.T, a.heads, 128).permute(0, 2, 1, 3)[:2] for x in (q, k, v))
    ref = (torch.softmax(qf @ kf.transpose(-1, -2) / 128 ** 0.5, -1) @ vf).permute(0, 2, 1, 3).reshape(2 * a.T, -1)
    qn = q.float().view(-1, a.heads, 128).norm(dim=-1).max()
    kn = k.float().view(-1, a.heads, 128).norm(dim=-1).max()
    shift = float(qn * kn / 128 ** 0.5 * 1.4426950408889634) + 0.25
    flops = 4.0 * a.items * a.heads * a.T * a.T * 128
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    per_cta = -(-(a.items * a.heads) // n_sm)

    def run(name, fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        mhz = sm_clock()
        us = e0.elapsed_time(e1) * 1e3 / a.iters
        err = float((o[: 2 * a.T].float() - ref).norm() / ref.norm())
        print(f"{name:28s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s  ~{us * mhz / per_cta:7.0f} clk/work item "
              f"@{mhz:.0f} MHz  rel-L2 {err:.2e}", flush=True)

    st = _capi.stream_ptr
    run("v1 (attention_tc.cuh)", lambda: _capi.check(lib.sab_test_attention_tc(
        a.items, a.heads, a.T, q.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr(), o.data_ptr(), 0, 0, st())))
    qs = (q.float() * (1.4426950408889634 / 128 ** 0.5)).bfloat16()     # folded variant: scale already in q
    for name, sh, qq in (("exact", -1.0, q), ("single-pass", shift, q), ("folded", 0.0, qs)):
        for poly in (0, 2, 3, 4):
            run(f"v2 {name} poly={poly}/8", lambda: _capi.check(lib.sab_test_attention_tc2(
                a.items, a.heads, a.T, qq.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr(), o.data_ptr(),
                sh, poly, None, st())))


if __name__ == "__main__":
    main()
