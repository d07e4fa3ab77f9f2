"""Timeline of the tcgen05 self-attention kernel's hand-over points for CTA 0 (first 8 work items), from the
clock64() stamps attention_tc2_kernel writes when given a trace buffer.

    python tools/attn_trace.py [--exact] [--poly 3]

Prints, per item, SM clocks relative to the first stamp: MMA issue points, loader, storer, and per softmax warp
(tile, lane quarter, key half) S-ready / P-written / O-ready / group-synced / O-read / staging-free / staged."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=64)
    ap.add_argument("--heads", type=int, default=22)
    ap.add_argument("--T", type=int, default=250)
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--folded", action="store_true")
    ap.add_argument("--poly", type=int, default=3)
    a = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    from sam_audio_b200 import _capi
    lib = _capi.lib()
    gen = torch.Generator(device="cuda").manual_seed(0)
    n = a.items * a.T
    q, k, v = (torch.randn(n, a.heads * 128, device="cuda", generator=gen).bfloat16() for _ in range(3))
    mask = torch.ones(a.items, a.T, dtype=torch.uint8, device="cuda")
    o = torch.zeros(n, a.heads * 128, device="cuda", dtype=torch.bfloat16)
    qn = q.float().view(-1, a.heads, 128).norm(dim=-1).max()
    kn = k.float().view(-1, a.heads, 128).norm(dim=-1).max()
    shift = -1.0 if a.exact else float(qn * kn / 128 ** 0.5 * 1.4426950408889634) + 0.25
    if a.folded:
        q = (q.float() * (1.4426950408889634 / 128 ** 0.5)).bfloat16()
        shift = 0.0
    trace = torch.zeros(8 * 32 * 8, dtype=torch.int64, device="cuda")

    def run(tr):
        _capi.check(lib.sab_test_attention_tc2(a.items, a.heads, a.T, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                               mask.data_ptr(), o.data_ptr(), shift, a.poly,
                                               None if tr is None else tr.data_ptr(), _capi.stream_ptr()))
    for _ in range(3):
        run(None)
    run(trace)
    torch.cuda.synchronize()
    t = trace.cpu().view(8, 32, 8)
    t0 = int(t[t > 0].min())
    rel = lambda x: "      -" if x == 0 else f"{int(x) - t0:7d}"
    mma = ["S0free", "QK0", "S1free", "QK1", "Vin", "PV0", "PV1"]
    sm = ["Sready", "Pdone", "Oready", "synced", "Oread", "stgfree", "staged"]
    for j in range(8):
        if not (t[j] > 0).any():
            break
        print(f"--- item {j}")
        print("  mma    " + " ".join(f"{n}={rel(t[j, 0, i])}" for i, n in enumerate(mma)))
        print("  loader " + f"QKload={rel(t[j, 1, 0])} Vload={rel(t[j, 1, 1])}")
        print("  storer " + " ".join(f"{n}={rel(t[j, 2, i])}" for i, n in enumerate(["full0", "read0", "full1", "read1"])))
        for w in range(3, 19):
            m, h, qq = (w - 3) >> 3, ((w - 3) >> 2) & 1, w & 3
            print(f"  w{w:02d} t{m} q{qq} h{h} " + " ".join(f"{n}={rel(t[j, w, i])}" for i, n in enumerate(sm)))


if __name__ == "__main__":
    main()
