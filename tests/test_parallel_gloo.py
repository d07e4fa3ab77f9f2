"""CPU, world_size 2 over gloo: the N>1 plumbing (weight broadcast, clip sharding, waveform all-gather)
with a stand-in separator (the CUDA engine is not needed for the host-side logic)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sam_audio_b200 import SAMAudioProcessor
from sam_audio_b200.parallel import (all_gather_waveforms, broadcast_state_dict, separate_and_gather, separate_sharded,
                                     shard_range)


class _FakeResult:
    def __init__(self, t, r):
        self.target, self.residual = t, r


class _FakeModel:
    """separate(): target = 2 * padded clip, residual = -clip (pure function of the clip)."""

    def device(self):
        return torch.device("cpu")

    def separate(self, batch, noise=None, reranking_candidates=1, _on_decoded=None):
        if _on_decoded is not None:          # chunked decode hook: 2 clips per chunk, c candidates per clip
            c, B = reranking_candidates, batch.audios.shape[0]
            wavs = torch.zeros(B * c, 2, 6)
            for i0 in range(0, B * c, 2 * c):
                i1 = min(B * c, i0 + 2 * c)
                for i in range(i0, i1):      # waveform value = 100*rank + clip index, + 0.5 for candidates > 0
                    wavs[i] = 100.0 * batch.rank + i // c + (0.5 if i % c else 0.0)
                _on_decoded(i0, i1, wavs)
            return None
        hop = batch.hop_length
        n = (batch.sizes * hop).int()
        S = int(batch.sizes.max()) * hop
        a = torch.nn.functional.pad(batch.audios[:, 0], (0, S - batch.audios.shape[-1]))
        return _FakeResult([2 * a[i, : int(n[i])] for i in range(len(n))], [-a[i, : int(n[i])] for i in range(len(n))])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = {"a.weight": torch.arange(12.0).view(3, 4), "b": torch.tensor([1.5])} if rank == 0 else None
        got = broadcast_state_dict(sd, src=0)
        ok = torch.equal(got["a.weight"], torch.arange(12.0).view(3, 4)) and float(got["b"]) == 1.5
        lens = [4000, 1920, 2500]
        auds = [torch.full((1, n), float(i + 1)) for i, n in enumerate(lens)]
        proc = SAMAudioProcessor(1920, 48000)
        tgt, res = separate_sharded(_FakeModel(), proc, ["x"] * 3, auds)
        for i, n in enumerate(lens):
            S = -(-n // 1920) * 1920
            exp = torch.zeros(S)
            exp[:n] = float(i + 1)
            ok &= tgt[i].shape == (S,) and torch.equal(tgt[i], 2 * exp) and torch.equal(res[i], -exp)
        # overlapped gather: every chunk of decoded clips is all-gathered while "decoding" goes on; candidate 0 only
        for cand in (1, 3):
            fb = type("B", (), {})()
            fb.audios, fb.rank = torch.zeros(5, 1, 8), rank
            full = separate_and_gather(_FakeModel(), fb, None, [5, 5], reranking_candidates=cand)
            exp = torch.tensor([100.0 * r + i for r in range(world) for i in range(5)])
            ok &= full.shape == (10, 2, 6) and torch.equal(full[:, 0, 0], exp) and torch.equal(full[:, 1, 5], exp)
        lo, hi = shard_range(3, rank, world)
        local = torch.full((hi - lo, 2, 5), float(rank))
        g = all_gather_waveforms(local, [2, 1])
        ok &= g.shape == (3, 2, 5) and g[:2].eq(0).all().item() and g[2].eq(1).all().item()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res == [(0, True), (1, True)]
