"""CPU, world_size 2 over gloo: the N>1 plumbing (weight broadcast, clip sharding, waveform all-gather)
with a stand-in separator (the CUDA engine is not needed for the host-side logic)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sam_audio_b200 import SAMAudioProcessor
from sam_audio_b200.parallel import all_gather_waveforms, broadcast_state_dict, separate_sharded, shard_range


class _FakeResult:
    def __init__(self, t, r):
        self.target, self.residual = t, r


class _FakeModel:
    """separate(): target = 2 * padded clip, residual = -clip (pure function of the clip)."""

    def device(self):
        return torch.device("cpu")

    def separate(self, batch, noise=None, reranking_candidates=1):
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
