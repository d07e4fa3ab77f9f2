"""Data-parallel sharding of separate() over the GPUs of one box.

The path is embarrassingly parallel over clips (the reference's only
parallelism is a DistributedSampler over clips with a full model replica per rank,
eval/main.py:53-76).  One process per GPU over ``torch.distributed``:

* ``broadcast_state_dict`` — one broadcast of the (flattened) weights from rank 0
  (NCCL over NVLink on GPUs, gloo in the CPU tests);
* ``shard_range`` — contiguous split of the clips; a clip's candidates stay on one rank;
* ``all_gather_waveforms`` — one all-gather of the separated waveforms.

No collective sits inside the model: the ODE loop never communicates.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank's items; the first (n % world) ranks take one extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_state_dict(sd: Dict[str, torch.Tensor] | None, src: int = 0, device="cpu",
                         group=None) -> Dict[str, torch.Tensor]:
    """Rank ``src`` holds ``sd``; every rank returns an identical fp32 copy on ``device``.
    Metadata travels as one object broadcast, the payload as ONE flat tensor broadcast."""
    rank = dist.get_rank(group)
    meta = [None]
    if rank == src:
        assert sd is not None
        meta[0] = [(k, tuple(v.shape)) for k, v in sd.items()]
    dist.broadcast_object_list(meta, src=src, group=group)
    total = sum(int(torch.Size(s).numel()) for _, s in meta[0])
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if rank == src:
        off = 0
        for k, s in meta[0]:
            n = int(torch.Size(s).numel())
            flat[off:off + n] = sd[k].reshape(-1).to(device=device, dtype=torch.float32)
            off += n
    dist.broadcast(flat, src=src, group=group)
    out, off = {}, 0
    for k, s in meta[0]:
        n = int(torch.Size(s).numel())
        out[k] = flat[off:off + n].view(s)
        off += n
    return out


def all_gather_waveforms(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """local [n_local, 2, S] -> [sum(counts), 2, S] on every rank (one all-gather; ragged shards are
    padded to the largest and trimmed)."""
    world = dist.get_world_size(group)
    n_max = max(counts)
    if local.shape[0] < n_max:
        pad = local.new_zeros(n_max - local.shape[0], *local.shape[1:])
        local = torch.cat([local, pad], 0)
    buf = local.new_empty(world * n_max, *local.shape[1:])
    dist.all_gather_into_tensor(buf, local.contiguous(), group=group)
    if all(c == n_max for c in counts):
        return buf
    parts = [buf[r * n_max: r * n_max + counts[r]] for r in range(world)]
    return torch.cat(parts, 0)


def separate_and_gather(model, batch, noise, counts: Sequence[int], reranking_candidates: int = 1, group=None,
                        **separate_kwargs) -> torch.Tensor:
    """separate() of this rank's shard with the waveform all-gather OVERLAPPED with the decode: the codec decodes a few
    clips at a time and each finished chunk is all-gathered (asynchronously, on the collective's own stream) while the
    next chunk decodes.  Every rank must hold the same number of clips (`counts` all equal; use separate_sharded
    otherwise).  Returns [sum(counts), 2, S] (candidate 0 of every clip, rank-major) on every rank.

    Buffers are kept on the model and reused: fresh ~GB allocations per call would be recorded on the collective's
    stream and force the caching allocator into cudaMalloc / cudaFree (device-synchronising) every step — measured: 1.5x
    step time at N=2.  Each chunk is gathered with all_gather_into_tensor straight into its slice of a staging buffer
    laid out [chunk][rank][clip] (no temporary), and one strided copy per chunk puts it in rank-major order at the end."""
    world = dist.get_world_size(group)
    B = counts[0]
    assert all(n == B for n in counts) and len(counts) == world, "separate_and_gather needs equal shards"
    c = int(reranking_candidates)
    state = {"stage": None, "out": None, "works": [], "chunks": [], "keep": []}

    def buffers(wavs):
        S = wavs.shape[-1]
        cache = getattr(model, "_gather_bufs", None)
        if cache is None or cache[0].numel() != world * B * 2 * S or cache[0].device != wavs.device:
            cache = (wavs.new_empty(world * B * 2 * S), wavs.new_empty(world * B, 2, S))
            try:
                model._gather_bufs = cache
            except Exception:
                pass
        return cache

    def on_decoded(i0, i1, wavs):
        if state["stage"] is None:
            state["stage"], state["out"] = buffers(wavs)
        S = wavs.shape[-1]
        b0, b1 = i0 // c, i1 // c                                   # clips of this chunk; their candidate-0 waveforms
        local = wavs[i0:i1:c].contiguous() if c > 1 else wavs[i0:i1]
        n = (b1 - b0) * 2 * S
        flat = state["stage"][world * b0 * 2 * S: world * b0 * 2 * S + world * n]
        state["keep"].append(local)
        state["chunks"].append((b0, b1, flat.view(world, b1 - b0, 2, S)))
        # concatenated form ([world * rows, 2, S]): the one both NCCL and gloo accept
        state["works"].append(dist.all_gather_into_tensor(flat.view(world * (b1 - b0), 2, S), local, group=group,
                                                          async_op=True))

    model.separate(batch, noise=noise, reranking_candidates=c, _on_decoded=on_decoded, **separate_kwargs)
    with torch.inference_mode():      # the buffers were created inside separate()'s inference_mode
        out = state["out"].view(world, B, 2, -1)
        for w, (b0, b1, dst) in zip(state["works"], state["chunks"]):
            w.wait()
            out[:, b0:b1].copy_(dst)
    return state["out"]


def separate_sharded(model, processor, descriptions: List[str], audios: List[torch.Tensor], noise=None,
                     reranking_candidates: int = 1, group=None):
    """Every rank passes the same full clip list; each separates its contiguous shard and all ranks
    return the full, ordered (target, residual) lists.  Clips are equal-length-padded per shard exactly as
    the single-process call would pad them, so results match a single-GPU run clip for clip."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(audios)
    lo, hi = shard_range(n, rank, world)
    counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    dev = model.device()
    hop = processor.audio_hop_length
    longest = max(a.shape[-1] for a in audios)
    S = ((longest + hop - 1) // hop) * hop
    local = torch.zeros(max(hi - lo, 0), 2, S, device=dev)
    if hi > lo:
        # pad the shard to the global longest clip so that every rank's T matches the 1-GPU batch
        batch = processor(descriptions=descriptions[lo:hi], audios=audios[lo:hi])
        if batch.audios.shape[-1] < longest:
            batch.audios = torch.nn.functional.pad(batch.audios, (0, longest - batch.audios.shape[-1]))
            full = torch.zeros(hi - lo, (longest + hop - 1) // hop, dtype=torch.bool)
            full[:, : batch.audio_pad_mask.shape[1]] = batch.audio_pad_mask
            batch.audio_pad_mask = full
            batch.process_anchors(batch.anchors)
        batch = batch.to(dev)
        c = reranking_candidates
        nz = None if noise is None else noise[lo * c: hi * c].to(dev)
        out = model.separate(batch, noise=nz, reranking_candidates=c)
        for i, (t, r) in enumerate(zip(out.target, out.residual)):
            local[i, 0, : t.numel()] = t
            local[i, 1, : r.numel()] = r
    full = all_gather_waveforms(local, counts, group=group)
    sizes = [((a.shape[-1] + hop - 1) // hop) * hop for a in audios]
    return [full[i, 0, : sizes[i]] for i in range(n)], [full[i, 1, : sizes[i]] for i in range(n)]
