"""Synthetic KV-Compress state (random-init, Llama-shaped) for tests and bench.py.

Builds exactly the tensors the reference engine owns when compression is scheduled
(vllm/kvcompress/block.py:94-126 block tables / context lens; metrics.py:220-275 metric,
position and per-block metadata arrays; worker/cache_engine.py:106-123 unified cache) with
deterministic contents, on any torch device.  Pure tensor plumbing, no kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

MAX_INT = 2147483000


@dataclass
class KVCState:
    S: int
    L: int
    H: int
    B: int
    D: int
    x: int
    NB: int
    MB: int
    block_tables: torch.Tensor      # [L, S, H, MB] i32
    context_lens: torch.Tensor      # [L, S, H] i32
    k_cache: torch.Tensor           # [NB, D/x, B, x]
    v_cache: torch.Tensor           # [NB, D, B]
    kv_metrics: torch.Tensor        # [NB, B] f32
    kv_position: torch.Tensor       # [NB, B] i32
    seq_by_block: torch.Tensor      # [NB] i32 (-1 = free)
    layer_by_block: torch.Tensor    # [NB] i32
    head_by_block: torch.Tensor     # [NB] i32
    logical_by_block: torch.Tensor  # [NB] i32


def make_state(S: int, L: int, H: int, B: int, D: int, ctx, dtype=torch.float16,
               device="cpu", seed: int = 0, slack: float = 1.25, max_blocks: Optional[int] = None,
               fill_cache: bool = True, blocks_per_head_alloc: Optional[int] = None) -> KVCState:
    """ctx: int or int tensor [L, S, H] of per-head context lengths.

    blocks_per_head_alloc: give every head a fixed slab of that many physical blocks (its block
    table row is fully populated) instead of exactly ceil(ctx/B); used by the decode harness
    whose heads grow and shrink inside their slab."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if isinstance(ctx, int):
        ctx = torch.full((L, S, H), ctx, dtype=torch.int32)
    ctx = ctx.to(torch.int32).cpu()
    nblk = (ctx + B - 1) // B                                # [L,S,H]
    if blocks_per_head_alloc is not None:
        assert int(nblk.max()) <= blocks_per_head_alloc
        nblk = torch.full_like(nblk, blocks_per_head_alloc)
    total = int(nblk.sum())
    NB = int(total * slack) + 8
    MB = int(max_blocks if max_blocks is not None else int(nblk.max()) + 1)
    perm = torch.randperm(NB, generator=g)[:total].to(torch.int32)
    flat_n = nblk.reshape(-1).to(torch.int64)
    head_of = torch.repeat_interleave(torch.arange(flat_n.numel()), flat_n)      # flat (l,s,h)
    start = torch.cumsum(flat_n, 0) - flat_n
    logical = torch.arange(total) - start[head_of]
    block_tables = torch.zeros((L * S * H, MB), dtype=torch.int32)
    block_tables[head_of, logical] = perm
    block_tables = block_tables.view(L, S, H, MB)
    seq_by_block = torch.full((NB,), -1, dtype=torch.int32)
    layer_by_block = torch.zeros(NB, dtype=torch.int32)
    head_by_block = torch.zeros(NB, dtype=torch.int32)
    logical_by_block = torch.zeros(NB, dtype=torch.int32)
    p64 = perm.long()
    layer_by_block[p64] = (head_of // (S * H)).to(torch.int32)
    seq_by_block[p64] = ((head_of // H) % S).to(torch.int32)
    head_by_block[p64] = (head_of % H).to(torch.int32)
    logical_by_block[p64] = logical.to(torch.int32)
    # token position == logical index right after prefill (block.py:275-321)
    kv_position = (logical_by_block[:, None] * B + torch.arange(B, dtype=torch.int32)[None]).contiguous()
    kv_position[seq_by_block < 0] = -1
    dev = torch.device(device)
    x = 16 // torch.empty((), dtype=dtype).element_size()
    if fill_cache:
        gd = torch.Generator(device=dev).manual_seed(seed + 1)
        scale = float(D) ** -0.5  # vllm/utils.py:664-720: U(-1/sqrt(D), 1/sqrt(D))
        k_cache = torch.empty((NB, D // x, B, x), device=dev, dtype=dtype)
        v_cache = torch.empty((NB, D, B), device=dev, dtype=dtype)
        step = 1 << 18   # blocks per chunk: bounds the fp32 scratch to 2 GB at large caches
        for cache in (k_cache, v_cache):
            for b0 in range(0, NB, step):
                n = min(step, NB - b0)
                cache[b0:b0 + n] = (torch.rand((n,) + tuple(cache.shape[1:]), generator=gd, device=dev,
                                               dtype=torch.float32).mul_(2 * scale).sub_(scale).to(dtype))
    else:
        k_cache = torch.zeros((NB, D // x, B, x), device=dev, dtype=dtype)
        v_cache = torch.zeros((NB, D, B), device=dev, dtype=dtype)
    # distinct metrics (ties are defined away)
    kv_metrics = (torch.randperm(NB * B, generator=g).to(torch.float32) / float(NB * B)).view(NB, B)
    mv = lambda t: t.to(dev)
    return KVCState(S, L, H, B, D, x, NB, MB, mv(block_tables), mv(ctx), k_cache, v_cache,
                    mv(kv_metrics), mv(kv_position), mv(seq_by_block), mv(layer_by_block),
                    mv(head_by_block), mv(logical_by_block))


def hanging_token_counts(context_lens: torch.Tensor, B: int) -> torch.Tensor:
    """vllm/kvcompress/block.py:330-335: ctx % B or B."""
    r = context_lens % B
    return torch.where(r == 0, torch.full_like(r, B), r)


def evicted_kv_offsets(context_lens_lsh: torch.Tensor, B: int) -> torch.Tensor:
    """vllm/kvcompress/scheduler.py:274-280: exclusive cumsum of ceil(ctx/B)*B in (s,l,h) order."""
    alloc = ((context_lens_lsh.transpose(0, 1) + B - 1) // B) * B
    flat = alloc.flatten().cumsum(0)
    off = torch.cat([torch.zeros_like(flat[:1]), flat[:-1]])
    return off.reshape(alloc.shape).to(torch.int32)


@dataclass
class SortedMetrics:
    """What CompressionMetrics.sort_seq_metrics hands to schedule_cache_evictions
    (vllm/kvcompress/metrics.py:849-966)."""
    sorted_indices: torch.Tensor
    seq_block_offsets: torch.Tensor
    layer_by_block: torch.Tensor
    head_by_block: torch.Tensor
    logical_block_num_by_block: torch.Tensor
    token_positions: torch.Tensor
    masked_blocks: torch.Tensor  # physical ids of the masked blocks (ours, for checks)


def sort_seq_metrics(st: KVCState) -> SortedMetrics:
    """Mask the allocated blocks (physical order), sort their slots by (seq, metric)."""
    # allocated blocks = blocks below ceil(ctx/B) of their head (a slab may hold spare blocks)
    lsh = (st.layer_by_block.long() * st.S + st.seq_by_block.long().clamp(min=0)) * st.H + st.head_by_block.long()
    nblk_of = (st.context_lens.flatten()[lsh] + st.B - 1) // st.B
    mask = (st.seq_by_block >= 0) & (st.logical_by_block < nblk_of)
    blocks = torch.nonzero(mask).flatten()
    m = st.kv_metrics[blocks].flatten()
    seqs = st.seq_by_block[blocks].repeat_interleave(st.B)
    order = torch.sort(m, stable=True).indices
    order = order[torch.sort(seqs[order], stable=True).indices]
    counts = torch.bincount(st.seq_by_block[blocks].long(), minlength=st.S)
    offs = torch.cumsum(counts, 0) - counts
    return SortedMetrics(order.to(torch.int32).contiguous(), offs.to(torch.int32),
                         st.layer_by_block[blocks].contiguous(), st.head_by_block[blocks].contiguous(),
                         st.logical_by_block[blocks].contiguous(), st.kv_position[blocks].contiguous(),
                         blocks)
