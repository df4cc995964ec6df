"""Compression scheduling: which sequences to compress, by how many blocks, and the device-side cycle that turns
those budgets into eviction lists, move lists and freed blocks.

Mirror of the reference's vllm/kvcompress/scheduler.py (CompressionScheduler :44-575) with the same method names,
argument meaning and results, so that the engine code that calls it (`scheduler.schedule_compression(seqs,
sampling_params)` in vllm/core/scheduler.py, outputs consumed by vllm/worker/cache_engine.py:139-151) needs no
change:

  _schedule_seq_evictions   policy of one sequence (:100-181): compress_once, max_cache_tokens rounded up to a
                            block multiple, target_compression_rate with the protected window rounded up to whole
                            blocks, even_layer_evict rounding -- pure host arithmetic on two integers per sequence
  _schedule_compression     batch choice (:193-233): sequences ordered by iterations since their last compression
                            (ties: larger seq_id first), skipped when their budget is 0, cut off where the running
                            KV total exceeds max_kv_per_compression; then the cycle below, ordered by batch slot
  schedule_compression      compression_interval / new_token_limit gating (:565-575)

What differs is where the work happens.  The reference asks the block manager for every sequence's KV and block
count with one device->host sync each (block_manager.py:347-354) and runs selection as a chain of ~6 torch.sort
calls with a host sync per sequence (metrics.py:441-847).  Here ONE reduction + one small D2H copy fetches the
counts of all candidate sequences, and the cycle is four launches on the batch's gathered tables:
prepare_compression (hanging counts, segment offsets) -> select_evictions -> schedule_t1_cache_moves -> (caller:
execute_cache_moves) -> alloc_free_compressed.  `backend` supplies those operations: DeviceBackend below (CUDA, this
package's ops) or a test double over the CPU oracle (tests/test_scheduler_chain.py).

reference_v2_batch_quirk: the live reference shrinks the block budget of every sequence but the first of a
multi-sequence call by the number of non-evictable (+inf) blocks of the sequences in front of it
(`[:unevicted_offset]` counted from row 0 instead of the sequence's own offset, metrics.py:718), and can even cut
into the evictions of earlier sequences.  Off by default (select_evictions follows the reference's CUDA kernel K1,
the acceptance oracle its author used, :690-697); `True` reproduces the live behaviour bit for bit through
reference_v2_budgets -- that is what the golden replay of the reference's own scheduler is compared with.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

MAX_INT = 2147483000


@dataclass
class CacheMoves:
    """reference scheduler.py:19-31: inputs of execute_cache_moves"""
    index: object     # [max_kv, 2] int32 (dst, src) physical slot ids
    count: object     # [S, L, H] int32
    offsets: object   # [S, L, H] int32


@dataclass
class CompressionOutputs:
    """reference scheduler.py:34-41"""
    cache_moves: CacheMoves
    freed_block_count: Dict[int, object]   # seq_id -> [L, H] int32


@dataclass
class SchedulerConfig:
    """the KVCompressConfig fields the scheduler reads (vllm/config.py KVCompressConfig)"""
    num_layers: int
    num_kv_heads: int
    block_size: int = 16
    max_kv_per_compression: int = 5_000_000
    compression_interval: int = 1
    new_token_limit: int = -1
    even_layer_evict: bool = False
    reference_v2_batch_quirk: bool = False


def reference_v2_budgets(budgets: Sequence[int], total_blocks: Sequence[int],
                         finite_blocks: Sequence[int]) -> List[int]:
    """Effective per-sequence block budgets of the live reference selection for a multi-sequence call.

    metrics.py:709-729 walks the sequences over one array of block rows sorted by (sequence, completion metric),
    +inf rows last within a sequence.  For sequence i starting at row offset_i it keeps rows
    [offset_i, U_i) with U_i = offset_i + budget_i - #inf rows in [0, offset_i + budget_i) -- counted from row 0,
    so the +inf rows of all earlier sequences are subtracted as well -- and overwrites [U_i, end_i) with
    MAX_INT; when U_i < offset_i that range reaches back into the rows of earlier sequences.  A row of sequence
    k therefore survives iff it lies below min_{i >= k} U_i.  finite_blocks[i] = rows of sequence i whose completion
    metric is finite (see finite_completion_blocks)."""
    n = len(budgets)
    offs, u = [], []
    off = inf_before = 0
    for i in range(n):
        b, tot, fin = int(budgets[i]), int(total_blocks[i]), int(finite_blocks[i])
        offs.append(off)
        u.append(off + b - (inf_before + max(0, b - fin)))
        inf_before += tot - fin
        off += tot
    out = []
    lim = math.inf
    for k in range(n - 1, -1, -1):
        lim = min(lim, u[k])
        out.append(max(0, int(lim) - offs[k]))
    return out[::-1]


class CompressionScheduler:
    """See the module docstring.  `seq` objects need `.seq_id`, `.data.get_len()` and a writable `.compressed`;
    `sampling_params` entries need target_compression_rate, max_cache_tokens, protected_window_size,
    compress_once (vllm/sampling_params.py in the fork)."""

    def __init__(self, config: SchedulerConfig, backend) -> None:
        self.config = config
        self.backend = backend
        self.block_size = config.block_size
        self.iteration_count = 0
        self.new_tokens = 0
        self._iters_since_compression: Dict[int, int] = {}

    # ------------------------------------------------------------------ bookkeeping (:86-97)
    def complete_seqs(self, completed_seqs) -> None:
        for seq in completed_seqs:
            self._iters_since_compression.pop(seq.seq_id, None)

    def _increment_iters_since_compression(self, compressed_seqs) -> None:
        for seq_id in self._iters_since_compression:
            self._iters_since_compression[seq_id] += 1
        for seq in compressed_seqs:
            self._iters_since_compression[seq.seq_id] = 0

    def increment_new_tokens(self, new_token_count: int) -> None:
        self.new_tokens += new_token_count

    # ------------------------------------------------------------------ policy (:100-181)
    def _schedule_seq_evictions(self, seq, kv_count: int, block_count: int, target_compression_rate: float,
                                max_cache_tokens: int, protected_window_size: int,
                                compress_once: bool) -> Tuple[int, int]:
        """Blocks of this sequence to free in the next compression.  kv_count / block_count: KVs and allocated
        blocks of the sequence over all layers and heads (the reference reads them from the block manager)."""
        if compress_once and seq.compressed:
            return 0, 0
        seq.compressed = True
        B = self.block_size
        if max_cache_tokens > 0:     # round up: never free a block that holds KVs below the limit (:117-121)
            max_cache_tokens = (max_cache_tokens + B - 1) // B * B
        if target_compression_rate < 1.0 and max_cache_tokens > 0:
            raise RuntimeError("both compression_rate and max_cache_tokens specified during compression")
        total_kv_heads = self.config.num_layers * self.config.num_kv_heads
        if max_cache_tokens >= 0:
            max_cache_kv = max_cache_tokens * total_kv_heads
            max_cache_blocks = (max_cache_kv + B - 1) // B
            evict_block_count = max(0, block_count - max_cache_blocks)
        else:
            protected_tokens = (protected_window_size + B - 1) // B * B
            compressible_token_count = seq.data.get_len() - protected_tokens
            if compressible_token_count <= 0:
                return 0, 0
            compressible_kv_count = compressible_token_count * total_kv_heads
            protected_kv = protected_tokens * total_kv_heads
            target_kv_count = math.ceil(compressible_kv_count * target_compression_rate) + protected_kv
            evict_kv_count = max(0, kv_count - target_kv_count)
            evict_block_count = (evict_kv_count + B - 1) // B
        if self.config.even_layer_evict:
            evict_block_count = evict_block_count // self.config.num_layers * self.config.num_layers
        assert evict_block_count <= max(block_count - (protected_window_size + B - 1) // B * total_kv_heads, 0)
        return evict_block_count * B, evict_block_count

    # ------------------------------------------------------------------ batch choice + cycle (:184-560)
    def _schedule_compression(self, seqs, sampling_params) -> Optional[CompressionOutputs]:
        be = self.backend
        slots_all = [be.get_slot_index(s) for s in seqs]
        kv_counts, block_counts = be.sequence_counts(slots_all)        # one reduction, one D2H copy
        by_id = {s.seq_id: (int(kv), int(nb)) for s, kv, nb in zip(seqs, kv_counts, block_counts)}
        total_kv_count = 0
        chosen = []
        order = sorted([(self._iters_since_compression.get(s.seq_id, 0), s.seq_id, s, sp)
                        for s, sp in zip(seqs, sampling_params)], key=lambda t: (t[0], t[1]), reverse=True)
        for _, _, seq, sp in order:
            kv, nb = by_id[seq.seq_id]
            _, evicted_block_count = self._schedule_seq_evictions(
                seq, kv, nb, target_compression_rate=sp.target_compression_rate,
                max_cache_tokens=sp.max_cache_tokens, protected_window_size=sp.protected_window_size,
                compress_once=sp.compress_once)
            if evicted_block_count == 0:
                continue
            total_kv_count += nb * self.block_size
            if total_kv_count > self.config.max_kv_per_compression:
                break
            chosen.append((seq, evicted_block_count, sp.protected_window_size))
        if not chosen:
            return None
        chosen.sort(key=lambda c: be.get_slot_index(c[0]))
        seqs_c = [c[0] for c in chosen]
        budgets = [c[1] for c in chosen]
        protected = [c[2] for c in chosen]
        slot_indices = [be.get_slot_index(s) for s in seqs_c]
        last_positions = [s.data.get_len() - 1 for s in seqs_c]       # :257-260
        if self.config.reference_v2_batch_quirk and len(seqs_c) > 1:
            total = [by_id[s.seq_id][1] for s in seqs_c]
            finite = be.finite_completion_blocks(slot_indices, last_positions, protected)
            budgets = reference_v2_budgets(budgets, total, finite)
        moves, evicted_block_count = be.compress(slot_indices, last_positions, budgets, protected)
        freed = {s.seq_id: evicted_block_count[i] for i, s in enumerate(seqs_c)}
        self._increment_iters_since_compression(seqs_c)
        be.free_compressed(slot_indices, evicted_block_count)          # block_manager.py:466-530 + remove_metadata
        return CompressionOutputs(moves, freed)

    def schedule_compression(self, seqs, sampling_params, force: bool = False) -> Optional[CompressionOutputs]:
        self.iteration_count += 1
        if force or (self.iteration_count >= self.config.compression_interval
                     or (self.config.new_token_limit > -1 and self.new_tokens > self.config.new_token_limit)):
            self.iteration_count = 0
            self.new_tokens = 0
            return self._schedule_compression(seqs, sampling_params)
        return None


class DeviceBackend:
    """The scheduler's view of the device state: this package's allocator (capi.DeviceBlockAllocator over the
    reference's own tensors: block tables [L, max_seqs, H, MB], context lens [L, max_seqs, H], free mask,
    per-block metadata), the metric / position arrays and the CUDA ops.  All results stay on the device except the
    two small count vectors the policy needs."""

    def __init__(self, alloc, kv_metrics, kv_position, slot_of_seq: Dict[int, int], max_kv_per_compression: int,
                 num_sinks: int = 0):
        import torch
        from . import _custom_ops as ops
        self.torch, self.ops = torch, ops
        self.alloc = alloc
        self.kv_metrics, self.kv_position = kv_metrics, kv_position
        self.slot_of_seq = slot_of_seq
        self.num_sinks = num_sinks
        dev = kv_metrics.device
        self.B = alloc.B
        # workspaces sized once, like the reference's (scheduler.py:75-86)
        self.evicted_logical = torch.empty(max_kv_per_compression, dtype=torch.int32, device=dev)
        self.cache_moves = torch.empty((max_kv_per_compression, 2), dtype=torch.int32, device=dev)

    def get_slot_index(self, seq) -> int:
        return self.slot_of_seq[seq.seq_id]

    def sequence_counts(self, slots: List[int]):
        t = self.torch
        idx = t.tensor(slots, dtype=t.long, device=self.alloc.context_lens.device)
        ctx = self.alloc.context_lens[:, idx]                                   # [L, n, H]
        both = t.stack([ctx.sum(dim=(0, 2)), ((ctx + self.B - 1) // self.B).sum(dim=(0, 2))]).cpu()
        return both[0].tolist(), both[1].tolist()

    def finite_completion_blocks(self, slots, last_positions, protected) -> List[int]:
        """rows of each sequence whose block-completion metric is finite in the reference's V2 selection
        (metrics.py:539-596): a head with e evictable KVs and r0 hanging tokens completes floor((e - r0) / B) + 1
        blocks (0 if e < r0)"""
        t, B = self.torch, self.B
        dev = self.alloc.context_lens.device
        idx = t.tensor(slots, dtype=t.long, device=dev)
        ctx = self.alloc.context_lens[:, idx]                                    # [L, n, H]
        bt = self.alloc.block_tables[:, idx].long()                              # [L, n, H, MB]
        nblk = (ctx + B - 1) // B
        MB = bt.shape[-1]
        live_blk = t.arange(MB, device=dev)[None, None, None] < nblk[..., None]
        pos = self.kv_position[bt.clamp(min=0)]                                  # [L, n, H, MB, B]
        tok = (t.arange(MB * B, device=dev).view(MB, B))[None, None, None] < ctx[..., None, None]
        lim = (t.tensor(last_positions, device=dev) - t.tensor(protected, device=dev)).view(1, -1, 1, 1, 1)
        ok = tok & live_blk[..., None] & (pos <= lim) & (pos >= self.num_sinks)
        e = ok.sum(dim=(3, 4))
        r0 = t.where(ctx % B == 0, t.full_like(ctx, B), ctx % B)
        fin = t.where(e >= r0, (e - r0) // B + 1, t.zeros_like(e))
        fin = t.minimum(fin, nblk)
        return fin.sum(dim=(0, 2)).cpu().tolist()

    def compress(self, slots, last_positions, budgets, protected):
        t, ops, B = self.torch, self.ops, self.B
        dev = self.alloc.context_lens.device
        idx = t.tensor(slots, dtype=t.long, device=dev)
        ctx = self.alloc.context_lens[:, idx].contiguous()                       # batch view copies, as the
        bt = self.alloc.block_tables[:, idx].contiguous()                        # reference's (block.py:326-327)
        L, n, H = ctx.shape
        hanging = t.empty((n, L, H), dtype=t.int32, device=dev)
        offs = t.empty((n, L, H), dtype=t.int32, device=dev)
        scratch = t.empty((n,), dtype=t.int32, device=dev)
        ops.prepare_compression(ctx, hanging, offs, scratch, B, 0)
        i32 = lambda v: t.tensor(v, dtype=t.int32, device=dev)
        kv_count = t.empty((n, L, H), dtype=t.int32, device=dev)
        blk_count = t.empty((n, L, H), dtype=t.int32, device=dev)
        ops.select_evictions(self.evicted_logical, kv_count, blk_count, self.kv_metrics, self.kv_position, bt, ctx,
                             hanging, offs, i32(budgets), i32(last_positions), i32(protected), B, MAX_INT,
                             num_sinks=self.num_sinks)
        moves_count = t.empty((n, L, H), dtype=t.int32, device=dev)
        ops.schedule_cache_moves(self.cache_moves, moves_count, self.evicted_logical, kv_count, offs, bt, ctx, B)
        return CacheMoves(self.cache_moves, moves_count, offs), blk_count

    def free_compressed(self, slots, evicted_block_count) -> None:
        t = self.torch
        self.alloc.free_compressed(t.tensor(slots, dtype=t.int32, device=evicted_block_count.device),
                                   evicted_block_count)
