"""ctypes/numpy front-end of oracle/kvc_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl
reference`` legs may import this module.  The product package
(``vllm_kvcompress_b200``) never does: it fails loudly when its CUDA extension
is missing instead of falling back to this CPU code.

Function names and argument order mirror the reference's Python op shim
(vllm/_custom_ops.py:934-1256) with numpy arrays in place of torch tensors;
each docstring cites the reference lines the C body follows.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "kvc_oracle.c")
_SO = os.path.join(_HERE, "libkvc_oracle.so")
_lib = None

MAX_INT = 2147483000  # vllm/kvcompress/scheduler.py:16


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -shared oracle/kvc_oracle.c -> oracle/libkvc_oracle.so"""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def num_threads() -> int:
    return int(lib().kvc_oracle_num_threads())


def use_all_cores() -> int:
    """Run the OpenMP loops on every core this process may use (torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would make the CPU baseline single-threaded under --gpus N > 1)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().kvc_oracle_set_num_threads(int(n))
    return num_threads()


def _p(a: Optional[np.ndarray]):
    if a is None:
        return ctypes.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"], "oracle expects C-contiguous arrays"
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def schedule_cache_evictions(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                             remaining_kv_count, evicted_kv_offsets, sorted_indices,
                             seq_block_offsets, layer_by_block, head_by_block,
                             virtual_block_num_by_block, evicted_blocks_per_seq, context_lens,
                             hanging_token_count, kv_position, last_position,
                             protected_window_size, block_size, evict_evenly_per_layer,
                             control_layers, max_evicted_kv, null_eviction_index, truncate):
    """K1 (+K2): csrc/kvcompress_eviction_kernels.cu:27-188, launcher :492-555.

    All arrays int32; context_lens / hanging_token_count are [S, L, H].
    Outputs (first four arrays) are written in place.
    """
    S, L, H = evicted_kv_count.shape
    for a in (evicted_kv_indices, evicted_logical_indices, evicted_kv_count, remaining_kv_count):
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    total_blocks = layer_by_block.shape[0]
    cl = None if control_layers is None else _i32(control_layers)
    lib().kvc_oracle_schedule_cache_evictions(
        _p(evicted_kv_indices), _p(evicted_logical_indices), _p(evicted_kv_count),
        _p(remaining_kv_count), _p(_i32(evicted_kv_offsets)), _p(_i32(sorted_indices)),
        _p(_i32(seq_block_offsets)), _p(_i32(layer_by_block)), _p(_i32(head_by_block)),
        _p(_i32(virtual_block_num_by_block)), _p(_i32(evicted_blocks_per_seq)),
        _p(_i32(context_lens)), _p(_i32(hanging_token_count)), _p(_i32(kv_position)),
        _p(_i32(last_position)), _p(_i32(protected_window_size)),
        ctypes.c_int(S), ctypes.c_int(L), ctypes.c_int(H), ctypes.c_int(total_blocks),
        ctypes.c_int(block_size), ctypes.c_int(int(evict_evenly_per_layer)), _p(cl),
        ctypes.c_int(0 if cl is None else cl.shape[0]), ctypes.c_int(max_evicted_kv),
        ctypes.c_int(null_eviction_index), ctypes.c_int(int(truncate)))


def truncate_cache_evictions(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                             evicted_kv_offsets, hanging_token_count, block_size,
                             max_evicted_kv, null_eviction_index):
    """K2: csrc/kvcompress_eviction_kernels.cu:158-188."""
    for a in (evicted_kv_indices, evicted_logical_indices, evicted_kv_count):
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    lib().kvc_oracle_truncate_cache_evictions(
        _p(evicted_kv_indices), _p(evicted_logical_indices), _p(evicted_kv_count),
        _p(_i32(evicted_kv_offsets)), _p(_i32(hanging_token_count)),
        ctypes.c_int(evicted_kv_count.size), ctypes.c_int(block_size),
        ctypes.c_int(max_evicted_kv), ctypes.c_int(null_eviction_index))


def count_block_evictions(evicted_block_count, evicted_logical_indices, evicted_kv_offsets,
                          hanging_token_count, block_size, null_value):
    """K3: csrc/kvcompress_eviction_kernels.cu:190-221."""
    for a in (evicted_block_count, evicted_logical_indices):
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    lib().kvc_oracle_count_block_evictions(
        _p(evicted_block_count), _p(evicted_logical_indices), _p(_i32(evicted_kv_offsets)),
        _p(_i32(hanging_token_count)), ctypes.c_int(evicted_block_count.size),
        ctypes.c_int(evicted_logical_indices.size), ctypes.c_int(block_size),
        ctypes.c_int(null_value))


def schedule_cache_moves(cache_moves_idx, cache_moves_count, evicted_logical_indices,
                         evicted_kv_count, evicted_kv_offsets, block_tables, context_lens,
                         block_size):
    """K4: csrc/kvcompress_eviction_kernels.cu:223-289.

    block_tables [L, S, H, MB], context_lens [L, S, H]; the rest [S, L, H].
    NB: unlike the reference wrapper (_custom_ops.py:1168) this does not
    zero-fill cache_moves_idx first; callers that compare whole buffers do.
    """
    for a in (cache_moves_idx, cache_moves_count):
        assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    S, L, H = evicted_kv_count.shape
    lib().kvc_oracle_schedule_t1_cache_moves(
        _p(cache_moves_idx), _p(cache_moves_count), _p(_i32(evicted_logical_indices)),
        _p(_i32(evicted_kv_count)), _p(_i32(evicted_kv_offsets)), _p(_i32(block_tables)),
        _p(_i32(context_lens)), ctypes.c_int(S), ctypes.c_int(L), ctypes.c_int(H),
        ctypes.c_int(block_tables.shape[3]), ctypes.c_int(block_size))


def execute_cache_moves(k_cache, v_cache, kv_metrics, kv_position, cache_moves_idx,
                        cache_moves_count, evicted_kv_offsets):
    """K5: csrc/kvcompress_eviction_kernels.cu:359-435.  In-place bit copy.

    k_cache [NB, D/x, B, x], v_cache [NB, D, B] of any 1/2/4-byte dtype
    (pass fp16/bf16 as uint16 views), kv_metrics [NB, B] f32, kv_position [NB, B] i32.
    """
    assert k_cache.flags["C_CONTIGUOUS"] and v_cache.flags["C_CONTIGUOUS"]
    assert kv_metrics.dtype == np.float32 and kv_position.dtype == np.int32
    NB, D, B = v_cache.shape
    x = k_cache.shape[3]
    lib().kvc_oracle_execute_cache_moves(
        _p(k_cache), _p(v_cache), _p(kv_metrics), _p(kv_position),
        _p(_i32(cache_moves_idx)), _p(_i32(cache_moves_count)), _p(_i32(evicted_kv_offsets)),
        ctypes.c_int(cache_moves_count.size), ctypes.c_int(D), ctypes.c_int(x),
        ctypes.c_int(B), ctypes.c_int(k_cache.dtype.itemsize))


def reshape_and_cache_kvc(key, value, key_cache, value_cache, kv_metrics, slot_mapping,
                          kv_metric_head_bias):
    """K6: csrc/kvcompress_cache_kernels.cu:27-89 (kv_cache_dtype == 'auto')."""
    T, H, D = key.shape
    B, x = key_cache.shape[2], key_cache.shape[3]
    assert key.flags["C_CONTIGUOUS"] and value.flags["C_CONTIGUOUS"]
    sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
    lib().kvc_oracle_reshape_and_cache(
        _p(key), _p(value), _p(key_cache), _p(value_cache), _p(kv_metrics), _p(sm),
        _p(np.ascontiguousarray(kv_metric_head_bias, dtype=np.float32)),
        ctypes.c_int(T), ctypes.c_int(H), ctypes.c_int(D), ctypes.c_int(B), ctypes.c_int(x),
        ctypes.c_int64(H * D), ctypes.c_int64(H * D), ctypes.c_int(key.dtype.itemsize))


def paged_attention_kvc(out, kv_metric_out, query, key_cache, value_cache, num_kv_heads,
                        scale, block_tables, context_lens, kv_position, last_position,
                        kv_metric_buffer_len, alibi_slopes, is_bf16, record_kv_metrics):
    """K7/K8 dense fp32 restatement: csrc/attention/kvcompress_attention_kernels.cu:95-451.

    query/key_cache/value_cache are uint16 views of fp16 (or bf16) data;
    out [S, Hq, D] float32; kv_metric_out [NB, B, Q] float32 (in place).
    """
    S, Hq, D = query.shape
    NB, _, B = value_cache.shape
    x = key_cache.shape[3]
    assert out.dtype == np.float32 and kv_metric_out.dtype == np.float32
    assert query.dtype == np.uint16 and key_cache.dtype == np.uint16
    al = None if alibi_slopes is None else np.ascontiguousarray(alibi_slopes, dtype=np.float32)
    lib().kvc_oracle_paged_attention(
        _p(out), _p(kv_metric_out), _p(np.ascontiguousarray(query)), _p(key_cache),
        _p(value_cache), _p(_i32(block_tables)), _p(_i32(context_lens)), _p(_i32(kv_position)),
        _p(_i32(last_position)), _p(_i32(kv_metric_buffer_len)), _p(al),
        ctypes.c_int(S), ctypes.c_int(Hq), ctypes.c_int(num_kv_heads), ctypes.c_int(D),
        ctypes.c_int(B), ctypes.c_int(x), ctypes.c_int(block_tables.shape[2]),
        ctypes.c_int64(Hq * D), ctypes.c_float(scale), ctypes.c_int(int(is_bf16)),
        ctypes.c_int(int(record_kv_metrics)))


def aggregate_decode(metrics, temp_metrics, use_l2=True):
    """vllm/kvcompress/metrics.py:429-439: metrics += (temp ** 2).sum(-1)."""
    assert metrics.dtype == np.float32 and temp_metrics.dtype == np.float32
    lib().kvc_oracle_aggregate_decode(_p(metrics), _p(temp_metrics),
                                      ctypes.c_int64(metrics.size),
                                      ctypes.c_int(temp_metrics.shape[-1]),
                                      ctypes.c_int(int(use_l2)))


def sort_evictions_by_head(evicted_kv_indices, evicted_logical_indices):
    """Host step between K1/K2 and K4 (vllm/kvcompress/scheduler.py:472-477):
    re-sort so that every head segment holds its evicted logical indices
    ascending with NULLs last.  Pure index work (stable sorts), numpy.
    """
    order = np.argsort(evicted_logical_indices, kind="stable")
    order2 = np.argsort(evicted_kv_indices[order], kind="stable")
    perm = order[order2]
    return evicted_logical_indices[perm].copy()


def prefill_kv_metrics(query, key, seq_lens, scale, kv_metric_buffer_len, n_observed=32,
                       max_observed_block_size=4096, use_l2=True, use_average=False, use_maxpool=True,
                       is_bf16=False):
    """_naive_kvc_attention (vllm/attention/backends/flash_attn.py:1122-1211) with an un-repeated
    key [T, H, D]; query/key are uint16 bit patterns.  Returns kv_metric_out [T, Hq] f32."""
    T, Hq, D = query.shape
    H = key.shape[1]
    assert query.dtype == np.uint16 and key.dtype == np.uint16
    starts = np.concatenate([[0], np.cumsum(np.asarray(seq_lens, dtype=np.int64))]).astype(np.int32)
    assert starts[-1] == T
    out = np.zeros((T, Hq), np.float32)
    lib().kvc_oracle_prefill_kv_metrics(
        _p(out), _p(np.ascontiguousarray(query)), _p(np.ascontiguousarray(key)), _p(starts),
        _p(_i32(kv_metric_buffer_len)), ctypes.c_int(T), ctypes.c_int(len(seq_lens)), ctypes.c_int(Hq),
        ctypes.c_int(H), ctypes.c_int(D), ctypes.c_float(scale), ctypes.c_int(n_observed),
        ctypes.c_int(max_observed_block_size), ctypes.c_int(int(use_l2)), ctypes.c_int(int(use_average)),
        ctypes.c_int(int(use_maxpool)), ctypes.c_int(int(is_bf16)))
    return out


def aggregate_prefill(metrics, prefill_metrics, slot_mapping):
    """CompressionMetrics.aggregate_prefill (vllm/kvcompress/metrics.py:395-427), in place on
    `metrics` (any shape, indexed flat by slot)."""
    T, Hq = prefill_metrics.shape
    H = slot_mapping.shape[1]
    assert metrics.dtype == np.float32 and prefill_metrics.dtype == np.float32
    sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
    lib().kvc_oracle_aggregate_prefill(_p(metrics), _p(np.ascontiguousarray(prefill_metrics)), _p(sm),
                                       ctypes.c_int(T), ctypes.c_int(Hq), ctypes.c_int(H))


def select_evictions(kv_metrics, kv_position, block_tables, context_lens, hanging_token_count,
                     evicted_kv_offsets, evicted_blocks_per_seq, last_position, protected_window_size,
                     block_size, null_eviction_index, head_bias=None, position_bins=None, bias_weight=0.0,
                     use_average=False, num_sinks=0):
    """CompressionMetrics.schedule_evictions (vllm/kvcompress/metrics.py:441-847) restated per head in
    numpy (small cases): key = metric [/ (last_position - position), :495-501] [+ bias[l, h,
    bin(position)] * weight, :503-506]; candidates = logical index < ctx, position <= last_position -
    protected and position >= num_sinks (:539-544); a head's block k completes at its sorted rank
    hanging-1 + k*B (:583-596); per sequence the `budget` smallest completion keys are granted
    (:671-729, ties by (key, head, block)); a head with k granted blocks evicts its (k-1)*B + hanging
    lowest keys (:774-792), listed ascending by logical index with `null` after them (:822-834).
    Returns (evicted_logical_indices [N], evicted_kv_count [S,L,H], evicted_block_count [S,L,H])."""
    B = block_size
    L, S, H = context_lens.shape
    nblk = (context_lens + B - 1) // B
    N = int(nblk.sum()) * B
    logical = np.full(N, null_eviction_index, np.int32)
    kv_count = np.zeros((S, L, H), np.int32)
    blk_count = np.zeros((S, L, H), np.int32)
    m_flat, p_flat = kv_metrics.reshape(-1).astype(np.float32), kv_position.reshape(-1)
    for s in range(S):
        heads, events = [], []
        for l in range(L):
            for h in range(H):
                n = int(context_lens[l, s, h])
                idx = np.arange(n)
                slots = block_tables[l, s, h, idx // B].astype(np.int64) * B + idx % B
                pos = p_flat[slots]
                key = m_flat[slots].copy()
                with np.errstate(divide="ignore", invalid="ignore"):
                    if use_average:
                        key = key / (int(last_position[s]) - pos).astype(np.float32)
                    if head_bias is not None:
                        bins = (pos[:, None] >= np.asarray(position_bins)[None]).sum(-1) - 1
                        b = np.where(pos >= 0, head_bias[l, h, bins], 0).astype(np.float32)
                        key = key + b * np.float32(bias_weight)
                cand = (pos <= int(last_position[s]) - int(protected_window_size[s])) & (pos >= num_sinks)
                ci = idx[cand]
                order = ci[np.lexsort((ci, key[cand]))]          # by (key, logical index)
                r0 = int(hanging_token_count[s, l, h])
                nev = (len(order) - r0) // B + 1 if len(order) >= r0 >= 1 else 0
                for k in range(nev):
                    events.append((key[order[r0 - 1 + k * B]], len(heads), k))
                heads.append((l, h, order, r0))
        events.sort()
        granted = np.zeros(len(heads), np.int64)
        for _, hd, _k in events[: int(evicted_blocks_per_seq[s])]:
            granted[hd] += 1
        for hd, (l, h, order, r0) in enumerate(heads):
            k = int(granted[hd])
            cnt = (k - 1) * B + r0 if k > 0 else 0
            kv_count[s, l, h], blk_count[s, l, h] = cnt, k
            off = int(evicted_kv_offsets[s, l, h])
            logical[off: off + cnt] = np.sort(order[:cnt])
    return logical, kv_count, blk_count


class BlockManagerOracle:
    """numpy restatement of ParallelBlockAllocator (vllm/kvcompress/block_manager.py:75-140) and the
    block-state updates of BlockSpaceManagerKVC: _add_sequence (:196-222), _append_to_sequence_batch
    (:269-294, metadata block.py:505-560 + metrics.py:344-362), free_compressed_blocks (:466-530 +
    block.py:184-210, metrics.py:364-370) and _remove_sequence_batch (:243-262)."""

    def __init__(self, L, H, B, max_seqs, MB, NB):
        self.L, self.H, self.B, self.MS, self.MB, self.NB = L, H, B, max_seqs, MB, NB
        self.block_tables = np.full((L, max_seqs, H, MB), -1, np.int32)
        self.context_lens = np.zeros((L, max_seqs, H), np.int32)
        self.free_mask = np.ones(NB, bool)
        self.seq_index_by_block = np.full(NB, -1, np.int32)
        self.layer_index_by_block = np.zeros(NB, np.int32)
        self.head_index_by_block = np.zeros(NB, np.int32)
        self.logical_block_num_by_block = np.zeros(NB, np.int32)
        self.token_positions = np.full((NB, B), -1, np.int32)

    @property
    def free_count(self):
        return int(self.free_mask.sum())

    def _allocate(self, n):  # :106-114: the n lowest-numbered free blocks, ascending
        if n > self.free_count:
            raise ValueError("Out of memory!")
        ids = np.nonzero(self.free_mask)[0][:n]
        self.free_mask[ids] = False
        return ids.astype(np.int32)

    def add_sequences(self, slots, lens):
        for slot, n in zip(slots, lens):
            nblk = (n + self.B - 1) // self.B
            ids = self._allocate(self.L * self.H * nblk).reshape(self.L, self.H, nblk)
            self.context_lens[:, slot] = n
            self.block_tables[:, slot, :, :nblk] = ids
            for l in range(self.L):
                for h in range(self.H):
                    for b in range(nblk):
                        self._meta(ids[l, h, b], slot, l, h, b, b * self.B)

    def _meta(self, blk, slot, l, h, b, pos0):
        self.seq_index_by_block[blk] = slot
        self.layer_index_by_block[blk] = l
        self.head_index_by_block[blk] = h
        self.logical_block_num_by_block[blk] = b
        self.token_positions[blk] = pos0 + np.arange(self.B)

    def append_slots(self, slots, positions):
        """+1 token per head of the batch; returns slot_mapping [L, batch, H] (block.py:303-321)."""
        ctx = self.context_lens[:, slots]                                   # [L, batch, H] copy
        need = (ctx % self.B) == 0
        ids = self._allocate(int(need.sum()))
        out = np.zeros(ctx.shape, np.int64)
        k = 0
        for l in range(self.L):
            for sb, slot in enumerate(slots):
                for h in range(self.H):
                    c = int(ctx[l, sb, h])
                    if need[l, sb, h]:
                        blk = ids[k]
                        k += 1
                        self.block_tables[l, slot, h, c // self.B] = blk
                        self._meta(blk, slot, l, h, c // self.B, positions[sb])
                    blk = self.block_tables[l, slot, h, c // self.B]
                    self.token_positions[blk, c % self.B] = positions[sb]
                    out[l, sb, h] = int(blk) * self.B + c % self.B
                    self.context_lens[l, slot, h] = c + 1
        return out

    def free_compressed(self, slots, counts):
        """counts [batch, L, H] = freed blocks per head."""
        for sb, slot in enumerate(slots):
            for l in range(self.L):
                for h in range(self.H):
                    k, c = int(counts[sb, l, h]), int(self.context_lens[l, slot, h])
                    nblk = (c + self.B - 1) // self.B
                    ids = self.block_tables[l, slot, h, max(nblk - k, 0):nblk]
                    self.free_mask[ids] = True
                    self.seq_index_by_block[ids] = -1
                    hanging = c % self.B or self.B
                    self.context_lens[l, slot, h] = c - max(k * self.B - (self.B - hanging), 0)

    def remove_sequences(self, slots):
        for slot in slots:
            for l in range(self.L):
                for h in range(self.H):
                    nblk = (int(self.context_lens[l, slot, h]) + self.B - 1) // self.B
                    ids = self.block_tables[l, slot, h, :nblk]
                    self.free_mask[ids] = True
                    self.seq_index_by_block[ids] = -1
            self.context_lens[:, slot] = 0
