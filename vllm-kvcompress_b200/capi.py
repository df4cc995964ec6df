"""ctypes view of the C ABI (include/kvc_b200.h) operating on torch CUDA tensors.

This is the raw boundary a non-torch host (cgo / JNI / plain C) would bind; the parity
tests call the kernels through it.  Tensors are only used for their data_ptr(), shape and
the current CUDA stream."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import native

KVC_F16, KVC_BF16, KVC_F32, KVC_U8 = 0, 1, 2, 3
_DT = {torch.float16: KVC_F16, torch.bfloat16: KVC_BF16, torch.float32: KVC_F32,
       torch.uint8: KVC_U8, torch.float8_e5m2: KVC_U8}


class KvcError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"kvc status {status}: {message}")
        self.status = status


def _check(rc: int) -> None:
    if rc != 0:
        raise KvcError(rc, native.load_lib().kvc_last_error().decode())


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.int32 or not t.is_contiguous() or not t.is_cuda:
        raise TypeError(f"{name} must be a contiguous int32 CUDA tensor")
    return t


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def abi_version() -> int:
    return int(native.load_lib().kvc_abi_version())


def schedule_cache_evictions(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                             remaining_kv_count, evicted_kv_offsets, sorted_indices,
                             seq_block_offsets, layer_by_block, head_by_block,
                             virtual_block_num_by_block, evicted_blocks_per_seq, context_lens,
                             hanging_token_count, kv_position, last_position,
                             protected_window_size, block_size, evict_evenly_per_layer,
                             control_layers, max_evicted_kv, null_eviction_index, truncate):
    lib = native.load_lib()
    S, L, H = evicted_kv_count.shape
    total_blocks = layer_by_block.shape[0]
    nbytes = lib.kvc_schedule_cache_evictions_workspace_bytes(S, L, H, total_blocks, block_size)
    ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=evicted_kv_count.device)
    ts = [evicted_kv_indices, evicted_logical_indices, evicted_kv_count, remaining_kv_count,
          evicted_kv_offsets, sorted_indices, seq_block_offsets, layer_by_block, head_by_block,
          virtual_block_num_by_block, evicted_blocks_per_seq, context_lens, hanging_token_count,
          kv_position, last_position, protected_window_size]
    for i, t in enumerate(ts):
        _i32(t, f"arg{i}")
    ncl = 0 if control_layers is None else control_layers.numel()
    _check(lib.kvc_schedule_cache_evictions(
        *[_ptr(t) for t in ts], S, L, H, total_blocks, block_size, int(evict_evenly_per_layer),
        _ptr(control_layers), ncl, max_evicted_kv, null_eviction_index, int(truncate),
        _ptr(ws), ctypes.c_size_t(ws.numel()), _stream()))


def truncate_cache_evictions(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                             evicted_kv_offsets, hanging_token_count, block_size, max_evicted_kv,
                             null_eviction_index):
    S, L, H = evicted_kv_count.shape
    _check(native.load_lib().kvc_truncate_cache_evictions(
        _ptr(_i32(evicted_kv_indices, "evicted_kv_indices")),
        _ptr(_i32(evicted_logical_indices, "evicted_logical_indices")),
        _ptr(_i32(evicted_kv_count, "evicted_kv_count")), _ptr(_i32(evicted_kv_offsets, "evicted_kv_offsets")),
        _ptr(_i32(hanging_token_count, "hanging_token_count")), S, L, H, block_size, max_evicted_kv,
        null_eviction_index, _stream()))


def count_block_evictions(evicted_block_count, evicted_logical_indices, evicted_kv_offsets,
                          hanging_token_count, block_size, null_value):
    S, L, H = evicted_block_count.shape
    _check(native.load_lib().kvc_count_block_evictions(
        _ptr(_i32(evicted_block_count, "evicted_block_count")),
        _ptr(_i32(evicted_logical_indices, "evicted_logical_indices")),
        _ptr(_i32(evicted_kv_offsets, "evicted_kv_offsets")),
        _ptr(_i32(hanging_token_count, "hanging_token_count")), S, L, H,
        evicted_logical_indices.numel(), block_size, null_value, _stream()))


def schedule_t1_cache_moves(cache_moves_idx, cache_moves_count, evicted_logical_indices,
                            evicted_kv_count, evicted_kv_offsets, block_tables, context_lens,
                            block_size):
    S, L, H = evicted_kv_count.shape
    _check(native.load_lib().kvc_schedule_t1_cache_moves(
        _ptr(_i32(cache_moves_idx, "cache_moves_idx")), _ptr(_i32(cache_moves_count, "cache_moves_count")),
        _ptr(_i32(evicted_logical_indices, "evicted_logical_indices")),
        _ptr(_i32(evicted_kv_count, "evicted_kv_count")), _ptr(_i32(evicted_kv_offsets, "evicted_kv_offsets")),
        _ptr(_i32(block_tables, "block_tables")), _ptr(_i32(context_lens, "context_lens")), S, L, H,
        block_tables.shape[3], block_size, _stream()))


def execute_cache_moves(k_cache, v_cache, kv_metrics, kv_position, cache_moves_idx,
                        cache_moves_count, evicted_kv_offsets, blocks_per_head=1,
                        threads_per_head=16):
    S, L, H = cache_moves_count.shape
    NB, D, B = v_cache.shape
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and kv_metrics.dtype == torch.float32
    _check(native.load_lib().kvc_execute_cache_moves(
        _ptr(k_cache), _ptr(v_cache), _ptr(kv_metrics), _ptr(_i32(kv_position, "kv_position")),
        _ptr(_i32(cache_moves_idx, "cache_moves_idx")), _ptr(_i32(cache_moves_count, "cache_moves_count")),
        _ptr(_i32(evicted_kv_offsets, "evicted_kv_offsets")), S, L, H, ctypes.c_int64(NB), D,
        k_cache.shape[3], B, _DT[k_cache.dtype], blocks_per_head, threads_per_head, _stream()))


def execute_cache_moves_deferred(k_cache, v_cache, kv_metrics, kv_position, cache_moves_idx,
                                 cache_moves_count, evicted_kv_offsets, pend_kv, pend_slot, pend_dst=None):
    S, L, H = cache_moves_count.shape
    NB, D, B = v_cache.shape
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and kv_metrics.dtype == torch.float32
    assert pend_kv.is_contiguous() and pend_kv.dtype == k_cache.dtype and pend_kv.numel() == S * L * H * 2 * D
    assert pend_slot.is_contiguous() and pend_slot.dtype == torch.int64 and pend_slot.numel() == S * L * H
    _check(native.load_lib().kvc_execute_cache_moves_deferred(
        _ptr(k_cache), _ptr(v_cache), _ptr(kv_metrics), _ptr(_i32(kv_position, "kv_position")),
        _ptr(_i32(cache_moves_idx, "cache_moves_idx")), _ptr(_i32(cache_moves_count, "cache_moves_count")),
        _ptr(_i32(evicted_kv_offsets, "evicted_kv_offsets")), _ptr(pend_kv), _ptr(pend_slot), _ptr(pend_dst), S, L, H,
        ctypes.c_int64(NB), D, k_cache.shape[3], B, _DT[k_cache.dtype], _stream()))


def reshape_and_cache(key, value, key_cache, value_cache, kv_metrics, slot_mapping,
                      kv_metric_head_bias):
    T, H, D = key.shape
    assert slot_mapping.dtype == torch.int64 and slot_mapping.is_contiguous()
    _check(native.load_lib().kvc_reshape_and_cache(
        _ptr(key), _ptr(value), _ptr(key_cache), _ptr(value_cache), _ptr(kv_metrics),
        _ptr(slot_mapping), _ptr(kv_metric_head_bias), T, H, D, key_cache.shape[2],
        key_cache.shape[3], ctypes.c_int64(key.stride(0)), ctypes.c_int64(value.stride(0)),
        _DT[key_cache.dtype], _stream()))


def _attn_args(query, key_cache, value_cache, num_kv_heads, scale, block_tables, context_lens,
               kv_position, last_position, kv_metric_buffer_len, block_size, max_context_len,
               alibi_slopes):
    S, Hq, D = query.shape
    return [_ptr(query), _ptr(key_cache), _ptr(value_cache), S, Hq, num_kv_heads, D,
            ctypes.c_float(scale), _ptr(_i32(block_tables, "block_tables")),
            _ptr(_i32(context_lens, "context_lens")), _ptr(_i32(kv_position, "kv_position")),
            _ptr(_i32(last_position, "last_position")),
            _ptr(_i32(kv_metric_buffer_len, "kv_metric_buffer_len")), block_size, max_context_len,
            block_tables.shape[2], ctypes.c_int64(query.stride(0)), ctypes.c_int64(key_cache.stride(0)),
            _ptr(alibi_slopes), _DT[query.dtype]]


def paged_attention_v1(out, kv_metric_out, query, key_cache, value_cache, num_kv_heads, scale,
                       block_tables, context_lens, kv_position, last_position,
                       kv_metric_buffer_len, block_size, max_context_len, alibi_slopes,
                       record_kv_metrics):
    a = _attn_args(query, key_cache, value_cache, num_kv_heads, scale, block_tables, context_lens,
                   kv_position, last_position, kv_metric_buffer_len, block_size, max_context_len,
                   alibi_slopes)
    _check(native.load_lib().kvc_paged_attention_v1(_ptr(out), _ptr(kv_metric_out), *a,
                                                    int(record_kv_metrics), _stream()))


def paged_attention_v2(out, kv_metric_out, exp_sums, max_logits, tmp_out, tmp_kv_metric_out, query,
                       key_cache, value_cache, num_kv_heads, scale, block_tables, context_lens,
                       kv_position, last_position, kv_metric_buffer_len, block_size,
                       max_context_len, alibi_slopes, record_kv_metrics):
    a = _attn_args(query, key_cache, value_cache, num_kv_heads, scale, block_tables, context_lens,
                   kv_position, last_position, kv_metric_buffer_len, block_size, max_context_len,
                   alibi_slopes)
    _check(native.load_lib().kvc_paged_attention_v2(
        _ptr(out), _ptr(kv_metric_out), _ptr(exp_sums), _ptr(max_logits), _ptr(tmp_out),
        _ptr(tmp_kv_metric_out), *a, int(record_kv_metrics), _stream()))


def paged_attention_fused_metrics(out, kv_metrics, query, key_cache, value_cache, num_kv_heads,
                                  scale, block_tables, context_lens, kv_position, last_position,
                                  kv_metric_buffer_len, block_size, max_context_len, alibi_slopes):
    a = _attn_args(query, key_cache, value_cache, num_kv_heads, scale, block_tables, context_lens,
                   kv_position, last_position, kv_metric_buffer_len, block_size, max_context_len,
                   alibi_slopes)
    lib = native.load_lib()
    lib.kvc_paged_attention_fused_metrics_workspace_bytes.restype = ctypes.c_size_t
    nbytes = int(lib.kvc_paged_attention_fused_metrics_workspace_bytes(query.shape[0], query.shape[1], max_context_len))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=query.device) if nbytes else None
    _check(lib.kvc_paged_attention_fused_metrics(_ptr(out), _ptr(kv_metrics), *a, _ptr(ws),
                                                 ctypes.c_size_t(nbytes), _stream()))


def select_evictions(evicted_logical_indices, evicted_kv_count, evicted_block_count, kv_metrics,
                     kv_position, block_tables, context_lens, hanging_token_count, evicted_kv_offsets,
                     evicted_blocks_per_seq, last_position, protected_window_size, block_size,
                     null_eviction_index, head_bias=None, position_bins=None, bias_weight=0.0,
                     use_average=False, num_sinks=0):
    lib = native.load_lib()
    lib.kvc_select_evictions_workspace_bytes_ex.restype = ctypes.c_size_t
    S, L, H = evicted_kv_count.shape
    max_kv = evicted_logical_indices.numel()
    nbytes = lib.kvc_select_evictions_workspace_bytes_ex(ctypes.c_int64(max_kv), block_size, S,
                                                         block_tables.shape[-1])
    ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=evicted_kv_count.device)
    assert kv_metrics.dtype == torch.float32 and kv_metrics.is_contiguous()
    ints = [kv_position, block_tables, context_lens, hanging_token_count, evicted_kv_offsets,
            evicted_blocks_per_seq, last_position, protected_window_size]
    for i, t in enumerate(ints):
        _i32(t, f"arg{i}")
    if head_bias is not None:
        assert head_bias.dtype == torch.float32 and head_bias.is_contiguous() and head_bias.shape[:2] == (L, H)
        _i32(position_bins, "position_bins")
    _check(lib.kvc_select_evictions_ex(
        _ptr(_i32(evicted_logical_indices, "evicted_logical_indices")),
        _ptr(_i32(evicted_kv_count, "evicted_kv_count")), _ptr(_i32(evicted_block_count, "evicted_block_count")),
        _ptr(kv_metrics), *[_ptr(t) for t in ints], S, L, H, block_tables.shape[3], block_size,
        ctypes.c_int64(max_kv), null_eviction_index, _ptr(head_bias), _ptr(position_bins),
        0 if head_bias is None else head_bias.shape[2], ctypes.c_float(bias_weight), int(use_average),
        int(num_sinks), _ptr(ws), ctypes.c_size_t(ws.numel()), _stream()))


def prefill_kv_metrics(kv_metric_out, kv_metrics, slot_mapping, query, key, seq_start_loc,
                       kv_metric_buffer_len, max_prompt_len, scale, n_observed=32, use_l2=True,
                       use_average=False, use_maxpool=True):
    lib = native.load_lib()
    lib.kvc_prefill_kv_metrics_workspace_bytes.restype = ctypes.c_size_t
    T, Hq, D = query.shape
    H = key.shape[1]
    num_seqs = seq_start_loc.numel() - 1
    nbytes = lib.kvc_prefill_kv_metrics_workspace_bytes(T, num_seqs, Hq, H, n_observed, max_prompt_len)
    ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=query.device)
    assert query.stride(-1) == 1 and query.stride(1) == D and key.stride(-1) == 1 and key.stride(1) == D
    for t in (kv_metric_out, kv_metrics):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    assert slot_mapping is None or (slot_mapping.dtype == torch.int64 and slot_mapping.is_contiguous())
    _check(lib.kvc_prefill_kv_metrics(
        _ptr(kv_metric_out), _ptr(kv_metrics), _ptr(slot_mapping), _ptr(query), _ptr(key),
        _ptr(_i32(seq_start_loc, "seq_start_loc")), _ptr(_i32(kv_metric_buffer_len, "kv_metric_buffer_len")),
        T, num_seqs, Hq, H, D, ctypes.c_int64(query.stride(0)), ctypes.c_int64(key.stride(0)),
        int(max_prompt_len), ctypes.c_float(scale), int(n_observed), int(use_l2), int(use_average),
        int(use_maxpool), _DT[query.dtype], _ptr(ws), ctypes.c_size_t(ws.numel()), _stream()))


def aggregate_prefill(kv_metrics, prefill_metrics, slot_mapping):
    lib = native.load_lib()
    T, Hq = prefill_metrics.shape
    H = slot_mapping.shape[1]
    assert kv_metrics.dtype == torch.float32 and kv_metrics.is_contiguous()
    assert prefill_metrics.dtype == torch.float32 and prefill_metrics.is_contiguous()
    assert slot_mapping.dtype == torch.int64 and slot_mapping.is_contiguous()
    _check(lib.kvc_aggregate_prefill(_ptr(kv_metrics), _ptr(prefill_metrics), _ptr(slot_mapping), T, Hq, H,
                                     _stream()))


def set_attention_pdl(mode: int) -> None:
    """kvc_set_attention_pdl (include/kvc_b200.h): 0 plain launches; 1 PDL, every warp waits at entry
    (default, always safe); 2 PDL, tables / cache not written by the preceding kernel (wait before q/k/v);
    3 PDL, q/k/v not produced by the preceding kernel either (wait at exit; synthetic benchmarks only)."""
    _check(native.load_lib().kvc_set_attention_pdl(int(mode)))


def get_attention_pdl() -> int:
    return int(native.load_lib().kvc_get_attention_pdl())


def set_attention_v2_mode(mode: int) -> None:
    """kvc_set_attention_v2_mode: partition kernel of paged attention v2: 0 auto, 1 mma.sync, 2 tcgen05 / TMEM."""
    _check(native.load_lib().kvc_set_attention_v2_mode(int(mode)))


def get_attention_v2_mode() -> int:
    return int(native.load_lib().kvc_get_attention_v2_mode())


def set_select_mode(mode: int) -> None:
    """kvc_set_select_mode: 0 = choose by shape (default), 1 = rank / threshold / emit launches,
    2 = one CTA per sequence with the pruned sort whenever block size and context lengths allow,
    3 = like 0 but the long-context radix path always runs as seven launches (never the cooperative one)."""
    _check(native.load_lib().kvc_set_select_mode(int(mode)))


def get_select_mode() -> int:
    return int(native.load_lib().kvc_get_select_mode())


def set_moves_mode(mode: int) -> None:
    """kvc_set_moves_mode: 0 = register-gather kernels (default), 1 = TMA-staged kernel for post-prefill
    compaction (bit-identical results)."""
    _check(native.load_lib().kvc_set_moves_mode(int(mode)))


def get_moves_mode() -> int:
    return int(native.load_lib().kvc_get_moves_mode())


class DeviceBlockAllocator:
    """The allocator / block-state ops of include/kvc_b200.h (kvc_alloc_*) bound to one set of state
    tensors -- the same tensors the reference keeps in ParallelBlockAllocator, BlockState and
    CompressionMetrics (block_manager.py:75-140, block.py:94-126, metrics.py:220-275)."""

    def __init__(self, num_blocks, block_tables, context_lens, kv_position, block_size, metadata=None):
        dev = block_tables.device
        self.NB, self.B = int(num_blocks), int(block_size)
        self.block_tables, self.context_lens, self.kv_position = block_tables, context_lens, kv_position
        self.L, self.MS, self.H, self.MB = block_tables.shape
        self.free_mask = torch.empty(self.NB, dtype=torch.bool, device=dev)
        self.free_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        # (seq, layer, head, logical block) index by block, or None
        self.metadata = metadata if metadata is not None else (None, None, None, None)
        lib = native.load_lib()
        lib.kvc_alloc_workspace_bytes.restype = ctypes.c_size_t
        nbytes = lib.kvc_alloc_workspace_bytes(ctypes.c_int64(self.NB), self.MS, self.L, self.H)
        self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _check(lib.kvc_alloc_init(_ptr(self.free_mask), _ptr(self.free_count), _ptr(self.status),
                                  ctypes.c_int64(self.NB), _stream()))

    def _state(self):
        return (_ptr(self.free_mask), _ptr(self.free_count), _ptr(self.status), ctypes.c_int64(self.NB),
                _ptr(_i32(self.block_tables, "block_tables")), _ptr(_i32(self.context_lens, "context_lens")))

    def _meta(self):
        return [_ptr(t) for t in self.metadata]

    def _ws(self):
        return _ptr(self.ws), ctypes.c_size_t(self.ws.numel())

    def append_slots(self, seq_slots, seq_len, slot_mapping):
        n = seq_len.numel()
        assert slot_mapping.dtype == torch.int64 and slot_mapping.is_contiguous()
        _check(native.load_lib().kvc_alloc_append_slots(
            *self._state(), _ptr(None if seq_slots is None else _i32(seq_slots, "seq_slots")),
            _ptr(_i32(seq_len, "seq_len")), _ptr(slot_mapping), _ptr(_i32(self.kv_position, "kv_position")),
            *self._meta(), n, self.MS, self.L, self.H, self.MB, self.B, *self._ws(), _stream()))

    def add_sequences(self, seq_slots, seq_lens, max_seq_len):
        _check(native.load_lib().kvc_alloc_add_sequences(
            *self._state(), _ptr(_i32(seq_slots, "seq_slots")), _ptr(_i32(seq_lens, "seq_lens")),
            _ptr(_i32(self.kv_position, "kv_position")), *self._meta(), seq_lens.numel(), int(max_seq_len),
            self.MS, self.L, self.H, self.MB, self.B, *self._ws(), _stream()))

    def free_compressed(self, seq_slots, evicted_block_count):
        n = evicted_block_count.shape[0]
        _check(native.load_lib().kvc_alloc_free_compressed(
            _ptr(self.free_mask), _ptr(self.free_count), _ptr(self.block_tables), _ptr(self.context_lens),
            _ptr(_i32(evicted_block_count, "evicted_block_count")),
            _ptr(None if seq_slots is None else _i32(seq_slots, "seq_slots")), self._meta()[0], n, self.MS,
            self.L, self.H, self.MB, self.B, _stream()))

    def remove_sequences(self, seq_slots):
        _check(native.load_lib().kvc_alloc_remove_sequences(
            _ptr(self.free_mask), _ptr(self.free_count), _ptr(self.block_tables), _ptr(self.context_lens),
            _ptr(_i32(seq_slots, "seq_slots")), self._meta()[0], seq_slots.numel(), self.MS, self.L, self.H,
            self.MB, self.B, _stream()))
