"""Python op shim with the reference's names, argument order and pre-processing.

Mirror of the KV-Compress part of the reference's vllm/_custom_ops.py
(paged_attention_kvc_v1/v2 :135-203, reshape_and_cache_kvc :641-658,
schedule_cache_evictions :934-1006, count_block_evictions :1065-1086,
schedule_cache_moves :1158-1179, execute_cache_moves :1220-1256): callers in
vllm/kvcompress/{metrics,scheduler}.py, vllm/worker/cache_engine.py and
vllm/attention/ops/paged_attn.py can import these instead.  The ops resolve to
torch.ops._C / _C_cache_ops / _C_kvc_ops registered by this package's _C.so, i.e. to
the sm_100a kernels; there is no Python / CPU fallback (the reference keeps slow
``ref_*`` Python twins next to each wrapper -- here those live in oracle/ and are
test-only).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import native

native.load_torch_ops()

MAX_INT = 2147483000  # vllm/_custom_ops.py:773


def paged_attention_kvc_v1(out, kv_metric_out, query, key_cache, value_cache, num_kv_heads: int,
                           scale: float, block_tables, context_lens, kv_position, last_position,
                           kv_metric_buffer_len, block_size: int, max_context_len: int,
                           alibi_slopes: Optional[torch.Tensor], kv_cache_dtype: str,
                           k_scale: float, v_scale: float, record_kv_metrics: bool) -> None:
    """reference vllm/_custom_ops.py:135-163"""
    torch.ops._C.kvcompress_paged_attention_v1(
        out, kv_metric_out, query, key_cache, value_cache, num_kv_heads, scale, block_tables,
        context_lens, kv_position, last_position, kv_metric_buffer_len, block_size,
        max_context_len, alibi_slopes, kv_cache_dtype, k_scale, v_scale, record_kv_metrics)


def paged_attention_kvc_v2(out, kv_metric_out, exp_sum, max_logits, tmp_out, tmp_kv_metric_out,
                           query, key_cache, value_cache, num_kv_heads: int, scale: float,
                           block_tables, context_lens, kv_position, last_position,
                           kv_metric_buffer_len, block_size: int, max_context_len: int,
                           alibi_slopes: Optional[torch.Tensor], kv_cache_dtype: str,
                           k_scale: float, v_scale: float, record_kv_metrics: bool) -> None:
    """reference vllm/_custom_ops.py:166-203"""
    torch.ops._C.kvcompress_paged_attention_v2(
        out, kv_metric_out, exp_sum, max_logits, tmp_out, tmp_kv_metric_out, query, key_cache,
        value_cache, num_kv_heads, scale, block_tables, context_lens, kv_position, last_position,
        kv_metric_buffer_len, block_size, max_context_len, alibi_slopes, kv_cache_dtype, k_scale,
        v_scale, record_kv_metrics)


def paged_attention_kvc_fused_metrics(out, kv_metrics, query, key_cache, value_cache,
                                      num_kv_heads: int, scale: float, block_tables, context_lens,
                                      kv_position, last_position, kv_metric_buffer_len,
                                      block_size: int, max_context_len: int,
                                      alibi_slopes: Optional[torch.Tensor] = None,
                                      kv_cache_dtype: str = "auto") -> None:
    """ADDITIVE op: attention + `metrics += sum_q p^2` in one launch; replaces
    clear_temp_metrics / record / aggregate_decode (vllm/kvcompress/metrics.py:337-342,
    429-439; engine call sites llm_engine.py:1621-1634)."""
    torch.ops._C.kvcompress_paged_attention_fused_metrics(
        out, kv_metrics, query, key_cache, value_cache, num_kv_heads, scale, block_tables,
        context_lens, kv_position, last_position, kv_metric_buffer_len, block_size,
        max_context_len, alibi_slopes, kv_cache_dtype)


def reshape_and_cache_kvc(key, value, key_cache, value_cache, kv_metrics, slot_mapping,
                          kv_metric_head_bias, kv_cache_dtype: str, k_scale: float,
                          v_scale: float) -> None:
    """reference vllm/_custom_ops.py:641-658"""
    torch.ops._C_cache_ops.kvcompress_reshape_and_cache(
        key, value, key_cache, value_cache, kv_metrics, slot_mapping, kv_metric_head_bias,
        kv_cache_dtype, k_scale, v_scale)


def schedule_cache_evictions(out_evicted_kv_indices, out_evicted_logical_indices,
                             out_evicted_kv_count, out_evicted_kv_offsets, sorted_indices,
                             seq_block_offsets, layer_by_block, head_by_block,
                             logical_block_num_by_block, evicted_blocks_per_seq, context_lens,
                             hanging_token_count, kv_position, last_position,
                             protected_window_size, block_size: int, max_evicted_kv: int,
                             null_eviction_index: int, truncate: bool,
                             evict_evenly_per_layer: bool = False,
                             control_layers: Optional[torch.Tensor] = None) -> None:
    """reference vllm/_custom_ops.py:934-1006 (same pre-processing: zero the head-index
    buffer, clone hanging counts as the scratch counter, transpose context_lens [L,S,H] ->
    [S,L,H], cast index tensors to int32)."""
    if evict_evenly_per_layer and block_size > 1:
        raise RuntimeError(
            f"cannot evict evenly across layers when block_size > 1 (got {block_size=})")
    out_evicted_kv_indices.zero_()
    torch.ops._C_kvc_ops.schedule_cache_evictions(
        out_evicted_kv_indices, out_evicted_logical_indices, out_evicted_kv_count,
        hanging_token_count.clone(), out_evicted_kv_offsets,
        sorted_indices.type(torch.int).contiguous(), seq_block_offsets.type(torch.int).contiguous(),
        layer_by_block.contiguous(), head_by_block.contiguous(),
        logical_block_num_by_block.contiguous(), evicted_blocks_per_seq.contiguous(),
        context_lens.transpose(0, 1).contiguous(), hanging_token_count.contiguous(),
        kv_position.type(torch.int).contiguous(), last_position.type(torch.int),
        protected_window_size, block_size, evict_evenly_per_layer, control_layers, max_evicted_kv,
        null_eviction_index, truncate)


def truncate_cache_evictions(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                             evicted_kv_offsets, hanging_token_count, block_size: int,
                             max_evicted_kv: int, null_eviction_index: int) -> None:
    """op csrc/torch_bindings.cpp:389-394 (the reference has no Python wrapper for it)"""
    torch.ops._C_kvc_ops.truncate_cache_evictions(
        evicted_kv_indices, evicted_logical_indices, evicted_kv_count, evicted_kv_offsets,
        hanging_token_count, block_size, max_evicted_kv, null_eviction_index)


def count_block_evictions(evicted_block_count, evicted_logical_indices, evicted_kv_offsets,
                          hanging_token_count, block_size: int, null_value: int,
                          evicted_blocks_per_seq: Optional[torch.Tensor] = None) -> None:
    """reference vllm/_custom_ops.py:1065-1086.  NB: like the reference, a non-contiguous
    evicted_logical_indices would be updated on a temporary copy."""
    torch.ops._C_kvc_ops.count_block_evictions(
        evicted_block_count, evicted_logical_indices.contiguous(), evicted_kv_offsets.contiguous(),
        hanging_token_count.contiguous(), block_size, null_value)


def schedule_cache_moves(out_cache_moves_indices, out_cache_moves_count, evicted_logical_indices,
                         evicted_kv_count, evicted_kv_offsets, block_tables, context_lens,
                         block_size: int) -> None:
    """reference vllm/_custom_ops.py:1158-1179 (zero-fills the move workspace first)"""
    out_cache_moves_indices.fill_(0)
    torch.ops._C_kvc_ops.schedule_t1_cache_moves(
        out_cache_moves_indices, out_cache_moves_count, evicted_logical_indices.contiguous(),
        evicted_kv_count.contiguous(), evicted_kv_offsets.contiguous(), block_tables.contiguous(),
        context_lens.contiguous(), block_size)


def schedule_cache_moves_nofill(out_cache_moves_indices, out_cache_moves_count, evicted_logical_indices,
                                evicted_kv_count, evicted_kv_offsets, block_tables, context_lens,
                                block_size: int) -> None:
    """The op itself, without the reference wrapper's zero-fill of the whole [max_kv, 2] workspace
    (vllm/_custom_ops.py:1168; only entries [0, count) of each head segment are ever read)."""
    torch.ops._C_kvc_ops.schedule_t1_cache_moves(
        out_cache_moves_indices, out_cache_moves_count, evicted_logical_indices, evicted_kv_count,
        evicted_kv_offsets, block_tables, context_lens, block_size)


def execute_cache_moves(k_cache, v_cache, kv_metrics, kv_position, cache_moves_indices,
                        cache_moves_count, evicted_kv_offsets, blocks_per_head: int,
                        threads_per_head: int) -> None:
    """reference vllm/_custom_ops.py:1220-1256"""
    torch.ops._C_kvc_ops.execute_cache_moves(
        k_cache, v_cache, kv_metrics, kv_position, cache_moves_indices.contiguous(),
        cache_moves_count.contiguous(), evicted_kv_offsets.contiguous(), blocks_per_head,
        threads_per_head)


def execute_cache_moves_deferred(k_cache, v_cache, kv_metrics, kv_position, cache_moves_indices,
                                 cache_moves_count, evicted_kv_offsets, pend_kv, pend_slot,
                                 pend_dst: Optional[torch.Tensor] = None) -> None:
    """ADDITIVE op: execute_cache_moves that also resolves the deferred appends of
    paged_attention_kvc_fused_step_deferred (pend_kv [L,S,H,2D], pend_slot [L,S,H] int64): a move
    whose source is a pending slot copies from the compact buffer, a pending token no move consumes
    is stored to its own slot.  The cache ends up bit-identical to reshape_and_cache_kvc +
    execute_cache_moves.  With pend_dst [L,S,H] int64 (initialised to -1) a head whose only move takes
    its parked token only gets metric / position moved and the destination recorded there: the next
    paged_attention_kvc_fused_step_deferred stores the K/V (deferred move); flush_deferred resolves
    whatever is outstanding."""
    torch.ops._C_kvc_ops.execute_cache_moves_deferred(
        k_cache, v_cache, kv_metrics, kv_position, cache_moves_indices.contiguous(),
        cache_moves_count.contiguous(), evicted_kv_offsets.contiguous(), pend_kv, pend_slot, pend_dst)


def flush_deferred(k_cache, v_cache, pend_kv, pend_slot, pend_dst: Optional[torch.Tensor] = None) -> None:
    """ADDITIVE op: store every parked token where it belongs (pend_dst if set, else pend_slot) and clear
    the markers; afterwards the cache is what the reference's op sequence leaves."""
    torch.ops._C_kvc_ops.flush_deferred(k_cache, v_cache, pend_kv, pend_slot, pend_dst)


SELECT_EVICTIONS_LAUNCHES = 3  # rank, threshold, emit


def select_evictions(out_evicted_logical_indices, out_evicted_kv_count, out_evicted_block_count,
                     kv_metrics, kv_position, block_tables, context_lens, hanging_token_count,
                     evicted_kv_offsets, evicted_blocks_per_seq, last_position, protected_window_size,
                     block_size: int, null_eviction_index: int,
                     head_bias: Optional[torch.Tensor] = None, position_bins: Optional[torch.Tensor] = None,
                     bias_weight: float = 0.0, use_average: bool = False, num_sinks: int = 0) -> None:
    """ADDITIVE op: fused replacement of CompressionMetrics.schedule_evictions
    (vllm/kvcompress/metrics.py:441-847) / the sort_seq_metrics -> schedule_cache_evictions ->
    re-sort chain (metrics.py:849-966, scheduler.py:285-486).  Same outputs as
    schedule_evictions returns: (evicted_logical_indices, evicted_kv_count, evicted_block_count).
    head_bias [L, H, bins] / position_bins [bins] / bias_weight = KVHeadBias and
    kv_metric_bias_weight (metrics.py:45-81, 503-506); use_average (:495-501); num_sinks (:542)."""
    torch.ops._C_kvc_ops.select_evictions(
        out_evicted_logical_indices, out_evicted_kv_count, out_evicted_block_count, kv_metrics,
        kv_position.contiguous(), block_tables.contiguous(), context_lens.contiguous(),
        hanging_token_count.contiguous(), evicted_kv_offsets.contiguous(),
        evicted_blocks_per_seq.contiguous(), last_position.type(torch.int), protected_window_size,
        block_size, null_eviction_index, head_bias,
        None if position_bins is None else position_bins.type(torch.int).contiguous(), bias_weight,
        use_average, num_sinks)


def prepare_compression(context_lens, out_hanging_token_count, out_evicted_kv_offsets,
                        out_evicted_blocks_per_seq, block_size: int, max_cache_blocks_per_seq: int) -> None:
    """ADDITIVE op: hanging counts, segment offsets and per-sequence block budgets in two launches
    (reference: block.py:330-335, scheduler.py:131-141, 274-280 as torch ops)."""
    torch.ops._C_kvc_ops.prepare_compression(context_lens, out_hanging_token_count, out_evicted_kv_offsets,
                                             out_evicted_blocks_per_seq, block_size, max_cache_blocks_per_seq)


def free_and_append_slots(context_lens, block_tables, evicted_block_count, hanging_token_count, seq_len,
                          out_slot_mapping, kv_position, block_size: int, append: bool = True,
                          status=None) -> None:
    """ADDITIVE op: free the compressed trailing blocks (block.py:184-210) and append one token per head
    (block_manager.py:269-294, block.py:275-321) in one launch.  status (int32 [1], optional) is set to 2
    when a head's block table is full (the reference raises an index error there)."""
    torch.ops._C_kvc_ops.free_and_append_slots(context_lens, block_tables, evicted_block_count,
                                               hanging_token_count, seq_len, out_slot_mapping, kv_position,
                                               block_size, append, status)


def paged_attention_kvc_fused_step(out, kv_metrics, query, key, value, key_cache, value_cache, slot_mapping,
                                   kv_metric_head_bias, num_kv_heads: int, scale: float, block_tables,
                                   context_lens, kv_position, last_position, kv_metric_buffer_len,
                                   block_size: int, max_context_len: int,
                                   alibi_slopes: Optional[torch.Tensor] = None,
                                   kv_cache_dtype: str = "auto") -> None:
    """ADDITIVE op: one decode layer of the hot path in one launch = reshape_and_cache_kvc (cache
    write + metric init, flash_attn.py:871) + paged attention + aggregate_decode."""
    torch.ops._C.kvcompress_paged_attention_fused_step(
        out, kv_metrics, query, key, value, key_cache, value_cache, slot_mapping, kv_metric_head_bias,
        num_kv_heads, scale, block_tables, context_lens, kv_position, last_position, kv_metric_buffer_len,
        block_size, max_context_len, alibi_slopes, kv_cache_dtype)


def paged_attention_kvc_fused_step_deferred(out, kv_metrics, query, key, value, key_cache, value_cache,
                                            slot_mapping, kv_metric_head_bias, pend_kv, pend_slot, pend_dst,
                                            num_kv_heads: int, scale: float, block_tables, context_lens,
                                            kv_position, last_position, kv_metric_buffer_len, block_size: int,
                                            max_context_len: int, alibi_slopes: Optional[torch.Tensor] = None,
                                            kv_cache_dtype: str = "auto") -> None:
    """ADDITIVE op: paged_attention_kvc_fused_step with deferred appends: a new token that opens a
    fresh block is parked in pend_kv [S,H,2D] / pend_slot [S,H] (this layer's slices) instead of being
    stored to the cache; execute_cache_moves_deferred of the next step resolves it.  pend_dst [S,H] (or
    None): destinations of deferred moves recorded by execute_cache_moves_deferred, applied and reset here."""
    torch.ops._C.kvcompress_paged_attention_fused_step_deferred(
        out, kv_metrics, query, key, value, key_cache, value_cache, slot_mapping, kv_metric_head_bias,
        pend_kv, pend_slot, pend_dst, num_kv_heads, scale, block_tables, context_lens, kv_position, last_position,
        kv_metric_buffer_len, block_size, max_context_len, alibi_slopes, kv_cache_dtype)


def prefill_kv_metrics(query, key, seq_lens, scale: float, kv_metric_buffer_len, n_observed: int = 32,
                       use_l2: bool = True, use_average: bool = False, use_maxpool: bool = True,
                       kv_metrics: Optional[torch.Tensor] = None,
                       slot_mapping: Optional[torch.Tensor] = None,
                       return_metrics: bool = True,
                       seq_start_loc: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """ADDITIVE op: tensor-core (tcgen05) replacement of _naive_kvc_attention
    (vllm/attention/backends/flash_attn.py:1122-1211; same argument meaning, `seq_lens` is the
    Python list prefill_meta.seq_lens, `key` is the un-repeated [T, num_kv_heads, D] tensor).
    Returns kv_metric_out [T, num_heads] f32 like the reference; when kv_metrics / slot_mapping
    are given, CompressionMetrics.aggregate_prefill (metrics.py:395-427) is fused in.
    `seq_start_loc` (int32 [num_seqs + 1], on the device): the cumulative prompt starts the reference's prefill
    metadata already holds (flash_attn.py: prefill_meta.seq_start_loc); built from `seq_lens` when absent, which
    costs a host-to-device copy per call."""
    if seq_start_loc is None:
        starts = [0]
        for n in seq_lens:
            starts.append(starts[-1] + int(n))
        assert starts[-1] == query.shape[0], "seq_lens must cover the prefill tokens"
        seq_start_loc = torch.tensor(starts, dtype=torch.int32, device=query.device)
    else:
        assert seq_start_loc.dtype == torch.int32 and seq_start_loc.numel() == len(seq_lens) + 1
    out = (torch.empty((query.shape[0], query.shape[1]), dtype=torch.float32, device=query.device)
           if return_metrics else None)
    torch.ops._C_kvc_ops.prefill_kv_metrics(
        out, kv_metrics, None if slot_mapping is None else slot_mapping.contiguous(), query, key,
        seq_start_loc, kv_metric_buffer_len.type(torch.int).contiguous(), max(seq_lens) if seq_lens else 0,
        scale, n_observed, use_l2, use_average, use_maxpool)
    return out


def aggregate_prefill(kv_metrics, prefill_metrics, slot_mapping) -> None:
    """ADDITIVE op: CompressionMetrics.aggregate_prefill (vllm/kvcompress/metrics.py:395-427) as one
    launch: kv_metrics[slot_mapping[t, h]] += prefill_metrics[t, h*Q:(h+1)*Q].sum()."""
    torch.ops._C_kvc_ops.aggregate_prefill(kv_metrics, prefill_metrics.contiguous(), slot_mapping.contiguous())


# ---- ADDITIVE ops: device-side block allocator / block-state updates (include/kvc_b200.h kvc_alloc_*) ----
# They operate IN PLACE on the reference's own state tensors: ParallelBlockAllocator.free_mask
# (block_manager.py:101-103), BlockState.block_tables / context_lens (block.py:94-126) and the
# CompressionMetrics metadata arrays (metrics.py:220-275); free_count / status are int32 [1] tensors.
def alloc_init(free_mask, free_count, status) -> None:
    """ParallelBlockAllocator.free_all (block_manager.py:127-129)."""
    torch.ops._C_kvc_ops.alloc_init(free_mask, free_count, status)


def alloc_append_slots(free_mask, free_count, status, block_tables, context_lens, seq_slots, seq_len,
                       out_slot_mapping, token_positions, block_size: int, seq_index_by_block=None,
                       layer_index_by_block=None, head_index_by_block=None,
                       logical_block_num_by_block=None) -> None:
    """BlockSpaceManagerKVC._append_to_sequence_batch (block_manager.py:269-294) + insert_metadata
    (metrics.py:344-362) + BlockStateView.get_decode_slot_mapping (block.py:303-321) in 3 launches."""
    torch.ops._C_kvc_ops.alloc_append_slots(free_mask, free_count, status, block_tables, context_lens, seq_slots,
                                            seq_len, out_slot_mapping, token_positions, seq_index_by_block,
                                            layer_index_by_block, head_index_by_block,
                                            logical_block_num_by_block, block_size)


def alloc_add_sequences(free_mask, free_count, status, block_tables, context_lens, seq_slots, seq_lens,
                        max_seq_len: int, token_positions, block_size: int, seq_index_by_block=None,
                        layer_index_by_block=None, head_index_by_block=None,
                        logical_block_num_by_block=None) -> None:
    """BlockSpaceManagerKVC._add_sequence (block_manager.py:196-222) for a batch of prompts."""
    torch.ops._C_kvc_ops.alloc_add_sequences(free_mask, free_count, status, block_tables, context_lens, seq_slots,
                                             seq_lens, max_seq_len, token_positions, seq_index_by_block,
                                             layer_index_by_block, head_index_by_block,
                                             logical_block_num_by_block, block_size)


def alloc_free_compressed(free_mask, free_count, block_tables, context_lens, evicted_block_count, seq_slots,
                          block_size: int, seq_index_by_block=None) -> None:
    """BlockSpaceManagerKVC.free_compressed_blocks (block_manager.py:466-530) + remove_trailing_blocks
    (block.py:184-210) + remove_metadata (metrics.py:364-370)."""
    torch.ops._C_kvc_ops.alloc_free_compressed(free_mask, free_count, block_tables, context_lens,
                                               evicted_block_count.contiguous(), seq_slots, seq_index_by_block,
                                               block_size)


def alloc_remove_sequences(free_mask, free_count, block_tables, context_lens, seq_slots, block_size: int,
                           seq_index_by_block=None) -> None:
    """BlockSpaceManagerKVC._remove_sequence_batch (block_manager.py:243-262)."""
    torch.ops._C_kvc_ops.alloc_remove_sequences(free_mask, free_count, block_tables, context_lens, seq_slots,
                                                seq_index_by_block, block_size)
