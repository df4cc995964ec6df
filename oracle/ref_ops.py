"""TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by vllm-kvcompress_b200/).

The vllm/_custom_ops.py function surface bound to the REFERENCE's own CUDA kernels, compiled unmodified for
sm_100a into oracle/_ref/_kvc_ref.so (oracle/Makefile.ref; torch library namespace `_kvc_ref`).  Handing an
instance to vllm_kvcompress_b200.engine.DecodeHotPath(cfg, device, ops_module=RefOps()) runs the same decode
hot-path driver over the reference kernels in the order the fork's engine runs them (V1 selection chain:
torch.sort -> schedule_cache_evictions (K1 + K2) -> host re-sort -> schedule_t1_cache_moves (K4) ->
execute_cache_moves (K5); per layer kvcompress_reshape_and_cache (K6) + kvcompress_paged_attention_v1 (K7)
with temp metrics; clear_temp_metrics + aggregate_decode as torch ops, metrics.py:337-342, 429-439).

Used by tests/test_gpu_chain.py (multi-step parity), tools/bench_reference_kernels.py and bench.py's
`reference_cuda` leg (the fork's kernels timed beside the product, outside the product's timed region).
"""
from __future__ import annotations

import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_kvc_ref.so")


def available() -> bool:
    return os.path.exists(REF_SO)


class RefOps:
    MAX_INT = 2147483000
    SELECT_EVICTIONS_LAUNCHES = 0

    def __init__(self):
        torch.ops.load_library(REF_SO)
        self.r = torch.ops._kvc_ref

    # ---- K6 / K7 (reference csrc/kvcompress_cache_kernels.cu, csrc/attention/kvcompress_attention_kernels.cu)
    def reshape_and_cache_kvc(self, *a):
        self.r.kvcompress_reshape_and_cache(*a)

    def paged_attention_kvc_v1(self, *a):
        self.r.kvcompress_paged_attention_v1(*a)

    def paged_attention_kvc_v2(self, *a):
        self.r.kvcompress_paged_attention_v2(*a)

    # ---- K1 + K2 with the wrapper's pre-processing (vllm/_custom_ops.py:934-1006)
    def schedule_cache_evictions(self, ekv, elog, cnt, offs, sorted_indices, seq_block_offsets, layer_by_block,
                                 head_by_block, logical_by_block, budgets, context_lens, hanging, kv_position,
                                 last_position, protected, block_size, max_kv, null_idx, truncate):
        ekv.zero_()
        self.r.schedule_cache_evictions(ekv, elog, cnt, hanging.clone(), offs, sorted_indices.int().contiguous(),
                                        seq_block_offsets.int().contiguous(), layer_by_block.contiguous(),
                                        head_by_block.contiguous(), logical_by_block.contiguous(),
                                        budgets.contiguous(), context_lens.transpose(0, 1).contiguous(),
                                        hanging.contiguous(), kv_position.int().contiguous(), last_position.int(),
                                        protected, block_size, False, None, max_kv, null_idx, truncate)

    def count_block_evictions(self, evicted_block_count, evicted_logical_indices, evicted_kv_offsets, hanging,
                              block_size, null_value):
        self.r.count_block_evictions(evicted_block_count, evicted_logical_indices, evicted_kv_offsets.contiguous(),
                                     hanging.contiguous(), block_size, null_value)

    # ---- K4 (wrapper zero-fills the workspace, vllm/_custom_ops.py:1168)
    def schedule_cache_moves(self, moves, mcnt, logical, cnt, offs, block_tables, context_lens, block_size):
        moves.fill_(0)
        self.schedule_cache_moves_nofill(moves, mcnt, logical, cnt, offs, block_tables, context_lens, block_size)

    def schedule_cache_moves_nofill(self, moves, mcnt, logical, cnt, offs, block_tables, context_lens, block_size):
        self.r.schedule_t1_cache_moves(moves, mcnt, logical.contiguous(), cnt.contiguous(), offs.contiguous(),
                                       block_tables.contiguous(), context_lens.contiguous(), block_size)

    # ---- K5
    def execute_cache_moves(self, *a):
        self.r.execute_cache_moves(*a)

    # ---- block-state bookkeeping: torch formulas of block.py:184-210, 303-321 / block_manager.py:269-294 (the
    # reference has no kernel for it)
    def free_and_append_slots(self, context_lens, block_tables, evicted_block_count, hanging, seq_len,
                              slot_mapping, kv_position, block_size, append=True, status=None):
        B = block_size
        if evicted_block_count is not None:
            freed = (evicted_block_count * B - (B - hanging)).clamp_(min=0)          # [S,L,H]
            context_lens.sub_(freed.transpose(0, 1))
        if append:
            ctx = context_lens.long()
            blk = torch.gather(block_tables, 3, (ctx // B).unsqueeze(-1)).squeeze(-1).long()
            slots = blk * B + ctx % B                                                 # [L,S,H]
            slot_mapping.copy_(slots)
            kv_position.view(-1)[slots.flatten()] = seq_len.view(1, -1, 1).expand_as(ctx).flatten().int()
            context_lens.add_(1)
