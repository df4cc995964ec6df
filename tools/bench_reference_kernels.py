#!/usr/bin/env python
"""Supplementary measurement (NOT part of bench.py's contract): the same decode hot-path step driven
through the REFERENCE's own CUDA kernels, compiled unmodified for sm_100a into oracle/_ref/_kvc_ref.so,
in the order the fork's engine runs them (V1 selection chain: torch.sort -> schedule_cache_evictions ->
host re-sort -> schedule_t1_cache_moves -> execute_cache_moves; per layer kvcompress_reshape_and_cache +
kvcompress_paged_attention_v1 with temp metrics; clear_temp_metrics + aggregate_decode as torch ops).
Prints one JSON line.  Run on a B200: python tools/bench_reference_kernels.py [--batch 256] [--steps 5]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


class RefOps:
    """vllm/_custom_ops.py wrappers bound to the reference kernels (namespace _kvc_ref)."""
    MAX_INT = 2147483000
    SELECT_EVICTIONS_LAUNCHES = 0

    def __init__(self):
        torch.ops.load_library(os.path.join(ROOT, "oracle", "_ref", "_kvc_ref.so"))
        self.r = torch.ops._kvc_ref

    def reshape_and_cache_kvc(self, *a):
        self.r.kvcompress_reshape_and_cache(*a)

    def paged_attention_kvc_v1(self, *a):
        self.r.kvcompress_paged_attention_v1(*a)

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

    def schedule_cache_moves(self, moves, mcnt, logical, cnt, offs, block_tables, context_lens, block_size):
        moves.fill_(0)
        self.r.schedule_t1_cache_moves(moves, mcnt, logical.contiguous(), cnt.contiguous(), offs.contiguous(),
                                       block_tables.contiguous(), context_lens.contiguous(), block_size)

    def execute_cache_moves(self, *a):
        self.r.execute_cache_moves(*a)

    def free_and_append_slots(self, *a):
        from vllm_kvcompress_b200 import _custom_ops as ours
        ours.free_and_append_slots(*a)   # host-side block-state bookkeeping, not a reference kernel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    from vllm_kvcompress_b200 import engine as E
    ref = RefOps()
    E.ops = ref  # the engine module resolves its ops through this name
    E.torch.ops._C_kvc_ops.schedule_t1_cache_moves = ref.r.schedule_t1_cache_moves
    cfg = E.HotPathConfig(num_seqs=args.batch, selection="chain", fused_metrics=False, fused_cache_write=False)
    dev = torch.device("cuda", 0)
    eng = E.DecodeHotPath(cfg, dev)
    H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
    qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), device=dev) - 0.5).to(cfg.dtype)
    out = torch.empty((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    for _ in range(args.warmup):
        eng.step(qkv, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.step(qkv, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "reference CUDA kernels (sm_100a rebuild) + torch.sort chain, eager",
                      "metric": "decode tokens/sec at CR=64x (Llama-3-8B shape, KV-Compress hot path)",
                      "value": S / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms, "batch": S}))


if __name__ == "__main__":
    main()
