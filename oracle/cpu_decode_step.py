"""CPU restatement of one decode step of the KV-Compress hot path -- TEST/BASELINE
INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs, tests).

Same cycle as vllm_kvcompress_b200.engine.DecodeHotPath.step, built from the C oracle
(oracle/kvc_oracle.c, OpenMP over heads/sequences) and numpy sorts in the order the
reference engine runs them (vllm/engine/llm_engine.py:1556-1634):
  sort_seq_metrics (metrics.py:849-966, numpy argsort) -> schedule_cache_evictions + truncate
  -> host re-sort (scheduler.py:472-477) -> schedule_t1_cache_moves -> execute_cache_moves
  -> free blocks -> append slots -> per layer reshape_and_cache + paged attention with
  recorded probabilities -> aggregate_decode.
The reference has no CPU build of these ops (csrc/cpu registers none of them, SURVEY.md
section 2.2), so this port is the CPU baseline ("kind": "port").
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import kvc_oracle as O

MAX_INT = O.MAX_INT


class CpuStepConfig:
    """The shape fields CpuDecodeHotPath reads (same names as vllm_kvcompress_b200.engine.HotPathConfig, so either
    can be passed); lets bench.py's CPU arm run without importing the product's engine module."""

    def __init__(self, num_seqs=256, num_layers=32, num_kv_heads=8, queries_per_kv=4, head_size=128, block_size=16,
                 max_cache_tokens=128, protected_window=32, prompt_len=6000, blocks_per_head=12):
        self.num_seqs, self.num_layers, self.num_kv_heads = num_seqs, num_layers, num_kv_heads
        self.queries_per_kv, self.head_size, self.block_size = queries_per_kv, head_size, block_size
        self.max_cache_tokens, self.protected_window = max_cache_tokens, protected_window
        self.prompt_len, self.blocks_per_head = prompt_len, blocks_per_head


class CpuDecodeHotPath:
    def __init__(self, cfg, seed=0):
        """cfg: vllm_kvcompress_b200.engine.HotPathConfig (num_seqs = the CPU sample size)."""
        from vllm_kvcompress_b200 import workload as W  # tensor plumbing only, no kernels
        self.cfg = cfg
        S, L, H, B, D = cfg.num_seqs, cfg.num_layers, cfg.num_kv_heads, cfg.block_size, cfg.head_size
        st = W.make_state(S, L, H, B, D, cfg.max_cache_tokens, dtype=torch.float16, device="cpu", seed=seed,
                          slack=1.0, max_blocks=cfg.blocks_per_head, fill_cache=True,
                          blocks_per_head_alloc=cfg.blocks_per_head)
        self.W = W
        self.st = st
        n = cfg.max_cache_tokens
        heads = L * S * H
        rng = np.random.RandomState(seed + 7)
        slots = (st.block_tables.view(heads, -1)[:, : n // B].long()[:, :, None] * B
                 + torch.arange(B)[None, None]).view(heads, n).numpy()
        pos = np.sort(rng.randint(0, cfg.prompt_len - cfg.protected_window, size=(heads, n - cfg.protected_window)), 1)
        pos = np.concatenate([pos, np.broadcast_to(np.arange(cfg.prompt_len - cfg.protected_window, cfg.prompt_len),
                                                   (heads, cfg.protected_window))], 1).astype(np.int32)
        kvp = np.full(st.NB * B, -1, np.int32)
        kvp[slots.reshape(-1)] = pos.reshape(-1)
        self.kv_position = kvp.reshape(st.NB, B)
        self.kv_metrics = rng.uniform(0.01, 1.0, size=(st.NB, B)).astype(np.float32)
        self.k_cache = st.k_cache.view(torch.int16).numpy().view(np.uint16)
        self.v_cache = st.v_cache.view(torch.int16).numpy().view(np.uint16)
        self.block_tables = st.block_tables.numpy()
        self.context_lens = st.context_lens.numpy().copy()          # [L,S,H]
        self.seq_len = np.full(S, cfg.prompt_len, np.int32)
        self.protected = np.full(S, cfg.protected_window, np.int32)
        self.temp = np.zeros((st.NB, B, cfg.queries_per_kv), np.float32)
        self.bias = np.zeros(H, np.float32)

    def compress(self):
        cfg, st, W = self.cfg, self.st, self.W
        S, L, H, B = cfg.num_seqs, cfg.num_layers, cfg.num_kv_heads, cfg.block_size
        ctx = self.context_lens
        nblk = (ctx + B - 1) // B
        hanging = np.where(ctx % B == 0, B, ctx % B).transpose(1, 0, 2).astype(np.int32).copy()
        alloc = (nblk.transpose(1, 0, 2) * B).reshape(-1)
        offs = np.concatenate([[0], np.cumsum(alloc)[:-1]]).astype(np.int32).reshape(S, L, H)
        N = int(alloc.sum())
        max_blocks = (cfg.max_cache_tokens * L * H + B - 1) // B
        budgets = np.maximum(nblk.sum(axis=(0, 2)) - max_blocks, 0).astype(np.int32)
        # sort_seq_metrics: mask allocated blocks (physical order), sort slots by (seq, metric)
        lsh = (st.layer_by_block.numpy().astype(np.int64) * S + np.maximum(st.seq_by_block.numpy(), 0)) * H \
            + st.head_by_block.numpy()
        mask = (st.seq_by_block.numpy() >= 0) & (st.logical_by_block.numpy() < nblk.reshape(-1)[lsh])
        blocks = np.nonzero(mask)[0]
        seqs = np.repeat(st.seq_by_block.numpy()[blocks], B)
        order = np.lexsort((self.kv_metrics[blocks].reshape(-1), seqs)).astype(np.int32)
        counts = np.bincount(st.seq_by_block.numpy()[blocks], minlength=S)
        seq_off = (np.cumsum(counts) - counts).astype(np.int32)
        ekv, elog = np.zeros(N, np.int32), np.zeros(N, np.int32)
        cnt = np.zeros((S, L, H), np.int32)
        O.schedule_cache_evictions(ekv, elog, cnt, hanging.copy(), offs, order, seq_off,
                                   st.layer_by_block.numpy()[blocks], st.head_by_block.numpy()[blocks],
                                   st.logical_by_block.numpy()[blocks], budgets,
                                   np.ascontiguousarray(ctx.transpose(1, 0, 2)), hanging,
                                   np.ascontiguousarray(self.kv_position[blocks]), self.seq_len - 1,
                                   self.protected, B, False, None, N, MAX_INT, True)
        logical = O.sort_evictions_by_head(ekv, elog)
        moves = np.zeros((N, 2), np.int32)
        mcnt = np.zeros((S, L, H), np.int32)
        O.schedule_cache_moves(moves, mcnt, logical, cnt, offs, self.block_tables,
                               np.ascontiguousarray(ctx), B)
        O.execute_cache_moves(self.k_cache, self.v_cache, self.kv_metrics, self.kv_position, moves, mcnt, offs)
        nb_ev = np.where(cnt >= hanging, (cnt - hanging) // B + 1, 0)
        freed = np.maximum(nb_ev * B - (B - hanging), 0)
        self.context_lens = (ctx - freed.transpose(1, 0, 2)).astype(np.int32)
        self.last_moves = int(mcnt.sum())

    def append_slots(self):
        B = self.cfg.block_size
        ctx = self.context_lens
        blk = np.take_along_axis(self.block_tables, (ctx // B)[..., None].astype(np.int64), axis=3)[..., 0]
        self.slot_mapping = blk.astype(np.int64) * B + ctx % B
        self.kv_position.reshape(-1)[self.slot_mapping.reshape(-1)] = np.broadcast_to(
            self.seq_len[None, :, None], ctx.shape).reshape(-1)
        self.context_lens = ctx + 1
        self.seq_len = self.seq_len + 1

    def forward(self, qkv: np.ndarray, out: np.ndarray):
        """qkv [L, S, (Hq+2H)*D] uint16 (fp16 bits); out [L, S, Hq*D] float32."""
        cfg = self.cfg
        H, Q, D = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size
        Hq = H * Q
        S = cfg.num_seqs
        zeros = np.zeros(S, np.int32)
        self.temp[:] = 0  # clear_temp_metrics, metrics.py:337-342
        for l in range(cfg.num_layers):
            x = qkv[l]
            q = np.ascontiguousarray(x[:, : Hq * D]).reshape(S, Hq, D)
            k = np.ascontiguousarray(x[:, Hq * D: (Hq + H) * D]).reshape(S, H, D)
            v = np.ascontiguousarray(x[:, (Hq + H) * D:]).reshape(S, H, D)
            O.reshape_and_cache_kvc(k, v, self.k_cache, self.v_cache, self.kv_metrics,
                                    self.slot_mapping[l], self.bias)
            o = out[l].reshape(S, Hq, D)
            O.paged_attention_kvc(o, self.temp, q, self.k_cache, self.v_cache, H, D ** -0.5,
                                  np.ascontiguousarray(self.block_tables[l]),
                                  np.ascontiguousarray(self.context_lens[l]), self.kv_position,
                                  self.seq_len - 1, zeros, None, False, True)
        O.aggregate_decode(self.kv_metrics, self.temp)  # metrics.py:429-439

    def step(self, qkv, out):
        self.compress()
        self.append_slots()
        self.forward(qkv, out)
