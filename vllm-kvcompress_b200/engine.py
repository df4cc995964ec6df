"""Host-side driver of the KV-Compress hot path (one GPU, one independent KV cache).

This is the slice of the reference engine loop that touches the hot path
(LLMEngine.step, vllm/engine/llm_engine.py:1484-1634):

    schedule_compression  -> selection, count_block_evictions, schedule_cache_moves
                             (vllm/kvcompress/scheduler.py:184-560)
    execute_cache_moves      (vllm/worker/cache_engine.py:139-151)
    free_compressed_blocks   (vllm/kvcompress/block_manager.py:466-530: context_lens shrink)
    batch_append_slots       (block_manager.py:269-294: +1 token per head, slot mapping)
    per layer: kvcompress_reshape_and_cache + paged attention with metric recording
                             (vllm/attention/backends/flash_attn.py:871, 1067-1091)
    aggregate_decode         (vllm/kvcompress/metrics.py:429-439; fused into attention here)

for a batch of S sequences of a Llama-shaped model (L layers, H KV heads, Q query heads
per KV head, D=128, block size 16), with synthetic Q/K/V in place of the model's
projections (the dense GEMMs are outside this path).  All state lives in device tensors;
the only host<->device traffic is what the caller passes in (step_host).

Block allocation: every head owns a fixed slab of `blocks_per_head` physical blocks drawn
from a random permutation of the cache (scattered, like the reference allocator's output),
so the block table is static and "allocate / free" reduce to context_lens growing and
shrinking; the reference frees the trailing blocks and never rewrites table entries either
(block.py:184-210).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _custom_ops as ops
from . import _custom_ops as _default_ops
from . import capi
from . import workload as W

MAX_INT = W.MAX_INT


@dataclass
class HotPathConfig:
    num_seqs: int = 256
    num_layers: int = 32
    num_kv_heads: int = 8
    queries_per_kv: int = 4
    head_size: int = 128
    block_size: int = 16
    max_cache_tokens: int = 128       # benchmark_throughput.py:391-393 at input 6000, CR 64
    protected_window: int = 32
    prompt_len: int = 6000
    blocks_per_head: int = 12          # slab: max_cache_tokens/16 + headroom for variable-rate heads
    dtype: torch.dtype = torch.float16
    selection: str = "fused"           # "fused" (kvc_select kernels) | "chain" (torch.sort + K1 + K2)
    fused_metrics: bool = True         # attention accumulates sum_q p^2 itself
    fused_cache_write: bool = True     # ... and writes the step's new K/V (one launch per layer)
    # Launch mode of the attention kernels (kvc_set_attention_pdl, include/kvc_b200.h):
    #   "model" (default)  what a model can use: layer 0 follows the bookkeeping kernels that rewrite tables
    #                      and slot mapping -> mode 1 (every warp waits at entry); layers >= 1 follow a kernel
    #                      that touches neither this layer's blocks nor its tables -> mode 2 (cache blocks are
    #                      streamed at once, the wait sits in front of q / new k / new v)
    #   "independent"      layers >= 1 in mode 3: q/k/v are declared independent of the previous kernel as
    #                      well (true for this harness, whose inputs are resident before the step; never
    #                      for a model) -- A/B measurements only
    #   "safe" / "off"     mode 1 / mode 0 for every layer
    attention_pdl: str = "model"
    # "device": blocks come from / return to the device-side allocator every step (kvc_alloc_*: free the
    # compressed trailing blocks, allocate for heads crossing a block boundary, tables rewritten in place);
    # "static": every head owns a fixed slab and only context_lens moves (round-1 harness)
    allocator: str = "device"
    # deferred append: a new token that opens a fresh block (every head, every step, once continual compression
    # has reached its steady state) is parked in a compact side buffer by the fused step kernel and stored once,
    # at its final slot, by the next step's execute_cache_moves_deferred (see include/kvc_b200.h).  Needs the
    # fused cache write.  flush_pending() makes the cache canonical at any point.
    deferred_append: bool = True
    # deferred move (with deferred_append): when a head's only move takes the parked token, its K/V is stored by
    # the next fused attention step (which has the destination block staged and writes whole sectors back)
    # instead of by execute_cache_moves (partial-sector stores into a block it has to fetch for that).
    deferred_move: bool = True
    host_copy_layers: int = 4          # step_host: layers per H2D / D2H copy
    seed: int = 0


class DecodeHotPath:
    """One replica: owns the unified cache, per-head tables and metrics of `num_seqs` sequences
    already compressed to `max_cache_tokens` per head (state after a post-prefill compression)."""

    def __init__(self, cfg: HotPathConfig, device: torch.device, ops_module=None):
        """ops_module: object with the vllm/_custom_ops.py function surface (default: this package's shim over
        libkvc_b200).  Tests and the reference-kernel measurement bind the same driver to other implementations
        of that surface (oracle/ref_ops.py: the reference's own CUDA kernels)."""
        self.cfg = cfg
        self.dev = device
        self.ops = ops_module if ops_module is not None else ops
        if cfg.attention_pdl not in ("model", "independent", "safe", "off"):
            raise ValueError(f"attention_pdl: {cfg.attention_pdl!r}")
        # an explicit environment setting wins for every layer (A/B runs)
        env = os.environ.get("KVC_ATTN_PDL")
        first, rest = {"model": (1, 2), "independent": (1, 3), "safe": (1, 1), "off": (0, 0)}[cfg.attention_pdl]
        self.pdl_first, self.pdl_rest = (int(env), int(env)) if env is not None else (first, rest)
        self.attention_pdl = self.pdl_rest
        S, L, H, B, D = cfg.num_seqs, cfg.num_layers, cfg.num_kv_heads, cfg.block_size, cfg.head_size
        self.launches = 0
        if cfg.allocator == "device":
            # heads start with exactly ceil(ctx/B) scattered blocks; the rest of the cache is the free pool
            st = W.make_state(S, L, H, B, D, cfg.max_cache_tokens, dtype=cfg.dtype, device=device,
                              seed=cfg.seed, slack=float(cfg.blocks_per_head) / (cfg.max_cache_tokens // B),
                              max_blocks=cfg.blocks_per_head, fill_cache=True)
        else:
            st = W.make_state(S, L, H, B, D, cfg.max_cache_tokens, dtype=cfg.dtype, device=device,
                              seed=cfg.seed, slack=1.0, max_blocks=cfg.blocks_per_head, fill_cache=True,
                              blocks_per_head_alloc=cfg.blocks_per_head)
        self.st = st
        # token positions of the surviving KVs: a random subset of the prompt (what a CR=64 eviction
        # leaves), newest `protected_window` tokens always present
        n = cfg.max_cache_tokens
        heads = L * S * H
        g = torch.Generator(device=device).manual_seed(cfg.seed + 7)
        bt = st.block_tables.view(heads, -1)
        kv_pos = torch.full((st.NB * B,), -1, dtype=torch.int32, device=device)
        recent = torch.arange(cfg.prompt_len - cfg.protected_window, cfg.prompt_len, device=device, dtype=torch.int32)
        ar = torch.arange(B, device=device)
        for h0 in range(0, heads, 8192):    # chunks bound the [heads, prompt_len] scratch (1.5 GB at 64 K heads)
            h1 = min(heads, h0 + 8192)
            keep = torch.rand((h1 - h0, cfg.prompt_len - cfg.protected_window), generator=g, device=device).topk(
                n - cfg.protected_window, dim=1).indices.sort(dim=1).values.to(torch.int32)
            pos = torch.cat([keep, recent.expand(h1 - h0, -1)], dim=1)                      # [chunk, n]
            slots = (bt[h0:h1, : n // B].long()[:, :, None] * B + ar[None, None]).view(h1 - h0, n)
            kv_pos[slots.flatten()] = pos.flatten()
        st.kv_position = kv_pos.view(st.NB, B)
        st.kv_metrics.uniform_(0.01, 1.0, generator=torch.Generator(device=device).manual_seed(cfg.seed + 9))
        # [2, S]: row 0 = tokens so far (= position of the token being appended), row 1 = position of the last
        # cached token; one add per step advances both (no other element-wise torch op runs inside a step)
        self._pos2 = torch.empty((2, S), dtype=torch.int32, device=device)
        self._pos2[0].fill_(cfg.prompt_len)
        self._pos2[1].fill_(cfg.prompt_len - 1)
        self.seq_len, self.last_pos = self._pos2[0], self._pos2[1]
        self.head_bias = torch.zeros((L, H), dtype=torch.float32, device=device)
        self.zeros_s = torch.zeros(S, dtype=torch.int32, device=device)
        self.protected = torch.full((S,), cfg.protected_window, dtype=torch.int32, device=device)
        self.scale = 1.0 / math.sqrt(D)
        self.max_ctx = cfg.blocks_per_head * B
        # workspaces sized once (the reference allocates them in CompressionScheduler.__init__, :75-86)
        N = heads * cfg.blocks_per_head * B
        self.max_kv = N
        self.evicted_logical = torch.empty(N, dtype=torch.int32, device=device)
        self.evicted_heads = torch.empty(N, dtype=torch.int32, device=device)
        self.cache_moves = torch.empty((N, 2), dtype=torch.int32, device=device)
        self.moves_count = torch.empty((S, L, H), dtype=torch.int32, device=device)
        self.evicted_kv_count = torch.empty((S, L, H), dtype=torch.int32, device=device)
        self.evicted_block_count = torch.empty((S, L, H), dtype=torch.int32, device=device)
        self.slot_mapping = torch.empty((L, S, H), dtype=torch.int64, device=device)
        self.hanging = torch.empty((S, L, H), dtype=torch.int32, device=device)
        self.offs = torch.empty((S, L, H), dtype=torch.int32, device=device)
        self.budgets = torch.empty((S,), dtype=torch.int32, device=device)
        self.arangeB = torch.arange(B, dtype=torch.int32, device=device)
        if not cfg.fused_metrics:
            self.temp_metrics = torch.zeros((st.NB, B, cfg.queries_per_kv), dtype=torch.float32, device=device)
        self.deferred = bool(cfg.deferred_append and cfg.fused_metrics and cfg.fused_cache_write)
        if self.deferred:
            self.pend_kv = torch.zeros((L, S, H, 2 * D), dtype=cfg.dtype, device=device)
            self.pend_slot = torch.full((L, S, H), -1, dtype=torch.int64, device=device)
            self.pend_dst = (torch.full((L, S, H), -1, dtype=torch.int64, device=device)
                             if cfg.deferred_move else None)
        self.table_status = torch.zeros(1, dtype=torch.int32, device=device)
        self.alloc = None
        if cfg.allocator == "device":
            self.alloc = capi.DeviceBlockAllocator(st.NB, st.block_tables, st.context_lens, st.kv_position, B)
            used = st.block_tables.view(heads, -1)[:, : n // B].reshape(-1).long()
            self.alloc.free_mask[used] = False
            self.alloc.free_count.fill_(st.NB - used.numel())

    # ------------------------------------------------------------------ compression
    def _budgets(self) -> torch.Tensor:
        """_schedule_seq_evictions with max_cache_tokens >= 0 (scheduler.py:131-141): blocks above
        max_cache_tokens * L * H / B."""
        cfg = self.cfg
        nblk = (self.st.context_lens + cfg.block_size - 1) // cfg.block_size   # [L,S,H]
        total = nblk.sum(dim=(0, 2))
        return (total - self.max_cache_blocks).clamp_(min=0).to(torch.int32)

    @property
    def max_cache_blocks(self) -> int:
        """scheduler.py:117-137: max_cache_tokens is rounded up to a block multiple per head first."""
        cfg = self.cfg
        B = cfg.block_size
        per_head = (cfg.max_cache_tokens + B - 1) // B * B
        return (per_head * cfg.num_layers * cfg.num_kv_heads + B - 1) // B

    def compress(self) -> None:
        """Compression cycle + append, device-side end to end (no host round trip)."""
        cfg, st = self.cfg, self.st
        B = cfg.block_size
        ops = self.ops
        ctx = st.context_lens                                                  # [L,S,H]
        last_pos = self.last_pos
        if cfg.selection == "fused":
            max_blocks = self.max_cache_blocks
            ops.prepare_compression(ctx, self.hanging, self.offs, self.budgets, B, max_blocks)
            ops.select_evictions(self.evicted_logical, self.evicted_kv_count, self.evicted_block_count,
                                 st.kv_metrics, st.kv_position, st.block_tables, ctx, self.hanging, self.offs,
                                 self.budgets, last_pos, self.protected, B, MAX_INT)
            # prepare_compression = 2 launches; select_evictions = 1 (one CTA per sequence) for decode-sized
            # batches of compressed contexts, else rank / threshold / emit
            fused_sel = (cfg.num_seqs >= 96 and B == 16 and st.block_tables.shape[-1] * B <= 512
                         and capi.get_select_mode() != 1) or capi.get_select_mode() == 2
            self.launches += 2 + (1 if fused_sel else ops.SELECT_EVICTIONS_LAUNCHES)
            logical, hanging, offs = self.evicted_logical, self.hanging, self.offs
        else:
            hanging = W.hanging_token_counts(ctx, B).transpose(0, 1).contiguous()  # [S,L,H]
            offs = W.evicted_kv_offsets(ctx, B)
            budgets = self._budgets()
            sm = W.sort_seq_metrics(st)                                        # torch.sort x2 (library)
            ops.schedule_cache_evictions(self.evicted_heads, self.evicted_logical, self.evicted_kv_count, offs,
                                         sm.sorted_indices, sm.seq_block_offsets, sm.layer_by_block,
                                         sm.head_by_block, sm.logical_block_num_by_block, budgets, ctx, hanging,
                                         sm.token_positions, last_pos, self.protected, B, self.max_kv, MAX_INT, True)
            self.launches += 6
            # host re-sort (scheduler.py:472-477)
            order = torch.sort(self.evicted_logical, stable=True).indices
            order = order[torch.sort(self.evicted_heads[order], stable=True).indices]
            logical = self.evicted_logical[order].contiguous()
            self.evicted_block_count.copy_(torch.where(self.evicted_kv_count >= hanging,
                                                       (self.evicted_kv_count - hanging) // B + 1,
                                                       torch.zeros_like(hanging)))
            self.hanging.copy_(hanging)
        # the op itself, not the reference wrapper: that one zero-fills the whole [max_kv, 2] workspace
        # first (vllm/_custom_ops.py:1168, 100 MB per call here); only [0, count) per head is ever read
        ops.schedule_cache_moves_nofill(self.cache_moves, self.moves_count, logical, self.evicted_kv_count, offs,
                                        st.block_tables, ctx, B)
        if self.deferred:
            ops.execute_cache_moves_deferred(st.k_cache, st.v_cache, st.kv_metrics, st.kv_position,
                                             self.cache_moves, self.moves_count, offs, self.pend_kv, self.pend_slot,
                                             self.pend_dst)
        else:
            ops.execute_cache_moves(st.k_cache, st.v_cache, st.kv_metrics, st.kv_position, self.cache_moves,
                                    self.moves_count, offs, 1, 16)
        # schedule_t1_cache_moves + execute_cache_moves (run kernel; + the one-move kernel from 4096 heads up)
        self.launches += 2 + (1 if cfg.num_seqs * cfg.num_layers * cfg.num_kv_heads >= 4096 else 0)

    def flush_pending(self) -> None:
        """Store every parked token to its slot (deferred append): afterwards the cache holds exactly what
        the reference's reshape_and_cache + execute_cache_moves sequence leaves."""
        if not self.deferred:
            return
        self.ops.flush_deferred(self.st.k_cache, self.st.v_cache, self.pend_kv, self.pend_slot, self.pend_dst)

    # ------------------------------------------------------------------ append + forward
    def append_slots(self) -> None:
        """free_compressed_blocks (block.py:184-210) + batch_append_slots (block_manager.py:269-294):
        one launch; then the sequence length advances."""
        cfg, st = self.cfg, self.st
        if self.alloc is not None:
            # free_compressed_blocks (block_manager.py:466-530) then _append_to_sequence_batch (:269-294) against
            # the device allocator: 1 + 3 launches, block tables rewritten in place
            self.alloc.free_compressed(None, self.evicted_block_count)
            self.alloc.append_slots(None, self.seq_len, self.slot_mapping)
            self.launches += 4
        else:
            self.ops.free_and_append_slots(st.context_lens, st.block_tables, self.evicted_block_count, self.hanging,
                                           self.seq_len, self.slot_mapping, st.kv_position, cfg.block_size, True,
                                           self.table_status)
            self.launches += 1
        self._pos2.add_(1)

    def check_status(self) -> None:
        """Raise if the allocator ran out of blocks or a head outgrew its block table (host sync; the reference
        raises at the failing call, block_manager.py:108-110 / an index error in block.py:303-321)."""
        st = int(self.table_status.item())
        if self.alloc is not None:
            st = max(st, int(self.alloc.status.item()))
        if st == 1:
            raise RuntimeError("KV cache out of free blocks")
        if st == 2:
            raise RuntimeError("a head's block table is full (blocks_per_head too small for this workload)")

    def forward_layer(self, l: int, x: torch.Tensor, o: torch.Tensor, last_pos: torch.Tensor,
                      timing=None) -> None:
        """x [S, (Hq + 2H) * D], o [S, Hq * D]: cache write + decode attention of layer l."""
        cfg, st, ops = self.cfg, self.st, self.ops
        H, Q, D, B = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.block_size
        Hq = H * Q
        if self.ops is _default_ops:
            capi.set_attention_pdl(self.pdl_first if l == 0 else self.pdl_rest)
        q = x[:, : Hq * D].view(-1, Hq, D)
        k = x[:, Hq * D: (Hq + H) * D].view(-1, H, D)
        v = x[:, (Hq + H) * D:].view(-1, H, D)
        o = o.view(-1, Hq, D)
        if self.deferred:
            ops.paged_attention_kvc_fused_step_deferred(o, st.kv_metrics, q, k, v, st.k_cache, st.v_cache,
                                                        self.slot_mapping[l].reshape(-1), self.head_bias[l],
                                                        self.pend_kv[l], self.pend_slot[l],
                                                        None if self.pend_dst is None else self.pend_dst[l], H, self.scale,
                                                        st.block_tables[l], st.context_lens[l], st.kv_position,
                                                        last_pos, self.zeros_s, B, self.max_ctx)
            self.launches += 1
            return
        if cfg.fused_metrics and cfg.fused_cache_write:
            ops.paged_attention_kvc_fused_step(o, st.kv_metrics, q, k, v, st.k_cache, st.v_cache,
                                               self.slot_mapping[l].reshape(-1), self.head_bias[l], H, self.scale,
                                               st.block_tables[l], st.context_lens[l], st.kv_position, last_pos,
                                               self.zeros_s, B, self.max_ctx)
            self.launches += 1
            return
        ops.reshape_and_cache_kvc(k, v, st.k_cache, st.v_cache, st.kv_metrics,
                                  self.slot_mapping[l].reshape(-1), self.head_bias[l], "auto", 1.0, 1.0)
        if timing is not None:
            timing[0].record()
        if cfg.fused_metrics:
            ops.paged_attention_kvc_fused_metrics(o, st.kv_metrics, q, st.k_cache, st.v_cache, H, self.scale,
                                                  st.block_tables[l], st.context_lens[l], st.kv_position,
                                                  last_pos, self.zeros_s, B, self.max_ctx)
        else:
            ops.paged_attention_kvc_v1(o, self.temp_metrics, q, st.k_cache, st.v_cache, H, self.scale,
                                       st.block_tables[l], st.context_lens[l], st.kv_position, last_pos,
                                       self.zeros_s, B, self.max_ctx, None, "auto", 1.0, 1.0, True)
        if timing is not None:
            timing[1].record()
        self.launches += 2

    def _aggregate_unfused(self) -> None:
        if not self.cfg.fused_metrics:
            # clear_temp_metrics + aggregate_decode (metrics.py:337-342, 429-439), whole-cache passes
            self.st.kv_metrics += (self.temp_metrics ** 2).sum(dim=-1)
            self.temp_metrics.zero_()

    def forward(self, qkv: torch.Tensor, out: torch.Tensor) -> None:
        """qkv [L, S, (Hq + 2H) * D], out [L, S, Hq * D]."""
        last_pos = self.last_pos
        for l in range(self.cfg.num_layers):
            self.forward_layer(l, qkv[l], out[l], last_pos)
        self._aggregate_unfused()

    def step(self, qkv: torch.Tensor, out: torch.Tensor) -> None:
        """One decode step of the hot path with device-resident inputs (eager launches)."""
        self.compress()
        self.append_slots()
        self.forward(qkv, out)

    # ------------------------------------------------------------------ CUDA graph of the step
    def capture(self, qkv: torch.Tensor, out: torch.Tensor) -> bool:
        """Capture step(qkv, out) into a CUDA graph (the step has no host round trip when the
        fused selection is used).  Returns False when the configuration cannot be captured."""
        if self.cfg.selection != "fused":
            return False
        n0 = self.launches
        self.step(qkv, out)                       # warm up allocator + count launches
        self.launches_per_step = self.launches - n0
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.step(qkv, out)
        self.launches = n0 + self.launches_per_step
        return True

    def replay(self) -> None:
        self.graph.replay()
        self.launches += self.launches_per_step

    # ------------------------------------------------------------------ end to end with host buffers
    def step_host(self, qkv_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Public end-to-end call: the step's Q/K/V live in (pinned) host memory and the attention
        output is returned to host memory.  Per-layer H2D copies run on a side stream ahead of the
        compute, D2H copies behind it (in groups of cfg.host_copy_layers layers); PCIe is full duplex so both
        overlap the kernels."""
        cfg = self.cfg
        main = torch.cuda.current_stream()
        if not hasattr(self, "_h2d"):
            self._h2d, self._d2h = torch.cuda.Stream(), torch.cuda.Stream()
            self._qkv_dev = torch.empty(qkv_host.shape, dtype=qkv_host.dtype, device=self.dev)
            self._out_dev = torch.empty(out_host.shape, dtype=out_host.dtype, device=self.dev)
            self._ev_in = [torch.cuda.Event() for _ in range(cfg.num_layers)]
            self._ev_done = [torch.cuda.Event() for _ in range(cfg.num_layers)]
        self._h2d.wait_stream(main)   # previous step's compute has consumed the staging buffers
        # copies are issued per group of `host_copy_layers` layers: every DMA carries a fixed cost, and 2 x 32
        # small copies in flight on the full-duplex link cost ~0.4 ms per step more than 2 x 8 larger ones
        G = max(1, min(cfg.host_copy_layers, cfg.num_layers))
        groups = [(a, min(a + G, cfg.num_layers)) for a in range(0, cfg.num_layers, G)]
        with torch.cuda.stream(self._h2d):
            for gi, (a, b) in enumerate(groups):
                self._qkv_dev[a:b].copy_(qkv_host[a:b], non_blocking=True)
                self._ev_in[gi].record(self._h2d)
        self.compress()
        self.append_slots()
        last_pos = self.last_pos
        for gi, (a, b) in enumerate(groups):
            main.wait_event(self._ev_in[gi])
            for l in range(a, b):
                self.forward_layer(l, self._qkv_dev[l], self._out_dev[l], last_pos)
            self._ev_done[gi].record(main)
            self._d2h.wait_event(self._ev_done[gi])
            with torch.cuda.stream(self._d2h):
                out_host[a:b].copy_(self._out_dev[a:b], non_blocking=True)
        self._aggregate_unfused()
        main.wait_stream(self._d2h)   # the step is complete when its results are in host memory

    def capture_host(self, qkv_host: torch.Tensor, out_host: torch.Tensor) -> bool:
        """Capture step_host (copies on side streams included) into one CUDA graph bound to these
        pinned host buffers: the per-layer H2D / compute / D2H pipeline then costs one graph launch."""
        if self.cfg.selection != "fused":
            return False
        self.step_host(qkv_host, out_host)  # creates streams / staging buffers, warms the allocator
        torch.cuda.synchronize()
        self.host_graph = torch.cuda.CUDAGraph()
        n0 = self.launches
        with torch.cuda.graph(self.host_graph):
            self.step_host(qkv_host, out_host)
        self.launches = n0
        return True

    def replay_host(self) -> None:
        self.host_graph.replay()
        self.launches += getattr(self, "launches_per_step", 0)

    # ------------------------------------------------------------------ end to end, pipelined across steps
    def submit_host(self, qkv_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """Public end-to-end call, pipelined: enqueue one decode step whose Q/K/V live in pinned host memory
        and whose attention output goes back to pinned host memory, and return at once.  Two device staging
        buffers alternate, so the H2D copy of step t+1 and the D2H copy of step t-1 run under the compute of
        step t (PCIe is full duplex; three streams): a step costs max(H2D, compute, D2H) instead of their sum.
        out_host of step t is valid after wait_host() (or after the second following submit_host returns
        its events); the caller must not reuse the host buffers of a step before that."""
        cfg = self.cfg
        main = torch.cuda.current_stream()
        if not hasattr(self, "_pipe"):
            qshape = (cfg.num_layers, cfg.num_seqs, (cfg.num_kv_heads * cfg.queries_per_kv + 2 * cfg.num_kv_heads)
                      * cfg.head_size)
            oshape = (cfg.num_layers, cfg.num_seqs, cfg.num_kv_heads * cfg.queries_per_kv * cfg.head_size)
            assert tuple(qkv_host.shape) == qshape and tuple(out_host.shape) == oshape
            pipe = {"h2d": torch.cuda.Stream(), "d2h": torch.cuda.Stream(), "n": 0, "graphs": [None, None],
                    "qkv": [torch.empty(qshape, dtype=cfg.dtype, device=self.dev) for _ in range(2)],
                    "out": [torch.empty(oshape, dtype=cfg.dtype, device=self.dev) for _ in range(2)],
                    "in_ready": [torch.cuda.Event() for _ in range(2)],
                    "computed": [torch.cuda.Event() for _ in range(2)],
                    "drained": [torch.cuda.Event() for _ in range(2)]}
            for ev in pipe["computed"] + pipe["drained"]:
                ev.record(main)
            self._pipe = pipe
        pipe = self._pipe
        i = pipe["n"] & 1
        pipe["n"] += 1
        h2d, d2h = pipe["h2d"], pipe["d2h"]
        h2d.wait_event(pipe["computed"][i])          # the step two back has consumed this staging buffer
        with torch.cuda.stream(h2d):
            pipe["qkv"][i].copy_(qkv_host, non_blocking=True)
            pipe["in_ready"][i].record(h2d)
        main.wait_event(pipe["in_ready"][i])
        main.wait_event(pipe["drained"][i])          # ... and its output has left the device
        if cfg.selection == "fused":
            if pipe["graphs"][i] is None:            # first use of this buffer: run eagerly once, then capture
                n0 = self.launches
                self.step(pipe["qkv"][i], pipe["out"][i])
                self.launches_per_step = self.launches - n0
                graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(graph):        # capture does not execute: the state stays where it is
                    self.step(pipe["qkv"][i], pipe["out"][i])
                self.launches = n0 + self.launches_per_step
                pipe["graphs"][i] = graph
            else:
                pipe["graphs"][i].replay()
                self.launches += getattr(self, "launches_per_step", 0)
        else:
            self.step(pipe["qkv"][i], pipe["out"][i])
        pipe["computed"][i].record(main)
        d2h.wait_event(pipe["computed"][i])
        with torch.cuda.stream(d2h):
            out_host.copy_(pipe["out"][i], non_blocking=True)
            pipe["drained"][i].record(d2h)

    def wait_host(self) -> None:
        """Every submitted step is complete and its output is in host memory."""
        if hasattr(self, "_pipe"):
            torch.cuda.current_stream().wait_stream(self._pipe["d2h"])
        torch.cuda.current_stream().synchronize()

    # ------------------------------------------------------------------ measurement helper
    def profile_attention(self, qkv: torch.Tensor, out: torch.Tensor, steps: int = 5):
        """Device time of the attention launches alone, doing exactly the work they do inside a step: the
        compression and append of one more step run first (so the parked tokens and deferred moves of that step
        are outstanding), then the L launches are captured in a CUDA graph -- no host launch gap sits between the
        two CUDA events that bracket each replay (events recorded on the launching stream).  The state the
        launches consume (pend_dst / pend_kv) is re-armed before every replay, outside the events.  Returns
        (mean ms per launch, mean algorithmic bytes per launch)."""
        cfg, st = self.cfg, self.st
        L = cfg.num_layers
        H, Q, D, B = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.block_size
        Hq = H * Q
        self.compress()
        self.append_slots()
        last_pos = self.last_pos
        nbytes = self.attention_bytes() / L

        def attn_only():
            for l in range(L):
                self.forward_layer(l, qkv[l], out[l], last_pos)
            self._aggregate_unfused()

        armed = None
        if self.deferred:
            armed = [(t, t.clone()) for t in (self.pend_kv, self.pend_slot) +
                     ((self.pend_dst,) if self.pend_dst is not None else ())]

        def rearm():
            if armed is not None:
                for t, v in armed:
                    t.copy_(v)

        n0 = self.launches
        attn_only()
        torch.cuda.synchronize()
        rearm()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            attn_only()
        rearm()
        graph.replay()
        torch.cuda.synchronize()
        pairs = []
        for _ in range(steps):
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            rearm()
            ev[0].record()
            graph.replay()
            ev[1].record()
            pairs.append(ev)
        torch.cuda.synchronize()
        self.launches = n0
        ms = sum(a.elapsed_time(b) for a, b in pairs) / (len(pairs) * L)
        return ms, nbytes

    @property
    def tokens_per_step(self) -> int:
        return self.cfg.num_seqs

    def attention_bytes(self) -> int:
        """Algorithmic bytes of one step's attention launches: K+V of every
        context token once (2*D*2 B) + position read (4 B) + metric read-modify-write (8 B)."""
        ctx_total = int(self.st.context_lens.sum())
        n = ctx_total * (2 * self.cfg.head_size * 2 + 4 + 8)
        if self.cfg.fused_metrics and self.cfg.fused_cache_write:
            # the same launches also store the step's new token of every head (K6: K + V + metric) ...
            n += self.st.context_lens.numel() * (2 * self.cfg.head_size * 2 + 4)
            # ... and apply the deferred moves recorded by the compression (K5: 1048 B per move)
            if getattr(self, "pend_dst", None) is not None:
                n += int((self.pend_dst >= 0).sum()) * 1048
        return n
