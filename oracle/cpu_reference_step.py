"""TEST / BASELINE INFRASTRUCTURE ONLY.  The part of the decode hot path that the REFERENCE itself implements on
the CPU, timed with the reference's own kernels (vLLM CPU backend: csrc/cpu/attention.cpp paged_attention_v1 and
csrc/cpu/cache.cpp reshape_and_cache, compiled unmodified into oracle/_ref/_vllm_cpu_ref.so by Makefile.ref).

The fork has no CPU build of any KV-Compress op (eviction scheduling, moves, per-head block tables, metric
recording: csrc/cpu/torch_bindings.cpp registers none of them), so this is NOT the full step that
oracle/cpu_decode_step.py (the port, `cpu_baseline`) runs: it is the per-layer cache write + paged-attention decode
of the vanilla layout (one block table per sequence, every head the same length) at the compressed context length,
i.e. an upper bound on what the reference's CPU path could deliver for this workload if its compression were free.
bf16 (the CPU backend dispatches float / bfloat16 only)."""
from __future__ import annotations

import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "_vllm_cpu_ref.so")


def available() -> bool:
    return os.path.exists(SO)


def measure(num_seqs=16, num_layers=32, num_kv_heads=8, queries_per_kv=4, head_size=128, block_size=16,
            context_len=129, seconds=8.0, threads=None):
    torch.ops.load_library(SO)
    r = torch.ops._vllm_cpu_ref
    if threads:
        torch.set_num_threads(threads)
    S, L, H, Q, D, B = num_seqs, num_layers, num_kv_heads, queries_per_kv, head_size, block_size
    Hq = H * Q
    x = 16 // 2
    nblk = (context_len + B - 1) // B
    NB = S * nblk * L + 4                       # one unified cache for all layers, like the engine's
    g = torch.Generator().manual_seed(0)
    kc = (torch.rand((NB, H, D // x, B, x), generator=g) - 0.5).to(torch.bfloat16)
    vc = (torch.rand((NB, H, D, B), generator=g) - 0.5).to(torch.bfloat16)
    perm = torch.randperm(NB - 4, generator=g).to(torch.int32).view(L, S, nblk)
    seq_lens = torch.full((S,), context_len, dtype=torch.int32)
    q = (torch.rand((L, S, Hq, D), generator=g) - 0.5).to(torch.bfloat16)
    k = (torch.rand((L, S, H, D), generator=g) - 0.5).to(torch.bfloat16)
    v = (torch.rand((L, S, H, D), generator=g) - 0.5).to(torch.bfloat16)
    out = torch.empty((S, Hq, D), dtype=torch.bfloat16)
    slots = (perm[:, :, -1].long() * B + (context_len - 1) % B)          # [L, S] slot of the step's new token

    def step():
        for l in range(L):
            r.reshape_and_cache(k[l], v[l], kc, vc, slots[l], "auto", 1.0, 1.0)
            r.paged_attention_v1(out, q[l], kc, vc, H, D ** -0.5, perm[l], seq_lens, B, context_len, None, "auto",
                                 1.0, 1.0, 0, 0, 1, 64, 0)

    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or time.perf_counter() - t0 < seconds:
        step()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": S / dt, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "reference",
            "what": "vLLM CPU backend kernels of the reference (csrc/cpu: reshape_and_cache + paged_attention_v1, "
                    "vanilla layout, bf16) on the per-layer attention part of the step only -- the fork has no CPU "
                    "implementation of the KV-Compress ops",
            "sample": f"{S} sequences x {L} layers x {Hq} query heads, context {context_len} tokens, "
                      f"{reps} steps in {dt * reps:.1f} s", "ms_per_step": dt * 1e3}
