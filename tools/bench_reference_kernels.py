#!/usr/bin/env python
"""The decode hot-path step driven through the REFERENCE's own CUDA kernels (oracle/_ref/_kvc_ref.so =
the fork's three .cu files compiled unmodified for sm_100a), in the order the fork's engine runs them
(oracle/ref_ops.py).  bench.py runs the same measurement as its `reference_cuda` leg; this is the stand-alone
form.  Prints one JSON line.  Run on a B200: python tools/bench_reference_kernels.py [--batch 256] [--steps 5]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def measure(batch=256, layers=32, steps=5, warmup=2, preset=None, device=None):
    from oracle.ref_ops import RefOps
    from vllm_kvcompress_b200 import engine as E
    dev = device or torch.device("cuda", 0)
    cfg = E.HotPathConfig(num_seqs=batch, num_layers=layers, selection="chain", fused_metrics=False,
                          fused_cache_write=False, allocator="static", deferred_append=False, **(preset or {}))
    eng = E.DecodeHotPath(cfg, dev, ops_module=RefOps())
    H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
    qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), device=dev) - 0.5).to(cfg.dtype)
    out = torch.empty((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    for _ in range(warmup):
        eng.step(qkv, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.step(qkv, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del eng
    torch.cuda.empty_cache()
    return {"impl": "reference CUDA kernels (fork's .cu files, sm_100a rebuild) + torch.sort chain, eager launches",
            "value": S / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms, "batch": S, "steps": steps,
            "warmup": warmup}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    print(json.dumps(measure(a.batch, 32, a.steps, a.warmup)))
