#!/usr/bin/env python
"""Decode attention over LONG per-head contexts (the first decode steps after a 6000-token prefill, before the
cache is compressed; CR <= 4): paged_attention_kvc_v2 = 512-token partitions + reduce.  The partition kernel runs
on tcgen05 / TMEM (paged_attention_tc_kernel, --v2-mode 2) or on mma.sync (--v2-mode 1).  Prints one JSON line: ms per call, algorithmic GB/s (K + V once per (sequence,
kv head): ctx * 512 B) and the fraction of the measured HBM peak.
    python tools/long_context_bench.py [--seqs 8] [--ctx 6000] [--q 4] [--bf16]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=8)
    ap.add_argument("--ctx", type=int, default=6000)
    ap.add_argument("--q", type=int, default=4)
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--version", default="v2", choices=["v1", "v2"])
    ap.add_argument("--v2-mode", type=int, default=0, help="kvc_set_attention_v2_mode: 1 mma.sync, 2 tcgen05")
    a = ap.parse_args()
    from vllm_kvcompress_b200 import _custom_ops as ops
    from vllm_kvcompress_b200 import capi
    from vllm_kvcompress_b200 import workload as W
    capi.set_attention_v2_mode(a.v2_mode)
    dev = torch.device("cuda", 0)
    S, H, Q, D, B = a.seqs, 8, a.q, 128, 16
    dt = torch.bfloat16 if a.bf16 else torch.float16
    st = W.make_state(S, 1, H, B, D, a.ctx, dtype=dt, device=dev, seed=1, slack=1.0)
    q = (torch.rand((S, H * Q, D), device=dev) - 0.5).to(dt)
    out = torch.empty_like(q)
    metric = torch.zeros((st.NB, B, Q), dtype=torch.float32, device=dev)
    last = torch.full((S,), a.ctx - 1, dtype=torch.int32, device=dev)
    zeros = torch.zeros(S, dtype=torch.int32, device=dev)
    P = (a.ctx + 511) // 512
    es = torch.empty((S, H * Q, P), dtype=torch.float32, device=dev)
    ml = torch.empty_like(es)
    tmp = torch.empty((S, H * Q, P, D), dtype=dt, device=dev)
    tmpm = torch.zeros((1,), dtype=torch.float32, device=dev)

    def call():
        if a.version == "v2":
            ops.paged_attention_kvc_v2(out, metric, es, ml, tmp, tmpm, q, st.k_cache, st.v_cache, H, D ** -0.5,
                                       st.block_tables[0], st.context_lens[0], st.kv_position, last, zeros, B, a.ctx,
                                       None, "auto", 1.0, 1.0, True)
        else:
            ops.paged_attention_kvc_v1(out, metric, q, st.k_cache, st.v_cache, H, D ** -0.5, st.block_tables[0],
                                       st.context_lens[0], st.kv_position, last, zeros, B, a.ctx, None, "auto", 1.0,
                                       1.0, True)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: every timed call starts cold
    for _ in range(3):
        call()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = S * H * a.ctx * (2 * D * 2 + 4 + 4 * Q)
    peak = 6542.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    print(json.dumps({"version": a.version, "v2_mode": a.v2_mode, "seqs": S,
                      "ctx": a.ctx, "q": Q, "dtype": str(dt), "ms": ms, "gbs": nbytes / ms / 1e6,
                      "frac": nbytes / ms / 1e6 / peak, "bytes": nbytes}))


if __name__ == "__main__":
    main()
