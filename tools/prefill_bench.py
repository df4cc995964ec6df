#!/usr/bin/env python
"""prefill_kv_metrics timing over a few prompt mixes (CUDA events, median of 10 after 3 warm-up calls, operands from
DRAM: an L2-sized buffer is rewritten between calls).  python tools/prefill_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def run(device, lens, n_observed=32, Hq=32, H=8, D=128, dtype=torch.float16, reps=10):
    from vllm_kvcompress_b200 import _custom_ops as ops
    T = sum(lens)
    g = torch.Generator(device=device).manual_seed(3)
    q = torch.randn((T, Hq, D), generator=g, device=device, dtype=torch.float32).to(dtype)
    k = torch.randn((T, H, D), generator=g, device=device, dtype=torch.float32).to(dtype)
    buf = torch.zeros(len(lens), dtype=torch.int32, device=device)
    starts = torch.tensor([0] + list(torch.tensor(lens).cumsum(0).tolist()), dtype=torch.int32, device=device)
    # slots as the engine lays them out: every head owns randomly placed blocks of 16 consecutive tokens
    nblk = (T + 15) // 16
    blocks = torch.randperm(H * nblk, device=device).view(H, nblk)
    tok = torch.arange(T, device=device)
    slots = (blocks[:, tok // 16] * 16 + tok % 16).t().contiguous()
    metrics = torch.zeros((H * nblk, 16), dtype=torch.float32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ts = []
    for i in range(reps + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.prefill_kv_metrics(q, k, lens, D ** -0.5, buf, n_observed=n_observed, kv_metrics=metrics,
                               slot_mapping=slots, return_metrics=False, seq_start_loc=starts)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    # the same call replayed from a CUDA graph: device time without the host's launch path
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.prefill_kv_metrics(q, k, lens, D ** -0.5, buf, n_observed=n_observed, kv_metrics=metrics,
                               slot_mapping=slots, return_metrics=False, seq_start_loc=starts)
    tg = []
    for i in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            tg.append(a.elapsed_time(b))
    ms_graph = sorted(tg)[len(tg) // 2]
    flops = 2 * sum(2.0 * min(n_observed, n) * Hq * n * D for n in lens)
    kbytes = T * H * D * k.element_size()
    return {"prompts": f"{len(lens)} x {lens[0]}" if len(set(lens)) == 1 else lens, "n_observed": n_observed,
            "ms": round(ms, 4), "ms_min": round(min(ts), 4), "ms_graph": round(ms_graph, 4),
            "tflops": round(flops / (ms * 1e-3) / 1e12, 1), "tflops_graph": round(flops / (ms_graph * 1e-3) / 1e12, 1),
            "k_gbs": round(kbytes / (ms * 1e-3) / 1e9, 1)}


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    for lens, nobs in (([6000] * 4, 32), ([6000], 32), ([2000] * 16, 32), ([19000], 32), ([6000] * 4, 128),
                       ([6000] * 16, 32)):
        print(json.dumps(run(dev, lens, nobs)))
