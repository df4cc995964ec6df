#!/usr/bin/env python
"""kvcompress_reshape_and_cache on a prefill batch (every head's tokens fill randomly placed blocks of 16 in order)
and on a decode batch (one token per head).  CUDA events, median of 10, L2 rewritten before each call.
python tools/k6_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def run(device, lens, H=8, D=128, B=16, decode=False, reps=10):
    from vllm_kvcompress_b200 import _custom_ops as ops
    T = sum(lens)
    g = torch.Generator(device=device).manual_seed(5)
    k = torch.randn((T, H, D), generator=g, device=device, dtype=torch.float32).to(torch.float16)
    v = torch.randn((T, H, D), generator=g, device=device, dtype=torch.float32).to(torch.float16)
    # per prompt and head: blocks in order of the prompt's tokens, block ids random
    nblk_p = [(n + B - 1) // B for n in lens]
    NB = H * sum(nblk_p) if not decode else T * H
    perm = torch.randperm(NB, device=device)
    slots = torch.empty((T, H), dtype=torch.int64, device=device)
    if decode:   # T "sequences", one new token each, somewhere inside a block
        slots[:] = (perm.view(T, H) * B + torch.randint(0, B, (T, H), device=device))
    else:
        t0, b0 = 0, 0
        for n, nb in zip(lens, nblk_p):
            tok = torch.arange(n, device=device)
            blocks = perm[b0:b0 + H * nb].view(H, nb)
            slots[t0:t0 + n] = (blocks[:, tok // B] * B + tok % B).t()
            t0 += n
            b0 += H * nb
    kc = torch.zeros((NB, D // 8, B, 8), dtype=torch.float16, device=device)
    vc = torch.zeros((NB, D, B), dtype=torch.float16, device=device)
    metrics = torch.zeros((NB, B), dtype=torch.float32, device=device)
    bias = torch.zeros((H,), dtype=torch.float32, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ts = []
    for i in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.reshape_and_cache_kvc(k, v, kc, vc, metrics, slots.view(-1), bias, "auto", 1.0, 1.0)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    # check the result once against plain indexing
    blk, off = slots // B, slots % B
    ok = bool(torch.equal(vc[blk, :, off], v) and torch.equal(kc[blk, :, off, :].reshape(T, H, D), k))
    nbytes = 2 * T * H * D * 2 * 2   # K and V read + written
    return {"tokens": T, "decode": decode, "ms": round(ms, 4), "gbs": round(nbytes / (ms * 1e-3) / 1e9, 1), "ok": ok}


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    for lens, dec in (([6000] * 4, False), ([6000], False), ([1999, 37, 4100], False), ([1] * 256, True)):
        print(json.dumps(run(dev, lens, decode=dec)))
