#!/usr/bin/env python
"""Device allocator at prefill scale: add 4 prompts of 6000 tokens (32 layers x 8 heads: 1024 heads x 375 blocks),
free the blocks a CR 64 compression releases, remove the sequences.  CUDA events, one call each after a warm-up
cycle.  python tools/alloc_bench.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def measure():
    from vllm_kvcompress_b200 import capi
    dev = "cuda"
    L, H, B, MS, n_seq, P = 32, 8, 16, 8, 4, 6000
    MB = (19000 + B - 1) // B
    NB = 4 * L * H * ((P + B - 1) // B) + 100000
    bt = torch.full((L, MS, H, MB), -1, dtype=torch.int32, device=dev)
    ctx = torch.zeros((L, MS, H), dtype=torch.int32, device=dev)
    pos = torch.full((NB, B), -1, dtype=torch.int32, device=dev)
    meta = tuple(torch.full((NB,), -1 if i == 0 else 0, dtype=torch.int32, device=dev) for i in range(4))
    al = capi.DeviceBlockAllocator(NB, bt, ctx, pos, B, metadata=meta)
    slots = torch.arange(n_seq, dtype=torch.int32, device=dev)
    lens = torch.full((n_seq,), P, dtype=torch.int32, device=dev)
    keep_blocks = 128 // B
    counts = torch.full((n_seq, L, H), (P + B - 1) // B - keep_blocks, dtype=torch.int32, device=dev)
    res = {}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = round(a.elapsed_time(b), 4)

    for rep in range(2):   # the second cycle is reported
        timed("add_sequences_ms", lambda: al.add_sequences(slots, lens, P))
        timed("free_compressed_ms", lambda: al.free_compressed(slots, counts))
        timed("remove_sequences_ms", lambda: al.remove_sequences(slots))
    res["blocks_allocated"] = n_seq * L * H * ((P + B - 1) // B)
    res["num_blocks"] = NB
    return res


if __name__ == "__main__":
    print(json.dumps(measure()))
