#!/usr/bin/env python
"""Host <-> device copy ceiling of the box for the e2e step's traffic (100.7 MB H2D + 67.1 MB D2H per GPU per
step, pinned host memory, H2D and D2H concurrently on two streams), with N processes (one per GPU) copying at
the same time:   python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py      (or plain python: N=1)
Rank 0 prints one JSON line: per-GPU and aggregate GB/s for H2D alone, D2H alone and both together, and the
step time the copies alone would take -- the ceiling DecodeHotPath.submit_host can reach on this box."""
import json
import os
import time

import torch


def main():
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    A, B = 100663296, 67108864
    h_in = torch.empty(A, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(A, dtype=torch.uint8, device=dev)
    h_out = torch.empty(B, dtype=torch.uint8, pin_memory=True)
    d_out = torch.empty(B, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)

    res = {}
    for name, (h2d, d2h) in (("h2d", (1, 0)), ("d2h", (0, 1)), ("both", (1, 1))):
        for _ in range(3):
            run(h2d, d2h)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            run(h2d, d2h)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())
    if rank == 0:
        out = {"n_gpus": world, "h2d_bytes": A, "d2h_bytes": B,
               "h2d_gbs_per_gpu": A / res["h2d"] / 1e9, "d2h_gbs_per_gpu": B / res["d2h"] / 1e9,
               "both_ms": res["both"] * 1e3, "both_gbs_per_gpu": (A + B) / res["both"] / 1e9,
               "both_gbs_aggregate": world * (A + B) / res["both"] / 1e9,
               "e2e_tokens_per_s_ceiling_batch256": world * 256 / res["both"]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
