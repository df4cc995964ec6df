#!/usr/bin/env python
"""execute_cache_moves on the post-prefill shape of BASELINE.md config 5 (S=3 sequences of 6000 tokens, Llama-3-8B
heads: L=32, H=8; max_kv_per_compression = 5e6 slots caps a post-prefill call at 3 sequences), CR in {64, 16, 8, 4}
=> max_cache_tokens in {128, 368, 736, 1488}:

  achieved_gbs / frac   algorithmic bytes (1048 B per move) / time vs the measured HBM peak
  floor_bytes           sector floor of THIS move list (tools/moves_floor.py): distinct 32 B sectors read + written
                        + fill reads of partly overwritten sectors; what the cache layout costs in DRAM traffic
  floor_gbs / frac_floor_bw   floor bytes / time vs the HBM peak: how close the kernel runs to the layout's floor
  traffic / frac_of_floor      measured DRAM bytes per launch (committed ncu capture, profiles/traffic.json) and
                               floor_bytes / traffic

Protocol: 5 warm-up + 20 timed repetitions, each on a re-initialised cache (K / V /
metrics / positions restored from a pristine copy outside the timed region), CUDA events around the launch.

Used by bench.py (rank 0) and stand-alone:  python tools/compaction_bench.py [--cr 64 4] [--mode 0|1] [--once]
(--once: one launch per CR and nothing else timed -- the ncu target).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

KEEP = {64: 128, 16: 368, 8: 736, 4: 1488}


def _traffic(cr):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get("execute_cache_moves", {}).get(f"cr{cr}")
    return None if d is None else d["dram_bytes_read_per_launch"] + d["dram_bytes_write_per_launch"]


def build_schedule(device, cr, S=3, L=32, H=8, B=16, D=128, ctx=6000):
    """state + eviction selection + move schedule for one post-prefill call"""
    from vllm_kvcompress_b200 import _custom_ops as ops
    from vllm_kvcompress_b200 import workload as W
    keep = KEEP[cr]
    st = W.make_state(S, L, H, B, D, ctx, dtype=torch.float16, device=device, seed=cr, slack=1.0, fill_cache=True)
    hanging = W.hanging_token_counts(st.context_lens, B).transpose(0, 1).contiguous()
    offs = W.evicted_kv_offsets(st.context_lens, B)
    N = int(((st.context_lens + B - 1) // B).sum()) * B
    nblk_total = int(((st.context_lens + B - 1) // B).sum(dim=(0, 2))[0])
    budgets = torch.full((S,), nblk_total - keep * L * H // B, dtype=torch.int32, device=device)
    last = torch.full((S,), ctx - 1, dtype=torch.int32, device=device)
    prot = torch.full((S,), 32, dtype=torch.int32, device=device)
    logical = torch.empty(N, dtype=torch.int32, device=device)
    cnt = torch.empty((S, L, H), dtype=torch.int32, device=device)
    blk = torch.empty_like(cnt)
    sel = lambda: ops.select_evictions(logical, cnt, blk, st.kv_metrics, st.kv_position, st.block_tables,
                                       st.context_lens, hanging, offs, budgets, last, prot, B, W.MAX_INT)
    moves = torch.empty((N, 2), dtype=torch.int32, device=device)
    mcnt = torch.empty((S, L, H), dtype=torch.int32, device=device)
    k4 = lambda: ops.schedule_cache_moves_nofill(moves, mcnt, logical, cnt, offs, st.block_tables, st.context_lens, B)
    return dict(st=st, offs=offs, N=N, sel=sel, k4=k4, moves=moves, mcnt=mcnt, ops=ops)


def _timed(fn, reps, setup=None):
    ts = []
    for _ in range(reps):
        if setup is not None:
            setup()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return ts


def run(device, peak, crs=(64, 16, 8, 4), reps=20, warmup=5, floor=True):
    import moves_floor
    res = {}
    for cr in crs:
        c = build_schedule(device, cr)
        st, ops = c["st"], c["ops"]
        # the schedule builders: operands from DRAM (an L2-sized buffer is rewritten before every call, which also
        # keeps the device busy up to the start event), median of 10
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        sel_ms = sorted(_timed(c["sel"], 12, setup=flush.zero_)[2:])[5]
        k4_ms = sorted(_timed(c["k4"], 12, setup=flush.zero_)[2:])[5]
        del flush
        nmoves = int(c["mcnt"].sum())
        pristine = [t.clone() for t in (st.k_cache, st.v_cache, st.kv_metrics, st.kv_position)]

        def restore():
            for t, v in zip((st.k_cache, st.v_cache, st.kv_metrics, st.kv_position), pristine):
                t.copy_(v)

        fn = lambda: ops.execute_cache_moves(st.k_cache, st.v_cache, st.kv_metrics, st.kv_position, c["moves"],
                                             c["mcnt"], c["offs"], 1, 16)
        ts = _timed(fn, warmup + reps, setup=restore)[warmup:]
        t = sorted(ts)[len(ts) // 2]
        gbs = nmoves * 1048 / (t * 1e-3) / 1e9
        r = {"moves": nmoves, "ms": t, "ms_min": min(ts), "reps": reps, "achieved_gbs": gbs, "frac": gbs / peak,
             "cache_gb": st.NB * 8192 / 1e9,
             # the schedule builders in front of it (12 B per candidate slot, 20 B per move)
             "select_evictions_ms": sel_ms, "select_gbs": c["N"] * 12 / (sel_ms * 1e-3) / 1e9,
             "schedule_moves_ms": k4_ms}
        if floor:
            mv = moves_floor.gather_moves(c["moves"].cpu().numpy(), c["mcnt"].cpu().numpy(), c["offs"].cpu().numpy())
            f = moves_floor.sector_floor(mv)
            f64 = moves_floor.sector_floor(mv, 2)
            r.update({"floor_bytes": f["floor_bytes"], "floor_bytes_per_move": f["per_move"],
                      "floor_bytes_fetch64": f64["floor_bytes"],
                      "floor_gbs": f["floor_bytes"] / (t * 1e-3) / 1e9,
                      "frac_floor_bw": f["floor_bytes"] / (t * 1e-3) / 1e9 / peak})
            tr = _traffic(cr)
            r["traffic"] = tr
            r["frac_of_floor"] = None if not tr else f["floor_bytes"] / tr
            r["frac_of_floor_fetch64"] = None if not tr else f64["floor_bytes"] / tr
        res[f"cr{cr}"] = r
        del c, st, pristine
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cr", type=int, nargs="*", default=[64, 16, 8, 4])
    ap.add_argument("--mode", type=int, default=None, help="kvc_set_moves_mode: 0 register gather (default), 1 TMA-staged")
    ap.add_argument("--once", action="store_true", help="one execute_cache_moves launch per CR (ncu target)")
    ap.add_argument("--l2-fetch", type=int, default=None,
                    help="experiment: cudaLimitMaxL2FetchGranularity in bytes (32 / 64 / 128; device default 64)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.init()
    torch.zeros(1, device=dev)
    if a.l2_fetch is not None:
        from cuda.bindings import runtime as cudart
        err, = cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, a.l2_fetch)
        err2, val = cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity)
        print("l2 fetch granularity set:", err, "now", val, file=sys.stderr)
    from vllm_kvcompress_b200 import capi
    if a.mode is not None:
        capi.set_moves_mode(a.mode)
    if a.once:
        for cr in a.cr:
            c = build_schedule(dev, cr)
            c["sel"](), c["k4"]()
            st = c["st"]
            c["ops"].execute_cache_moves(st.k_cache, st.v_cache, st.kv_metrics, st.kv_position, c["moves"], c["mcnt"],
                                         c["offs"], 1, 16)
            torch.cuda.synchronize()
            print("cr", cr, "moves", int(c["mcnt"].sum()))
        return
    peak = 6542.7
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    print(json.dumps({"moves_mode": capi.get_moves_mode(), "peak_gbs": peak, **run(dev, peak, tuple(a.cr))}))


if __name__ == "__main__":
    main()
