#!/usr/bin/env python
"""bench.py -- KV-Compress hot path on B200: decode tokens/s at CR=64x (Llama-3-8B shape).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl {b200,reference}]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: for the batch of
256 sequences already compressed to max_cache_tokens=128 KVs per (layer, head)
(BASELINE.json configs[1]: Llama-3-8B shape, input 6000 / output 500, CR 64, batch 256):
  compression cycle (fused selection -> schedule_t1_cache_moves -> execute_cache_moves ->
  free blocks) -> append one token per head -> for each of the 32 layers:
  kvcompress_reshape_and_cache + paged attention over the per-head block tables with the
  attention metric aggregated (sum_q p^2) -- i.e. everything LLMEngine.step does on this path
  (vllm/engine/llm_engine.py:1556-1634) except the model's dense projections, which are
  outside the path and replaced by synthetic Q/K/V.
value   = batch tokens / step time, inputs resident in HBM (CUDA events, max over ranks).
e2e     = same step through the public API (DecodeHotPath.step_host) with the step's Q/K/V
          coming from pinned host memory and the attention output copied back, copies inside
          the timed region.
roofline= the dominant kernel (the fused attention step: cache write + deferred move + attention + metric
          aggregation): algorithmic bytes / its mean launch time, measured live with CUDA events on the
          launching stream around graph replays of one step's 32 launches (their deferred work armed,
          engine.profile_attention); traffic = DRAM bytes per launch from the committed ncu capture.
By default new tokens that open a fresh block are parked and stored once at their final slot (deferred
append / deferred move); --no-deferred-append / --no-deferred-move run the reference's op
sequence (store, read back, move) instead.
cpu_baseline = the oracle port of the same step on the host cores, on a bounded sample.
N>1: request-sharded replicas (one independent cache per GPU, weak scaling); NCCL only for
the start-up broadcast and the timing reduction.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-deferred-append", action="store_true",
                    help="store every new token to the cache at once (reference op sequence)")
    ap.add_argument("--no-deferred-move", action="store_true",
                    help="execute_cache_moves stores the parked tokens itself instead of the next attention step")
    ap.add_argument("--batch", type=int, default=256, help="sequences per GPU")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--model", default="llama3-8b",
                    choices=["llama3-8b", "llama3-8b-cr16", "llama3-8b-cr8", "llama3-8b-cr4", "llama3-70b"],
                    help="shape preset: 8b = L32 H8 Q4 CR64 (max_cache_tokens 128); -crN = the CR sweep; "
                         "70b = L80 H8 Q8 CR16 (368)")
    ap.add_argument("--selection", default="fused", choices=["fused", "chain"])
    ap.add_argument("--allocator", default="device", choices=["device", "static"],
                    help="device = blocks freed/allocated every step by the device-side allocator")
    ap.add_argument("--unfused-metrics", action="store_true")
    ap.add_argument("--separate-cache-write", action="store_true", help="K6 as its own launch per layer")
    ap.add_argument("--host-copy-layers", type=int, default=4, help="e2e: layers per H2D / D2H copy")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-compaction", action="store_true")
    ap.add_argument("--cpu-sample-seqs", type=int, default=16)
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the other BASELINE.json configurations (CR 16 / 8 / 4, Llama-3-70B) and the launch-mode A/B")
    ap.add_argument("--extra-steps", type=int, default=12, help="timed steps per extra configuration")
    ap.add_argument("--no-reference-cuda", action="store_true",
                    help="skip the leg that times the reference's own CUDA kernels (oracle/_ref) on the same step")
    return ap.parse_args()


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json),
    or None when no capture of this configuration exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get(kernel)
    return None if d is None else d["dram_bytes_read_per_launch"] + d["dram_bytes_write_per_launch"]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 2 ms from a thread
    (the timed region of the default run is ~0.3 s; nvidia-smi -lms cannot sample that fast), with
    the nvidia-smi recipe of B200_PROFILING.md as the fallback when pynvml is unavailable."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index = index
        self.sm, self.mx, self.bits = [], [], 0
        self.stop_flag = False
        self.thread = None
        self.proc = None

    def _poll_nvml(self, h, nv):
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx.append(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll_nvml, args=(h, nv), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if r and r[0].replace(".", "").isdigit():
                self.sm.append(int(float(r[0])))
            if len(r) > 1 and r[1].replace(".", "").isdigit():
                self.mx.append(int(float(r[1])))
            for (nm, bit), v in zip(self.REASONS, r[3:7]):
                if v.lower().startswith("active"):
                    self.bits |= bit

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self.stop_flag = True
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        self.thread.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": [nm for nm, bit in self.REASONS if self.bits & bit], "samples": len(sm)}


def cpu_step_rate(args, seconds=None, steps=None, warmup=1):
    """The oracle port of the step on the host cores (OpenMP over heads / sequences): tokens/s on a bounded
    sample of the batch.  Imports oracle/ only (no product library is mapped by the CPU arm)."""
    import numpy as np
    from oracle import kvc_oracle as O
    from oracle.cpu_decode_step import CpuDecodeHotPath, CpuStepConfig
    O.build()
    O.use_all_cores()
    ns = args.cpu_sample_seqs
    cfg = CpuStepConfig(num_seqs=ns, num_layers=args.layers)
    eng = CpuDecodeHotPath(cfg)
    Hq, H, D = cfg.num_kv_heads * cfg.queries_per_kv, cfg.num_kv_heads, cfg.head_size
    qkv = (np.random.RandomState(0).rand(cfg.num_layers, ns, (Hq + 2 * H) * D).astype(np.float16) - 0.5).view(np.uint16)
    out = np.zeros((cfg.num_layers, ns, Hq * D), np.float32)
    for _ in range(warmup):
        eng.step(qkv, out)
    t0 = time.perf_counter()
    reps = 0
    while (steps is not None and reps < steps) or (steps is None and (reps < 3 or (time.perf_counter() - t0 < seconds
                                                                                   and reps < 200))):
        eng.step(qkv, out)
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    kv_mb = ns * cfg.num_layers * H * cfg.max_cache_tokens * 2 * D * 2 / 1e6
    return {"value": ns / dt, "unit": "tokens/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"{ns} of {args.batch} sequences per step (K+V working set {kv_mb:.0f} MB), all {args.layers} "
                      f"layers, full cycle (selection, moves, append, cache write, attention, metric aggregation), "
                      f"{reps} steps in {dt * reps:.1f} s"}, dt


def cpu_reference_kernels(args, seconds):
    """The reference's OWN CPU kernels (vLLM csrc/cpu paged_attention_v1 + reshape_and_cache, oracle/_ref/
    _vllm_cpu_ref.so) on the per-layer attention part of the step: what exists of "the reference's CPU path" for
    this hot path (the fork has no CPU build of the KV-Compress ops).  Reported next to cpu_baseline (the full-step
    port); see oracle/cpu_reference_step.py."""
    try:
        from oracle import cpu_reference_step as C
        if not C.available():
            return {"unavailable": "oracle/_ref/_vllm_cpu_ref.so not built"}
        return C.measure(num_seqs=args.cpu_sample_seqs, num_layers=args.layers, seconds=seconds,
                         threads=os.cpu_count())
    except Exception as ex:  # pragma: no cover
        return {"error": repr(ex)}


def run_reference(args, rank, world):
    """--impl reference: the oracle port of the step on the host cores (rank 0 only).  The reference has no CPU
    build of these ops (csrc/cpu registers none of them), so the port is the CPU arm."""
    if rank != 0:
        return
    base, dt = cpu_step_rate(args, steps=args.steps, warmup=args.warmup)
    val = base["value"]
    line = {
        "impl": "reference", "metric": "decode tokens/sec at CR=64x (Llama-3-8B shape, KV-Compress hot path)",
        "value": val, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": workload_config(args, args.batch * args.gpus),
        "cpu_baseline": base,
        "cpu_reference_kernels": cpu_reference_kernels(args, seconds=5),
        "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# shape presets: BASELINE.json configs[1] (the headline), configs[3] (70B) and configs[4] (CR sweep); max_cache_tokens per
# compression rate follows benchmarks/benchmark_throughput.py:391-393 at input 6000: max(128, ((6000 // CR - 1) // 16) * 16)
PRESETS = {
    "llama3-8b": dict(name="Llama-3-8B", layers=32, q=4, cr=64, max_cache_tokens=128, blocks_per_head=12),
    "llama3-8b-cr16": dict(name="Llama-3-8B", layers=32, q=4, cr=16, max_cache_tokens=368, blocks_per_head=28),
    "llama3-8b-cr8": dict(name="Llama-3-8B", layers=32, q=4, cr=8, max_cache_tokens=736, blocks_per_head=52),
    "llama3-8b-cr4": dict(name="Llama-3-8B", layers=32, q=4, cr=4, max_cache_tokens=1488, blocks_per_head=100),
    "llama3-70b": dict(name="Llama-3-70B", layers=80, q=8, cr=16, max_cache_tokens=368, blocks_per_head=28),
}


def workload_config(args, batch, model=None):
    p = PRESETS[model or getattr(args, "model", "llama3-8b")]
    layers = p["layers"] if model or args.layers == 32 else args.layers
    wl = ("%s shape (L=%d, H_kv=8, Q=%d, D=128, block 16), input 6000 / output 500, CR %dx => max_cache_tokens=%d, "
          "protected_window=32, continual compression every step"
          % (p["name"], layers, p["q"], p["cr"], p["max_cache_tokens"]))
    return {"workload": wl,
            "global_batch": batch, "seq_len": 6000, "parallelism": "request-sharded replicas",
            "l2": "inputs larger than L2 (per-step KV read ~4.6 GB vs 126 MB L2)",
            "selection": args.selection, "allocator": getattr(args, "allocator", "device"),
            "deferred_append": not getattr(args, "no_deferred_append", False),
            "deferred_move": not getattr(args, "no_deferred_move", False),
            "fused_metrics": not args.unfused_metrics,
            "fused_cache_write": not getattr(args, "separate_cache_write", False),
            "launch": "one CUDA graph per step" if args.selection == "fused" else "eager"}


def compaction_roofline(device, peak):
    """execute_cache_moves on the post-prefill shape of BASELINE.md config 5, CR in {64, 16, 8, 4}: algorithmic
    GB/s (1048 B per move) vs HBM peak, the sector floor of the move list and the committed ncu traffic
    (tools/compaction_bench.py: 5 warm-up + 20 timed repetitions, each on a re-initialised cache)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import compaction_bench
    return compaction_bench.run(device, peak)


def prefill_metrics_bench(device):
    """Prefill metric initialisation (tcgen05 kernels) on 4 prompts of 6000 tokens, Llama-3-8B heads:
    2 passes x 2*rows*keys*D FLOPs on the tensor cores (rows = 32 observed queries x 32 query heads)."""
    from vllm_kvcompress_b200 import _custom_ops as ops
    lens, Hq, H, D = [6000] * 4, 32, 8, 128
    T = sum(lens)
    g = torch.Generator(device=device).manual_seed(3)
    q = torch.randn((T, Hq, D), generator=g, device=device, dtype=torch.float32).to(torch.float16)
    k = torch.randn((T, H, D), generator=g, device=device, dtype=torch.float32).to(torch.float16)
    buf = torch.zeros(len(lens), dtype=torch.int32, device=device)
    starts = torch.tensor([0] + list(torch.tensor(lens).cumsum(0).tolist()), dtype=torch.int32, device=device)
    # slots as the engine lays them out: every head owns randomly placed blocks of 16 consecutive tokens
    nblk = (T + 15) // 16
    blocks = torch.randperm(H * nblk, device=device).view(H, nblk)
    tok = torch.arange(T, device=device)
    slots = (blocks[:, tok // 16] * 16 + tok % 16).t().contiguous()
    metrics = torch.zeros((H * nblk, 16), dtype=torch.float32, device=device)
    ts = []
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.prefill_kv_metrics(q, k, lens, D ** -0.5, buf, n_observed=32, kv_metrics=metrics, slot_mapping=slots,
                               return_metrics=False, seq_start_loc=starts)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts[1:])[len(ts[1:]) // 2]
    flops = 2 * sum(2.0 * 32 * Hq * n * D for n in lens)
    return {"prompts": lens, "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12,
            "note": "3 launches (TMA + tcgen05 row stats, squared-probability column sums, max-pool + aggregate_prefill); "
                    "prompt starts on the device as in the reference's prefill metadata"}


def build_engine(args, model, device, rank, pdl="model"):
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig
    p = PRESETS[model]
    layers = p["layers"] if (model != "llama3-8b" or args.layers == 32) else args.layers
    cfg = HotPathConfig(num_seqs=args.batch, num_layers=layers, queries_per_kv=p["q"],
                        max_cache_tokens=p["max_cache_tokens"], blocks_per_head=p["blocks_per_head"],
                        selection=args.selection, fused_metrics=not args.unfused_metrics, allocator=args.allocator,
                        host_copy_layers=args.host_copy_layers, fused_cache_write=not args.separate_cache_write,
                        seed=rank, deferred_append=not args.no_deferred_append,
                        deferred_move=not args.no_deferred_move, attention_pdl=pdl)
    return DecodeHotPath(cfg, device)


def make_inputs(cfg, device, rank):
    H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
    g = torch.Generator(device=device).manual_seed(1234 + rank)
    qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), generator=g, device=device, dtype=torch.float32) - 0.5).to(cfg.dtype)
    out = torch.empty((L, S, H * Q * D), device=device, dtype=cfg.dtype)
    return qkv, out


def time_device_steps(eng, qkv, out, steps, warmup, barrier, sampler=None):
    """`steps` steps with device-resident inputs between two CUDA events (one CUDA graph per step)."""
    graphed = eng.capture(qkv, out)
    run_step = eng.replay if graphed else (lambda: eng.step(qkv, out))
    for _ in range(warmup):
        run_step()
    barrier()
    if sampler is not None:
        sampler.start()
    n0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run_step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps, (eng.launches - n0) // steps


def attention_roofline(eng, qkv, out, steps, peak, peak_src, traffic=None):
    ms, nbytes = eng.profile_attention(qkv, out, steps=steps)
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm",
            "kernel": "paged_attention_stream_kernel (cache write + deferred moves + attention + metric aggregation), "
                      "%d launches/step" % eng.cfg.num_layers,
            "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
            "bytes_per_launch": nbytes, "ms_per_launch": ms, "traffic": traffic,
            "launch_mode": "kvc_set_attention_pdl %d (layer 0) / %d (layers >= 1)" % (eng.pdl_first, eng.pdl_rest)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    from vllm_kvcompress_b200 import sharding

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = peaks()
    eng = build_engine(args, args.model, device, rank)
    cfg = eng.cfg
    # the only collective on this path: one start-up broadcast (stands in for the weight broadcast
    # of the 8-replica layout); every replica then owns an independent cache for its own requests
    hdr = sharding.broadcast_config([cfg.num_seqs, cfg.num_layers, cfg.max_cache_tokens], device)
    assert hdr == [cfg.num_seqs, cfg.num_layers, cfg.max_cache_tokens]
    my_requests = sharding.shard_requests(cfg.num_seqs * world, world, rank)  # ids of this replica's batch
    assert len(my_requests) == cfg.num_seqs
    L, S = cfg.num_layers, cfg.num_seqs
    qkv, out = make_inputs(cfg, device, rank)

    # ---------------- device-resident timing (headline `value`) ----------------
    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = time_device_steps(eng, qkv, out, args.steps, args.warmup, barrier, sampler)
    clocks = sampler.stop() if rank == 0 else None
    eng.check_status()

    # attention kernel time: the step's L launches alone as one CUDA graph between two events (separate pass
    # so the event records do not perturb the headline number)
    default_cfg = (args.model == "llama3-8b" and args.batch == 256 and not args.no_deferred_append
                   and not args.no_deferred_move)
    roof = attention_roofline(eng, qkv, out, min(args.steps, 5), peak, peak_src,
                              ncu_traffic("paged_attention_stream_kernel") if default_cfg else None)
    pdl_ab = None
    if rank == 0 and not args.no_extras and "KVC_ATTN_PDL" not in os.environ:
        # A/B of the launch modes on the same state (a model can only use "model"; "independent" is what round 1
        # reported: every layer's q/k/v declared independent of the previous kernel)
        pdl_ab = {"model": roof["ms_per_launch"]}
        keep = eng.pdl_first, eng.pdl_rest
        for name, (a, b) in (("independent", (1, 3)), ("safe", (1, 1)), ("off", (0, 0))):
            eng.pdl_first, eng.pdl_rest = a, b
            pdl_ab[name] = eng.profile_attention(qkv, out, steps=min(args.steps, 5))[0]
        eng.pdl_first, eng.pdl_rest = keep

    # ---------------- end to end through the public API with host buffers ----------------
    # two sets of pinned host buffers alternate (the pipelined API returns before a step's output has landed)
    qkv_host = [torch.empty(qkv.shape, dtype=cfg.dtype, pin_memory=True) for _ in range(2)]
    out_host = [torch.empty(out.shape, dtype=cfg.dtype, pin_memory=True) for _ in range(2)]
    for t in qkv_host:
        t.copy_(qkv)
    for i in range(max(4, min(args.warmup, 6))):       # includes the two eager + capture steps
        eng.submit_host(qkv_host[i & 1], out_host[i & 1])
    eng.wait_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.submit_host(qkv_host[i & 1], out_host[i & 1])
    eng.wait_host()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    eng.check_status()
    h2d_bytes = qkv_host[0].numel() * qkv_host[0].element_size()
    d2h_bytes = out_host[0].numel() * out_host[0].element_size()
    del qkv_host, out_host

    ms, ms_e2e = sharding.max_over_ranks([ms, ms_e2e], device)
    total_tokens = S * world
    line = {
        "metric": "decode tokens/sec at CR=64x (Llama-3-8B shape, KV-Compress hot path)",
        "value": total_tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(args, total_tokens),
        "e2e": {"value": total_tokens / (ms_e2e * 1e-3), "unit": "tokens/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "api": "DecodeHotPath.submit_host / wait_host: pipelined over two staging buffers (H2D of step "
                       "t+1 and D2H of step t-1 under the compute of step t)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "attention_launch": "programmatic dependent launch, layer 0 mode %d, layers >= 1 mode %d "
                            "(include/kvc_b200.h kvc_set_attention_pdl)" % (eng.pdl_first, eng.pdl_rest),
    }
    if pdl_ab is not None:
        line["attention_ms_per_launch_by_pdl"] = pdl_ab
    del eng, qkv, out
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE.json configurations (decode step, same measurement) ----------------
    if not args.no_extras and args.model == "llama3-8b":
        extras = {}
        for model in ("llama3-8b-cr16", "llama3-8b-cr8", "llama3-8b-cr4", "llama3-70b"):
            try:
                e = build_engine(args, model, device, rank)
                q2, o2 = make_inputs(e.cfg, device, rank)
                xsteps = max(4, min(args.steps, args.extra_steps))
                xms, xl = time_device_steps(e, q2, o2, xsteps, 3, barrier)
                e.check_status()
                xroof = attention_roofline(e, q2, o2, 3, peak, peak_src)
                xms, = sharding.max_over_ranks([xms], device)
                extras[model] = {"value": total_tokens / (xms * 1e-3), "unit": "tokens/s", "ms_per_step": xms,
                                 "steps": xsteps, "warmup": 3, "gpu_launches": xl, "roofline": xroof,
                                 "config": workload_config(args, total_tokens, model)["workload"],
                                 "cache_gb": e.st.NB * 8192 / 1e9}
                del e, q2, o2
            except Exception as ex:  # pragma: no cover
                extras[model] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        line["configs"] = extras

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if not args.no_compaction:
        try:
            line["execute_cache_moves"] = compaction_roofline(device, peak)
        except Exception as ex:  # pragma: no cover
            line["execute_cache_moves"] = {"error": repr(ex)}
        try:
            line["prefill_kv_metrics"] = prefill_metrics_bench(device)
        except Exception as ex:  # pragma: no cover
            line["prefill_kv_metrics"] = {"error": repr(ex)}
        try:
            # the other KV-Compress ops around a prefill: the per-head cache write of 4 x 6000-token prompts and the
            # allocator calls of the compression that follows it (tools/k6_bench.py, tools/alloc_bench.py)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import alloc_bench
            import k6_bench
            line["prefill_cache_write"] = k6_bench.run(device, [6000] * 4)
            line["allocator_post_prefill"] = alloc_bench.measure()
        except Exception as ex:  # pragma: no cover
            line["prefill_cache_write"] = {"error": repr(ex)}
    if not args.no_reference_cuda and world == 1 and args.model == "llama3-8b":
        # the fork's own CUDA kernels (unmodified .cu files rebuilt for sm_100a, oracle/_ref) driving the same
        # step in the fork's op order: the "reference fork's own decode tokens/s on the same box" (north_star).
        # Outside the product's timed region; reported, like cpu_baseline.
        try:
            from oracle import ref_ops
            if ref_ops.available():
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_reference_kernels as brk
                r = brk.measure(args.batch, args.layers, steps=3, warmup=1, device=device)
                r["ratio_value_over_reference_cuda"] = line["value"] / r["value"]
                line["reference_cuda"] = r
            else:
                line["reference_cuda"] = {"unavailable": "oracle/_ref/_kvc_ref.so not built"}
        except Exception as ex:  # pragma: no cover
            line["reference_cuda"] = {"error": repr(ex)}
    if not args.no_cpu_baseline and world == 1:   # reported at N=1 only (rank 0's host cores)
        line["cpu_baseline"] = cpu_step_rate(args, seconds=10)[0]
        line["cpu_reference_kernels"] = cpu_reference_kernels(args, seconds=5)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
