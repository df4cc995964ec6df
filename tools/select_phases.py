"""Time select_evictions at the benchmark's steady state; with KVC_SELECT_STOP_AFTER=n the one-CTA-per-sequence
kernel returns after phase n (1 staging, 2 bounds, 3 T1, 4 pruned sort, 5 grants), so differences give the
phases' shares.  python tools/select_phases.py"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from vllm_kvcompress_b200 import _custom_ops as ops
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig, MAX_INT
    dev = torch.device("cuda", 0)
    stop = os.environ.pop("KVC_SELECT_STOP_AFTER", None)
    cfg = HotPathConfig()
    eng = DecodeHotPath(cfg, dev)
    H, Q, D, L, S, B = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs, cfg.block_size
    qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), device=dev) - 0.5).to(cfg.dtype)
    out = torch.empty((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    for _ in range(12):
        eng.step(qkv, out)
    st = eng.st
    max_blocks = (cfg.max_cache_tokens * L * H + B - 1) // B
    ops.prepare_compression(st.context_lens, eng.hanging, eng.offs, eng.budgets, B, max_blocks)
    if stop is not None:
        os.environ["KVC_SELECT_STOP_AFTER"] = stop
    ts = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(8):
        flush.zero_()   # metrics / positions come from DRAM as in the real step
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.select_evictions(eng.evicted_logical, eng.evicted_kv_count, eng.evicted_block_count, st.kv_metrics,
                             st.kv_position, st.block_tables, st.context_lens, eng.hanging, eng.offs, eng.budgets,
                             eng.last_pos, eng.protected, B, MAX_INT)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"stop_after={stop}: {sorted(ts)[len(ts) // 2]:.1f} us")
else:
    for stop in ("1", "2", "3", "4", "5", "0"):
        env = dict(os.environ, KVC_SELECT_STOP_AFTER=stop)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
