"""Device time of every phase of one decode step (eager launches, CUDA events between phases; medians of the LAST 10 steps):
    python tools/step_breakdown.py [steps] [--model llama3-70b]
The stream is primed with a spin kernel before each step so the host runs ahead of the device and the
events see kernel time, not launch gaps."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_kvcompress_b200 import _custom_ops as ops
from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig, MAX_INT

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 40
dev = torch.device("cuda", 0)
preset = {}
if "llama3-70b" in sys.argv:
    preset = dict(num_layers=80, queries_per_kv=8, max_cache_tokens=368, blocks_per_head=28)
if "--no-deferred" in sys.argv:
    preset["deferred_append"] = False
if "--no-deferred-move" in sys.argv:
    preset["deferred_move"] = False
cfg = HotPathConfig(**preset)
eng = DecodeHotPath(cfg, dev)
H, Q, D, L, S, B = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs, cfg.block_size
qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), device=dev) - 0.5).to(cfg.dtype)
out = torch.empty((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
st = eng.st
names = ["prepare", "select", "sched_moves", "exec_moves", "free", "append", "attention x%d" % L]
acc = {n: [] for n in names}
for it in range(steps):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    torch.cuda._sleep(3_000_000)
    ctx = st.context_lens
    last_pos = eng.last_pos
    max_blocks = eng.max_cache_blocks
    evs[0].record()
    ops.prepare_compression(ctx, eng.hanging, eng.offs, eng.budgets, B, max_blocks)
    evs[1].record()
    ops.select_evictions(eng.evicted_logical, eng.evicted_kv_count, eng.evicted_block_count, st.kv_metrics,
                         st.kv_position, st.block_tables, ctx, eng.hanging, eng.offs, eng.budgets, last_pos,
                         eng.protected, B, MAX_INT)
    evs[2].record()
    torch.ops._C_kvc_ops.schedule_t1_cache_moves(eng.cache_moves, eng.moves_count, eng.evicted_logical,
                                                 eng.evicted_kv_count, eng.offs, st.block_tables, ctx, B)
    evs[3].record()
    if eng.deferred:
        ops.execute_cache_moves_deferred(st.k_cache, st.v_cache, st.kv_metrics, st.kv_position, eng.cache_moves,
                                         eng.moves_count, eng.offs, eng.pend_kv, eng.pend_slot, eng.pend_dst)
    else:
        ops.execute_cache_moves(st.k_cache, st.v_cache, st.kv_metrics, st.kv_position, eng.cache_moves,
                                eng.moves_count, eng.offs, 1, 16)
    evs[4].record()
    eng.alloc.free_compressed(None, eng.evicted_block_count)
    evs[5].record()
    eng.alloc.append_slots(None, eng.seq_len, eng.slot_mapping)
    eng._pos2.add_(1)
    evs[6].record()
    eng.forward(qkv, out)
    evs[7].record()
    torch.cuda.synchronize()
    for i, n in enumerate(names):
        acc[n].append(evs[i].elapsed_time(evs[i + 1]) * 1e3)
    if it in (0, 1, 2, 5, 10, steps - 1):
        print(f"step {it:3d}: moves {int(eng.moves_count.sum()):8d} evicted kv {int(eng.evicted_kv_count.sum()):8d} "
              f"blocks {int(eng.evicted_block_count.sum()):7d} ctx mean {float(st.context_lens.float().mean()):6.1f} "
              f"min {int(st.context_lens.min())} max {int(st.context_lens.max())} "
              f"at 16k+1: {float(((st.context_lens % 16) == 1).float().mean()) * 100:5.1f} %")
print("heads with > 1 move at the last step:", int((eng.moves_count > 1).sum()))
tot = 0.0
for n in names:
    v = sorted(acc[n][-10:])
    med = v[len(v) // 2]
    tot += med
    print(f"{n:16s} {med:9.1f} us")
print(f"{'sum':16s} {tot:9.1f} us")
