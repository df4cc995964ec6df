"""Per-step trace of the decode hot path (moves, evicted tokens, heads on a block boundary, step time):
    python tools/step_trace.py [steps]
shows how far the synthetic state is from the steady state of continual compression."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
cfg = HotPathConfig()
eng = DecodeHotPath(cfg, dev)
H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), device=dev) - 0.5).to(cfg.dtype)
out = torch.empty((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
for i in range(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.step(qkv, out)
    e1.record()
    torch.cuda.synchronize()
    ctx = eng.st.context_lens
    print(f"step {i:3d}: {e0.elapsed_time(e1):7.3f} ms  moves {int(eng.moves_count.sum()):8d}  evicted kv "
          f"{int(eng.evicted_kv_count.sum()):8d}  evicted blocks {int(eng.evicted_block_count.sum()):7d}  "
          f"ctx mean {float(ctx.float().mean()):6.1f} min {int(ctx.min())} max {int(ctx.max())}  "
          f"heads at 16k+1: {float(((ctx % 16) == 1).float().mean()) * 100:5.1f} %")
