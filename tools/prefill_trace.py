#!/usr/bin/env python
"""Timeline of the two prefill_metrics passes (diagnostic build: KVC_EXTRA_NVCC_FLAGS=-DKVC_PREFILL_TRACE).
Per CTA: stamps of epilogue thread 0 (0 entry, 1 setup done, 2 Q staged, 3 statistics merged, 4 / 5 first / second
tile's MMAs complete, 6 loop done, 7 exit) and of the control thread (8 TMA issued, 9 Q ready, 10 K tile 0 landed,
11 last MMA issued).  python tools/prefill_trace.py [n_prompts] [prompt_len]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import prefill_bench
from vllm_kvcompress_b200 import native

lib = native.load_lib()
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
buf = torch.zeros((2, 4096, 16), dtype=torch.int64, device=dev)
prefill_bench.run(dev, [L] * n, 32, reps=1)
lib.kvc_debug_set_prefill_trace(ctypes.c_void_p(buf.data_ptr()))
buf.zero_()
prefill_bench.run(dev, [L] * n, 32, reps=1)   # the last call's stamps stay
torch.cuda.synchronize()
lib.kvc_debug_set_prefill_trace(ctypes.c_void_p(0))
t = buf.cpu().numpy()
out = {}
for ps in (0, 1):
    live = t[ps][t[ps][:, 0] > 0]
    t0 = live[:, 0].min()
    rel = (live - t0) / 1e3
    names = ["entry", "setup", "q_staged", "stats_merged", "mma0_done", "mma1_done", "loop_done", "exit",
             "c_tma_issued", "c_q_ready", "c_k0_landed", "c_last_mma"]
    out[f"pass{ps + 1}"] = {"ctas": int(live.shape[0]),
                            "median_us": {nm: round(float(np.median(rel[:, i])), 2) for i, nm in enumerate(names)},
                            "max_us": {nm: round(float(rel[:, i].max()), 2) for i, nm in enumerate(names)}}
    if ps == 0:
        base = t0
    else:
        out["pass2_start_after_pass1_start_us"] = round(float(t0 - base) / 1e3, 2)
print(json.dumps(out, indent=1))
