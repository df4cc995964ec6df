"""ncu target: one prefill-scale K6 call (4 x 6000-token prompts) and three post-prefill K4 calls (profiles/README.md).
python tools/ncu_k6_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import k6_bench
import compaction_bench as cb
dev = torch.device("cuda", 0)
k6_bench.run(dev, [6000] * 4, reps=1)
c = cb.build_schedule(dev, 64)
c["sel"]()
for _ in range(3):
    c["k4"]()
torch.cuda.synchronize()
