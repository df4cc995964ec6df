"""ncu target: four long-context select_evictions calls on 3 prompts x 6000 tokens, CR 64 (profiles/README.md).
python tools/ncu_select_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import compaction_bench as cb
dev = torch.device("cuda", 0)
c = cb.build_schedule(dev, 64)
for _ in range(4):
    c["sel"]()
torch.cuda.synchronize()
