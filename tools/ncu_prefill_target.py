"""ncu target: a few prefill_kv_metrics calls on n prompts of L tokens (profiles/README.md).
python tools/ncu_prefill_target.py 4 6000"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, prefill_bench
dev = torch.device("cuda", 0)
n = int(sys.argv[1]); L = int(sys.argv[2])
print(json.dumps(prefill_bench.run(dev, [L] * n, 32, reps=2)))
