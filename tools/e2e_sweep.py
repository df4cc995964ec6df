"""e2e step time against the host copy granularity (layers per H2D / D2H copy): python tools/e2e_sweep.py"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for g in (sys.argv[1:] or ["1", "2", "4", "8", "16"]):
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "60", "--warmup", "5",
                          "--no-cpu-baseline", "--no-compaction", "--host-copy-layers", g],
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    print(f"layers per copy {g:>2}: e2e {d['e2e']['value']:9.0f} tokens/s  {d['e2e']['ms_per_step']:.3f} ms   "
          f"(device-resident {d['value']:9.0f})")
