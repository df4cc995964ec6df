"""Phase shares of the cooperative long-context selection kernel (3 prompts x 6000 tokens, Llama-3-8B heads, CR 64):
KVC_SELECT_STOP_AFTER = 1 key pass, 2 four digit passes, 3 counts + grants, 0 whole kernel; mode 3 = seven launches.
python tools/select_long_phases.py"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import compaction_bench as cb
    from vllm_kvcompress_b200 import capi
    dev = torch.device("cuda", 0)
    capi.set_select_mode(int(os.environ.get("SEL_MODE", "0")))
    c = cb.build_schedule(dev, 64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for cold in (True, False):
        ts = cb._timed(c["sel"], 12, setup=(flush.zero_ if cold else None))
        print(f"stop_after={os.environ.get('KVC_SELECT_STOP_AFTER', '0')} mode={os.environ.get('SEL_MODE', '0')} "
              f"{'cold' if cold else 'warm'}: {sorted(ts[2:])[len(ts[2:]) // 2] * 1e3:.1f} us")
else:
    for stop, mode in (("1", "0"), ("2", "0"), ("3", "0"), ("0", "0"), ("0", "3")):
        env = dict(os.environ, KVC_SELECT_STOP_AFTER=stop, SEL_MODE=mode)
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip())
