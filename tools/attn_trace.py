#!/usr/bin/env python
"""Timeline of the fused attention launches inside one decode step (diagnostic build only).

Build the library with the trace hooks, run, rebuild without:
    KVC_EXTRA_NVCC_FLAGS=-DKVC_ATTN_TRACE python -c "import __graft_entry__ as g; g.build()"   (touch the .cu first)
    python tools/attn_trace.py [--pdl model] [--steps 70]
Every consumer warp (= one (sequence, kv head) item) records: globaltimer at entry / exit and clock64 at entry,
before / after the griddepcontrol.wait, before the first K unit, after the first K block, at the end of the
sweep.  The step's L launches are captured in one CUDA graph; each launch writes its own slice of the buffer.
Prints, per layer (median over layers 2..L-1): launch span, gap to the previous launch, and the split of a warp's
life into wait / first-data / sweep / epilogue."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pdl", default="model")
    ap.add_argument("--steps", type=int, default=70)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=32)
    a = ap.parse_args()
    from vllm_kvcompress_b200 import native
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig
    lib = native.load_lib()
    if not hasattr(lib, "kvc_debug_set_attention_trace"):
        raise SystemExit("library was built without -DKVC_ATTN_TRACE")
    dev = torch.device("cuda", 0)
    cfg = HotPathConfig(num_seqs=a.batch, num_layers=a.layers, attention_pdl=a.pdl)
    eng = DecodeHotPath(cfg, dev)
    H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
    qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), device=dev) - 0.5).to(cfg.dtype)
    out = torch.empty((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    for _ in range(a.steps):
        eng.step(qkv, out)
    torch.cuda.synchronize()
    items = S * H
    trace = torch.zeros((L, items, 16), dtype=torch.int64, device=dev)
    base = trace.data_ptr()
    orig = eng.forward_layer

    def traced_layer(l, *args, **kw):
        lib.kvc_debug_set_attention_trace(ctypes.c_void_p(base + l * items * 128))
        return orig(l, *args, **kw)

    eng.forward_layer = traced_layer
    eng.step(qkv, out)            # eager once (allocator warm, trace pointers exercised)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.step(qkv, out)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    trace.zero_()
    g.replay()
    torch.cuda.synchronize()
    lib.kvc_debug_set_attention_trace(ctypes.c_void_p(0))
    t = trace.cpu().numpy().astype(np.int64)
    g0, c0, cw0, cw1, ck0, ck1 = t[..., 0], t[..., 1], t[..., 2], t[..., 3], t[..., 4], t[..., 5]
    c6 = t[..., 6] & ((1 << 48) - 1)
    g7 = t[..., 7]
    t0 = g0.min()
    rows = []
    for l in range(L):
        rows.append(dict(layer=l, start_us=(g0[l].min() - t0) / 1e3, first_exit_us=(g7[l].min() - t0) / 1e3,
                         last_start_us=(g0[l].max() - t0) / 1e3, end_us=(g7[l].max() - t0) / 1e3))
    span = np.array([r["end_us"] - r["start_us"] for r in rows])
    period = np.diff(np.array([r["end_us"] for r in rows]))
    clk = 1.965e3  # cycles per us at the boost clock (ratios matter, not the absolute)
    sel = slice(2, L)
    phases = {
        "entry_to_wait_us": np.median((cw0 - c0)[sel]) / clk,
        "pdl_wait_us": np.median((cw1 - cw0)[sel]) / clk,
        "q_load_etc_us": np.median((ck0 - cw1)[sel]) / clk,
        "first_k_block_us": np.median((ck1 - ck0)[sel]) / clk,
        "sweep_rest_us": np.median((c6 - ck1)[sel]) / clk,
        "output_stores_us": np.median((t[..., 8] - c6)[sel]) / clk,
        "new_token_stores_us": np.median((t[..., 9] - t[..., 8])[sel]) / clk,
        "metric_epilogue_us": np.median((t[..., 10] - t[..., 9])[sel]) / clk,
        "warp_life_us_median": np.median((g7 - g0)[sel]) / 1e3,
        "warp_life_us_p95": np.percentile((g7 - g0)[sel], 95) / 1e3,
    }
    # where does the exit skew come from?  per-SM exit time of one mid-step layer, relative to the layer's first
    # warp entry; SM ids are dealt to GPCs round-robin-ish, so a pattern over the id shows a topology effect
    lay = min(L - 1, 8)
    smid = (t[lay, :, 6] >> 48) & 0xffff
    rel_exit = (g7[lay] - g0[lay].min()) / 1e3
    rel_entry = (g0[lay] - g0[lay].min()) / 1e3
    per_sm = {}
    for sm, ex in zip(smid.tolist(), rel_exit.tolist()):
        per_sm.setdefault(sm, []).append(ex)
    sm_last = {k: max(v) for k, v in per_sm.items()}
    sm_cnt = {k: len(v) for k, v in per_sm.items()}
    order = sorted(sm_last, key=lambda k: sm_last[k])
    skew = {"layer": lay, "exit_us_percentiles": [float(np.percentile(rel_exit, q)) for q in (0, 5, 25, 50, 75, 95, 100)],
            "entry_us_percentiles": [float(np.percentile(rel_entry, q)) for q in (0, 50, 95, 100)],
            "sm_last_exit_percentiles": [float(np.percentile(list(sm_last.values()), q)) for q in (0, 25, 50, 75, 100)],
            "fastest_sms": [(k, round(sm_last[k], 1), sm_cnt[k]) for k in order[:10]],
            "slowest_sms": [(k, round(sm_last[k], 1), sm_cnt[k]) for k in order[-10:]],
            "warps_per_sm_hist": {str(c): int(sum(1 for v in sm_cnt.values() if v == c)) for c in sorted(set(sm_cnt.values()))}}
    # sweep-time variance: between SMs vs within an SM, persistence of the per-SM effect across layers, and the
    # effect of the warp's position in its CTA (item = cta * consumers + warp)
    sweep = (c6 - ck0) / clk                                    # [L, items] us
    sm_all = (t[..., 6] >> 48) & 0xffff
    lays = list(range(2, L))
    sm_mean = np.full((L, 160), np.nan)
    within = []
    for l in lays:
        for sm in np.unique(sm_all[l]):
            v = sweep[l][sm_all[l] == sm]
            sm_mean[l, sm] = v.mean()
            within.append(v.var())
    between = np.nanvar(sm_mean[lays], axis=1).mean()
    persist = []
    for l in lays[:-1]:
        x, y = sm_mean[l], sm_mean[l + 1]
        ok = ~np.isnan(x) & ~np.isnan(y)
        persist.append(float(np.corrcoef(x[ok], y[ok])[0, 1]))
    kcons = 7
    by_warp = [float(sweep[lays][:, w::kcons].mean()) for w in range(kcons)]
    sm_avg = np.nanmean(sm_mean[lays], axis=0)
    ids = np.argsort(np.nan_to_num(sm_avg, nan=0))
    ids = [int(i) for i in ids if not np.isnan(sm_avg[i])]
    skew["sweep_us_percentiles"] = [float(np.percentile(sweep[lays], q)) for q in (0, 5, 25, 50, 75, 95, 100)]
    skew["sweep_var_between_sm"] = float(between)
    skew["sweep_var_within_sm"] = float(np.mean(within))
    skew["sm_mean_corr_adjacent_layers"] = float(np.mean(persist))
    skew["sweep_us_by_warp_in_cta"] = by_warp
    skew["sm_avg_sweep_fastest"] = [(i, round(float(sm_avg[i]), 2)) for i in ids[:12]]
    skew["sm_avg_sweep_slowest"] = [(i, round(float(sm_avg[i]), 2)) for i in ids[-12:]]
    skew["sm_avg_sweep_by_id"] = [round(float(v), 2) if not np.isnan(v) else None for v in sm_avg[:148]]
    print(json.dumps({"pdl": a.pdl, "skew": skew, "launch_span_us_median": float(np.median(span[2:])),
                      "launch_period_us_median": float(np.median(period[2:])),
                      "start_skew_us_median": float(np.median([r["last_start_us"] - r["start_us"] for r in rows[2:]])),
                      "exit_skew_us_median": float(np.median([r["end_us"] - r["first_exit_us"] for r in rows[2:]])),
                      "phases": {k: float(v) for k, v in phases.items()}, "layers": rows[:6]}))


if __name__ == "__main__":
    main()
