"""GPU parity tests (run with -m gpu on the B200 box).  Everything here calls the CUDA
kernels through the C ABI (vllm_kvcompress_b200.capi -> libkvc_b200.so) and checks them
against (1) the committed golden vectors produced by the reference's Python oracles,
(2) the C oracle on seeded random inputs, and (3) the reference's own CUDA kernels
(oracle/_ref/_kvc_ref.so) when that library is present.  Integer / byte results must be
bit-exact; attention is checked at the reference's tolerance (atol 1e-3, rtol 1e-5 on the
output; default allclose on recorded probabilities)."""
import os

import numpy as np
import pytest
import torch

from tests.golden_util import load_cases

pytestmark = pytest.mark.gpu

MAX_INT = 2147483000
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "_kvc_ref.so")


@pytest.fixture(scope="module")
def capi():
    from vllm_kvcompress_b200 import capi as c
    assert c.abi_version() == 1
    return c


@pytest.fixture(scope="module")
def O():
    from oracle import kvc_oracle
    kvc_oracle.build()
    return kvc_oracle


@pytest.fixture(scope="module")
def ref_ops():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/_kvc_ref.so (reference CUDA kernels) not built")
    torch.ops.load_library(REF_SO)
    return torch.ops._kvc_ref


def cu(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def cu16(a, is_bf16):
    """uint16 numpy bits -> fp16/bf16 cuda tensor"""
    t = torch.as_tensor(np.ascontiguousarray(a).view(np.int16)).cuda()
    return t.view(torch.bfloat16 if is_bf16 else torch.float16)


def bits16(t):
    return t.view(torch.int16).cpu().numpy().view(np.uint16)


# ---------------------------------------------------------------------------
# A. golden vectors through the C ABI
# ---------------------------------------------------------------------------
EVICT = load_cases("evictions.npz")
COUNT = load_cases("count_blocks.npz")
MOVES = load_cases("moves.npz")
ATTN = load_cases("attention.npz")
CACHE = load_cases("cache_write.npz")


def run_k1(capi, c, budgets=None, truncate=None, evenly=False, control=None, remaining=None):
    S, L, H, B = int(c["S"]), int(c["L"]), int(c["H"]), int(c["B"])
    N = int(c["total_kv"])
    ekv = torch.zeros(N, dtype=torch.int32, device="cuda")
    elog = torch.full((N,), -7, dtype=torch.int32, device="cuda")
    cnt = torch.full((S, L, H), 99, dtype=torch.int32, device="cuda")
    rem = cu(c["hanging_token_count"] if remaining is None else remaining, torch.int32)
    i = lambda k: cu(c[k], torch.int32)
    capi.schedule_cache_evictions(
        ekv, elog, cnt, rem, i("evicted_kv_offsets"), i("sorted_indices"), i("seq_block_offsets"),
        i("layer_by_block"), i("head_by_block"), i("virtual_block_num_by_block"),
        cu(c["budgets"] if budgets is None else budgets, torch.int32), i("context_lens"),
        i("hanging_token_count"), i("kv_position"), i("last_position"), i("protected"), B, evenly,
        None if control is None else cu(control, torch.int32), N, MAX_INT,
        bool(c["truncate"]) if truncate is None else truncate)
    torch.cuda.synchronize()
    return ekv.cpu().numpy(), elog.cpu().numpy(), cnt.cpu().numpy(), rem.cpu().numpy()


@pytest.mark.parametrize("name", sorted(EVICT))
def test_k1_golden(capi, name):
    c = EVICT[name]
    ekv, elog, cnt, _ = run_k1(capi, c)
    np.testing.assert_array_equal(cnt, c["out_evicted_kv_count"])
    if int(c["truncate"]):
        np.testing.assert_array_equal(elog, c["out_evicted_logical_indices"])
        np.testing.assert_array_equal(ekv, c["out_evicted_kv_indices"])
    else:
        offs = c["evicted_kv_offsets"].reshape(-1)
        written = np.zeros(len(elog), bool)
        for o, n in zip(offs, cnt.reshape(-1)):
            written[o:o + n] = True
        np.testing.assert_array_equal(elog[written], c["out_evicted_logical_indices"][written])
        np.testing.assert_array_equal(ekv[written], c["out_evicted_kv_indices"][written])
        assert (elog[~written] == -7).all() and (ekv[~written] == 0).all()


@pytest.mark.parametrize("name", sorted(COUNT))
def test_k3_golden(capi, name):
    c = COUNT[name]
    logical = cu(c["in_evicted_logical_indices"], torch.int32)
    out = torch.zeros(tuple(c["out_evicted_block_count"].shape), dtype=torch.int32, device="cuda")
    capi.count_block_evictions(out, logical, cu(c["evicted_kv_offsets"], torch.int32),
                               cu(c["hanging_token_count"], torch.int32), int(c["B"]), MAX_INT)
    np.testing.assert_array_equal(out.cpu().numpy(), c["out_evicted_block_count"])
    np.testing.assert_array_equal(logical.cpu().numpy(), c["out_evicted_logical_indices"])


@pytest.fixture(params=[0, 1], ids=["register_gather", "staged_tma"])
def moves_mode(request, capi):
    """execute_cache_moves kernel choice for post-prefill shapes (kvc_set_moves_mode): both must be bit-exact"""
    capi.set_moves_mode(request.param)
    yield request.param
    capi.set_moves_mode(0)


def torch_cache(a):
    if a.dtype == np.uint16:
        return cu16(a, False)
    return cu(a)


def cache_bits(t, like):
    return bits16(t) if like.dtype == np.uint16 else t.cpu().numpy()


@pytest.mark.parametrize("name", sorted(MOVES))
def test_k4_k5_golden(capi, name, moves_mode):
    c = MOVES[name]
    B = int(c["B"])
    moves = torch.zeros((int(c["total_kv"]), 2), dtype=torch.int32, device="cuda")
    mcnt = torch.zeros(tuple(c["evicted_kv_count"].shape), dtype=torch.int32, device="cuda")
    offs = cu(c["evicted_kv_offsets"], torch.int32)
    capi.schedule_t1_cache_moves(moves, mcnt, cu(c["evicted_logical_indices"], torch.int32),
                                 cu(c["evicted_kv_count"], torch.int32), offs,
                                 cu(c["block_tables"], torch.int32), cu(c["context_lens_lsh"], torch.int32), B)
    np.testing.assert_array_equal(mcnt.cpu().numpy(), c["out_cache_moves_count"])
    np.testing.assert_array_equal(moves.cpu().numpy(), c["out_cache_moves_idx"])
    kc, vc = torch_cache(c["k_cache"]), torch_cache(c["v_cache"])
    km, kp = cu(c["kv_metrics"]), cu(c["kv_position"], torch.int32)
    capi.execute_cache_moves(kc, vc, km, kp, moves, mcnt, offs)
    np.testing.assert_array_equal(cache_bits(kc, c["k_cache"]), c["out_k_cache"])
    np.testing.assert_array_equal(cache_bits(vc, c["v_cache"]), c["out_v_cache"])
    np.testing.assert_array_equal(km.cpu().numpy(), c["out_kv_metrics"])
    np.testing.assert_array_equal(kp.cpu().numpy(), c["out_kv_position"])


@pytest.fixture(params=[1, 2], ids=["v2_mma_sync", "v2_tcgen05"])
def v2_mode(request, capi):
    """partition kernel of paged attention v2 (kvc_set_attention_v2_mode); irrelevant for v1 / fused"""
    capi.set_attention_v2_mode(request.param)
    yield request.param
    capi.set_attention_v2_mode(0)


def run_attention(capi, c, version, record=True):
    S, Hq, H, D, B = (int(c[k]) for k in ("S", "Hq", "H", "D", "B"))
    Q = Hq // H
    bf = bool(c["is_bf16"])
    q, kc, vc = cu16(c["query"], bf), cu16(c["key_cache"], bf), cu16(c["value_cache"], bf)
    out = torch.zeros((S, Hq, D), dtype=q.dtype, device="cuda")
    NB = kc.shape[0]
    bt, cl = cu(c["block_tables"], torch.int32), cu(c["context_lens"], torch.int32)
    kp, lp, bl = (cu(c[k], torch.int32) for k in ("kv_position", "last_position", "kv_metric_buffer_len"))
    slopes = cu(c["alibi_slopes"]) if c["alibi_slopes"].size else None
    max_ctx = int(c["context_lens"].max())
    scale = float(c["scale"])
    if version == "fused":
        metrics = torch.zeros((NB, B), dtype=torch.float32, device="cuda")
        capi.paged_attention_fused_metrics(out, metrics, q, kc, vc, H, scale, bt, cl, kp, lp, bl, B,
                                           max_ctx, slopes)
        return out.float().cpu().numpy(), metrics.cpu().numpy()
    metric = torch.full((NB, B, Q), -1.0, dtype=torch.float32, device="cuda")
    if version == "v1":
        capi.paged_attention_v1(out, metric, q, kc, vc, H, scale, bt, cl, kp, lp, bl, B, max_ctx,
                                slopes, record)
    else:
        P = (max_ctx + 511) // 512
        es = torch.empty((S, Hq, P), dtype=torch.float32, device="cuda")
        ml = torch.empty_like(es)
        tmp = torch.empty((S, Hq, P, D), dtype=q.dtype, device="cuda")
        tmpm = torch.empty_like(metric)
        capi.paged_attention_v2(out, metric, es, ml, tmp, tmpm, q, kc, vc, H, scale, bt, cl, kp, lp,
                                bl, B, max_ctx, slopes, record)
    torch.cuda.synchronize()
    return out.float().cpu().numpy(), metric.cpu().numpy()


@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("name", sorted(ATTN))
def test_attention_golden(capi, name, version, v2_mode):
    if version != "v2" and v2_mode != 1:
        pytest.skip("v2 partition-kernel sweep")
    c = ATTN[name]
    out, metric = run_attention(capi, c, version)
    # reference tolerance: tests/kernels/test_kvcompress_attention.py:356-357
    np.testing.assert_allclose(out, c["out"], atol=1e-3, rtol=1e-5)
    untouched = c["kv_metric_out"] == -1.0
    assert ((metric == -1.0) == untouched).all()
    np.testing.assert_allclose(metric[~untouched], c["kv_metric_out"][~untouched], atol=2e-5, rtol=2e-3)


@pytest.mark.parametrize("name", sorted(ATTN))
def test_attention_fused_metrics_golden(capi, name):
    c = ATTN[name]
    out, metrics = run_attention(capi, c, "fused")
    np.testing.assert_allclose(out, c["out"], atol=1e-3, rtol=1e-5)
    gm = c["kv_metric_out"].copy()
    gm[gm == -1.0] = 0.0
    np.testing.assert_allclose(metrics, (gm ** 2).sum(-1), atol=1e-6, rtol=5e-3)


def test_attention_record_off_leaves_metrics_alone(capi):
    c = ATTN["attn0"]
    out, metric = run_attention(capi, c, "v1", record=False)
    np.testing.assert_allclose(out, c["out"], atol=1e-3, rtol=1e-5)
    assert (metric == -1.0).all()


@pytest.mark.parametrize("name", sorted(CACHE))
def test_k6_golden(capi, name):
    c = CACHE[name]
    kc, vc, m = cu16(c["key_cache"], False), cu16(c["value_cache"], False), cu(c["kv_metrics"])
    capi.reshape_and_cache(cu16(c["key"], False), cu16(c["value"], False), kc, vc, m,
                           cu(c["slot_mapping"], torch.int64), cu(c["kv_metric_head_bias"]))
    np.testing.assert_array_equal(bits16(kc), c["out_key_cache"])
    np.testing.assert_array_equal(bits16(vc), c["out_value_cache"])
    np.testing.assert_array_equal(m.cpu().numpy(), c["out_kv_metrics"])


# ---------------------------------------------------------------------------
# B. seeded random inputs vs the C oracle (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------
def k1_case(seed, S, L, H, B, max_blocks, prot, frac, shuffled_pos):
    from tests.golden.make_golden import build_eviction_inputs  # pure numpy builder
    rng = np.random.RandomState(seed)
    hang = rng.randint(1, B + 1, size=(S, L, H))
    ctx = rng.randint(0, max_blocks + 1, size=(S, L, H)) * B + hang
    inp = build_eviction_inputs(rng, S, L, H, B, ctx, interleave=True,
                                position_mode="shuffled" if shuffled_pos else "logical")
    nblk = (ctx + B - 1) // B
    inp["last_position"] = np.array([int(ctx[s].max()) - 1 for s in range(S)], np.int32)
    inp["protected"] = np.full(S, prot, np.int32)
    inp["budgets"] = np.array([int(nblk[s].sum() * frac) for s in range(S)], np.int32)
    inp["truncate"] = np.array(1)
    return inp


def oracle_k1(O, c, truncate, evenly=False, control=None):
    S, L, H, B = int(c["S"]), int(c["L"]), int(c["H"]), int(c["B"])
    N = int(c["total_kv"])
    ekv = np.zeros(N, np.int32)
    elog = np.full(N, -7, np.int32)
    cnt = np.full((S, L, H), 99, np.int32)
    rem = c["hanging_token_count"].astype(np.int32).copy()
    O.schedule_cache_evictions(ekv, elog, cnt, rem, c["evicted_kv_offsets"], c["sorted_indices"],
                               c["seq_block_offsets"], c["layer_by_block"], c["head_by_block"],
                               c["virtual_block_num_by_block"], c["budgets"], c["context_lens"],
                               c["hanging_token_count"], c["kv_position"], c["last_position"],
                               c["protected"], B, evenly, control, N, MAX_INT, truncate)
    return ekv, elog, cnt, rem


@pytest.mark.parametrize("truncate", [True, False])
@pytest.mark.parametrize("cfg", [
    # (seed, S, L, H, B, max_blocks, protected, budget fraction, shuffled positions)
    (1, 3, 2, 2, 4, 3, 0, 0.5, False),
    (2, 4, 4, 8, 16, 20, 32, 0.6, False),     # several 1024-entry chunks per sequence, cut mid-chunk
    (3, 2, 8, 8, 16, 40, 20, 0.9, True),
    (4, 5, 3, 4, 16, 30, 0, 1.5, False),      # budget never reached: every valid slot is evicted
    (5, 3, 2, 3, 2, 50, 3, 0.3, True),
    (6, 6, 4, 2, 1, 300, 5, 0.7, False),      # B=1
    (7, 1, 32, 8, 16, 12, 32, 0.8, False),    # Llama-3-8B head count (256 bins)
    (8, 2, 80, 8, 16, 3, 16, 0.5, False),     # Llama-3-70B head count (640 bins)
])
def test_k1_random_vs_oracle(capi, O, cfg, truncate):
    c = k1_case(*cfg)
    want = oracle_k1(O, c, truncate)
    got = run_k1(capi, c, truncate=truncate)
    for w, g, nm in zip(want, got, ("evicted_kv_indices", "evicted_logical_indices", "evicted_kv_count",
                                    "remaining_kv_count")):
        np.testing.assert_array_equal(g, w, err_msg=nm)


@pytest.mark.parametrize("control", [None, [1]])
def test_k1_evict_evenly_per_layer(capi, O, control):
    c = k1_case(11, 3, 4, 3, 1, 200, 4, 0.5, False)
    c["budgets"] = (c["budgets"] // 4) * 4
    ctl = None if control is None else np.array(control, np.int32)
    want = oracle_k1(O, c, True, evenly=True, control=ctl)
    got = run_k1(capi, c, truncate=True, evenly=True, control=ctl)
    for w, g in zip(want[:3], got[:3]):
        np.testing.assert_array_equal(g, w)


def test_unsupported_block_size_raises(capi):
    c = dict(k1_case(1, 1, 1, 1, 4, 2, 0, 0.5, False))
    c["B"] = np.array(8)
    with pytest.raises(capi.KvcError, match="Unsupported block size"):
        run_k1(capi, c)


def random_moves_state(seed, S, L, H, B, D, max_blocks, dtype=np.uint16):
    from tests.golden.make_golden import build_move_inputs
    return build_move_inputs(np.random.RandomState(seed), S, L, H, B, D, 16 // np.dtype(dtype).itemsize
                             if D >= 8 else 1, dtype, max_blocks)


@pytest.mark.parametrize("cfg", [(21, 4, 4, 8, 16, 128, 30), (22, 2, 2, 4, 16, 128, 200),
                                 (23, 3, 2, 2, 4, 4, 9), (24, 2, 3, 2, 1, 2, 40)])
def test_k4_k5_random_vs_oracle(capi, O, cfg, moves_mode):
    seed, S, L, H, B, D, mb = cfg
    c = random_moves_state(seed, S, L, H, B, D, mb)
    N = int(c["total_kv"])
    wm = np.zeros((N, 2), np.int32)
    wc = np.zeros((S, L, H), np.int32)
    O.schedule_cache_moves(wm, wc, c["evicted_logical_indices"], c["evicted_kv_count"],
                           c["evicted_kv_offsets"], c["block_tables"], c["context_lens_lsh"], B)
    moves = torch.zeros((N, 2), dtype=torch.int32, device="cuda")
    mcnt = torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
    offs = cu(c["evicted_kv_offsets"], torch.int32)
    capi.schedule_t1_cache_moves(moves, mcnt, cu(c["evicted_logical_indices"], torch.int32),
                                 cu(c["evicted_kv_count"], torch.int32), offs,
                                 cu(c["block_tables"], torch.int32), cu(c["context_lens_lsh"], torch.int32), B)
    np.testing.assert_array_equal(mcnt.cpu().numpy(), wc)
    np.testing.assert_array_equal(moves.cpu().numpy(), wm)
    kc, vc = c["k_cache"].copy(), c["v_cache"].copy()
    km, kp = c["kv_metrics"].copy(), c["kv_position"].copy()
    O.execute_cache_moves(kc, vc, km, kp, wm, wc, c["evicted_kv_offsets"])
    gk, gv = torch_cache(c["k_cache"]), torch_cache(c["v_cache"])
    gm, gp = cu(c["kv_metrics"]), cu(c["kv_position"], torch.int32)
    capi.execute_cache_moves(gk, gv, gm, gp, moves, mcnt, offs)
    np.testing.assert_array_equal(cache_bits(gk, c["k_cache"]), kc)
    np.testing.assert_array_equal(cache_bits(gv, c["v_cache"]), vc)
    np.testing.assert_array_equal(gm.cpu().numpy(), km)
    np.testing.assert_array_equal(gp.cpu().numpy(), kp)


@pytest.mark.parametrize("keep", [128, 368, 1488])
def test_k5_post_prefill_compaction_vs_oracle(capi, O, keep, moves_mode):
    """The post-prefill compaction of BASELINE config 5 in miniature: heads of 6000 tokens evicted down to `keep`
    (CR 64 / 16 / 4), protected window 32: runs whose 16 sources come from 16 different blocks (more than the
    staged kernel's ring holds: the refill path) as well as runs with several sources per block."""
    rng = np.random.RandomState(400 + keep)
    S, L, H, B, D, ctx_len = 1, 2, 6, 16, 128, 6000
    ctx = np.full((L, S, H), ctx_len, np.int32)
    ctx[0, 0, 1], ctx[1, 0, 2] = 5987, 5993                # ragged last blocks
    nblk = (ctx + B - 1) // B
    MB, NB = int(nblk.max()) + 1, int(nblk.sum()) + 5
    phys = rng.permutation(NB)
    bt = np.zeros((L, S, H, MB), np.int32)
    p = 0
    for l in range(L):
        for h in range(H):
            bt[l, 0, h, :nblk[l, 0, h]] = phys[p:p + nblk[l, 0, h]]
            p += nblk[l, 0, h]
    alloc = np.transpose(nblk, (1, 0, 2)).reshape(-1) * B
    offs = np.concatenate([[0], np.cumsum(alloc)[:-1]]).astype(np.int32).reshape(S, L, H)
    total = int(alloc.sum())
    logical = np.full(total, MAX_INT, np.int32)
    cnt = np.zeros((S, L, H), np.int32)
    for l in range(L):
        for h in range(H):
            n = int(ctx[l, 0, h])
            c = n - keep - int(rng.randint(0, 40))            # heads end up with slightly different lengths
            cnt[0, l, h] = c
            logical[offs[0, l, h]: offs[0, l, h] + c] = np.sort(rng.choice(n - 32, c, replace=False))
    wm, wc = np.zeros((total, 2), np.int32), np.zeros((S, L, H), np.int32)
    O.schedule_cache_moves(wm, wc, logical, cnt, offs, bt, ctx, B)
    moves = torch.zeros((total, 2), dtype=torch.int32, device="cuda")
    mcnt = torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
    capi.schedule_t1_cache_moves(moves, mcnt, cu(logical), cu(cnt), cu(offs), cu(bt), cu(ctx), B)
    np.testing.assert_array_equal(mcnt.cpu().numpy(), wc)
    np.testing.assert_array_equal(moves.cpu().numpy(), wm)
    kc = rng.randint(0, 65535, size=(NB, D // 8, B, 8)).astype(np.uint16)
    vc = rng.randint(0, 65535, size=(NB, D, B)).astype(np.uint16)
    km, kp = rng.rand(NB, B).astype(np.float32), rng.randint(0, 10000, size=(NB, B)).astype(np.int32)
    gk, gv, gm, gp = torch_cache(kc), torch_cache(vc), cu(km), cu(kp)
    capi.execute_cache_moves(gk, gv, gm, gp, moves, mcnt, cu(offs))
    O.execute_cache_moves(kc, vc, km, kp, wm, wc, offs)
    np.testing.assert_array_equal(cache_bits(gk, kc), kc)
    np.testing.assert_array_equal(cache_bits(gv, vc), vc)
    np.testing.assert_array_equal(gm.cpu().numpy(), km)
    np.testing.assert_array_equal(gp.cpu().numpy(), kp)


def test_k2_k3_random_vs_oracle(capi, O):
    rng = np.random.RandomState(31)
    S, L, H, B = 5, 4, 8, 16
    nblk = rng.randint(1, 9, size=(S, L, H))
    hanging = rng.randint(1, B + 1, size=(S, L, H)).astype(np.int32)
    alloc = (nblk * B).reshape(-1)
    offs = np.concatenate([[0], np.cumsum(alloc)[:-1]]).astype(np.int32).reshape(S, L, H)
    total = int(alloc.sum()) + 777  # slack: the last head's tail runs to max_evicted_kv
    cnt = np.minimum(rng.randint(0, 9 * B, size=(S, L, H)), nblk * B).astype(np.int32)
    ekv = rng.randint(0, 1000, size=total).astype(np.int32)
    elog = rng.randint(0, 1000, size=total).astype(np.int32)
    w = [ekv.copy(), elog.copy(), cnt.copy()]
    O.truncate_cache_evictions(w[0], w[1], w[2], offs, hanging, B, total, MAX_INT)
    g = [cu(ekv), cu(elog), cu(cnt)]
    capi.truncate_cache_evictions(g[0], g[1], g[2], cu(offs), cu(hanging), B, total, MAX_INT)
    for a, b in zip(w, g):
        np.testing.assert_array_equal(b.cpu().numpy(), a)
    # K3 on V2-format rows
    logical = np.full(total, MAX_INT, np.int32)
    for i, (o, n) in enumerate(zip(offs.reshape(-1), nblk.reshape(-1))):
        e = rng.randint(0, n + 1)
        logical[o:o + e * B] = rng.randint(0, n * B, size=e * B)
    wl, wc = logical.copy(), np.zeros((S, L, H), np.int32)
    O.count_block_evictions(wc, wl, offs, hanging, B, MAX_INT)
    gl, gc = cu(logical), torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
    capi.count_block_evictions(gc, gl, cu(offs), cu(hanging), B, MAX_INT)
    np.testing.assert_array_equal(gc.cpu().numpy(), wc)
    np.testing.assert_array_equal(gl.cpu().numpy(), wl)


def attn_case(seed, S, Hq, H, max_ctx, bf16, alibi, buf, min_ctx=1):
    rng = np.random.RandomState(seed)
    g = torch.Generator().manual_seed(seed)
    D, B, x = 128, 16, 8
    ctx = rng.randint(min_ctx, max_ctx + 1, size=(S, H)).astype(np.int32)
    ctx[0, 0] = max_ctx
    nblk = (ctx + B - 1) // B
    MB = int(nblk.max()) + 1
    NB = int(nblk.sum()) + 3
    phys = rng.permutation(NB)
    bt = np.zeros((S, H, MB), np.int32)
    p = 0
    for s in range(S):
        for h in range(H):
            bt[s, h, :nblk[s, h]] = phys[p:p + nblk[s, h]]
            p += nblk[s, h]
    dt = torch.bfloat16 if bf16 else torch.float16
    scale = D ** -0.5
    mk = lambda shape: (torch.rand(shape, generator=g) * 2 - 1).mul(scale * 4).to(dt)
    q, kc, vc = mk((S, Hq, D)), mk((NB, D // x, B, x)), mk((NB, D, B))
    v16 = lambda t: t.view(torch.int16).numpy().view(np.uint16)
    return dict(S=S, Hq=Hq, H=H, D=D, B=B, is_bf16=int(bf16), scale=np.float32(scale), query=v16(q),
                key_cache=v16(kc), value_cache=v16(vc), block_tables=bt, context_lens=ctx,
                kv_position=rng.randint(0, 300, size=(NB, B)).astype(np.int32),
                last_position=np.full(S, 299, np.int32), kv_metric_buffer_len=np.full(S, buf, np.int32),
                alibi_slopes=(rng.rand(Hq).astype(np.float32) if alibi else np.zeros(0, np.float32)))


@pytest.mark.parametrize("version", ["v1", "v2", "fused"])
@pytest.mark.parametrize("cfg", [
    (41, 5, 32, 8, 160, False, False, 0),     # Llama-3-8B GQA group (Q=4), compressed-length contexts
    (42, 3, 64, 8, 200, True, False, 10),     # Llama-3-70B group (Q=8), bf16
    (43, 2, 8, 8, 1500, False, True, 0),      # MHA, several partitions, alibi
    (44, 2, 24, 2, 700, False, False, 50),    # Q=12 (two n-tiles)
    (45, 4, 4, 4, 17, True, False, 0),        # tiny ragged contexts
    (46, 3, 32, 8, 400, False, False, 5),     # streaming kernel, logits beyond the smem cap (fused: global scratch)
    (47, 3, 32, 8, 2100, False, False, 0),    # long contexts: 5 partitions of 4 MMA tiles (v2: tcgen05 partition kernel)
    (48, 2, 64, 8, 1300, True, True, 16),     # Llama-3-70B group, bf16, alibi, metric buffer, long contexts
])
def test_attention_random_vs_oracle(capi, O, cfg, version, v2_mode):
    if version != "v2" and v2_mode != 1:
        pytest.skip("v2 partition-kernel sweep")
    c = attn_case(*cfg)
    S, Hq, H, D, B = (int(c[k]) for k in ("S", "Hq", "H", "D", "B"))
    Q = Hq // H
    want = np.zeros((S, Hq, D), np.float32)
    wmet = np.full((c["key_cache"].shape[0], B, Q), -1.0, np.float32)
    slopes = c["alibi_slopes"] if c["alibi_slopes"].size else None
    O.paged_attention_kvc(want, wmet, c["query"], c["key_cache"], c["value_cache"], H, float(c["scale"]),
                          c["block_tables"], c["context_lens"], c["kv_position"], c["last_position"],
                          c["kv_metric_buffer_len"], slopes, bool(c["is_bf16"]), True)
    out, metric = run_attention(capi, c, version)
    # bf16 output rounding alone is 2^-9 relative; values here are O(0.1)
    np.testing.assert_allclose(out, want, atol=2e-3 if c["is_bf16"] else 1e-3, rtol=1e-2)
    if version == "fused":
        wm = wmet.copy()
        wm[wm == -1.0] = 0
        np.testing.assert_allclose(metric, (wm ** 2).sum(-1), atol=1e-6, rtol=1e-2)
    else:
        untouched = wmet == -1.0
        assert ((metric == -1.0) == untouched).all()
        np.testing.assert_allclose(metric[~untouched], wmet[~untouched], atol=2e-5, rtol=5e-3)


def test_k6_random_vs_oracle(capi, O):
    rng = np.random.RandomState(51)
    T, H, D, B, x, NB = 37, 8, 128, 16, 8, 64
    key = rng.randint(0, 65535, size=(T, H, D)).astype(np.uint16)
    val = rng.randint(0, 65535, size=(T, H, D)).astype(np.uint16)
    kc = rng.randint(0, 65535, size=(NB, D // x, B, x)).astype(np.uint16)
    vc = rng.randint(0, 65535, size=(NB, D, B)).astype(np.uint16)
    m = rng.rand(NB, B).astype(np.float32)
    bias = rng.rand(H).astype(np.float32)
    slots = rng.choice(NB * B, T * H, replace=False).astype(np.int64).reshape(T, H)
    slots[3, :] = -1
    wk, wv, wm = kc.copy(), vc.copy(), m.copy()
    O.reshape_and_cache_kvc(key, val, wk, wv, wm, slots, bias)
    gk, gv, gm = cu16(kc, False), cu16(vc, False), cu(m)
    capi.reshape_and_cache(cu16(key, False), cu16(val, False), gk, gv, gm, cu(slots, torch.int64), cu(bias))
    np.testing.assert_array_equal(bits16(gk), wk)
    np.testing.assert_array_equal(bits16(gv), wv)
    np.testing.assert_array_equal(gm.cpu().numpy(), wm)


# ---------------------------------------------------------------------------
# C. kernel-vs-kernel against the reference's own CUDA kernels (oracle/_ref)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(2, 4, 4, 8, 16, 20, 32, 0.6, False), (7, 1, 32, 8, 16, 12, 32, 0.8, False),
                                 (3, 2, 8, 8, 16, 40, 20, 0.9, True)])
def test_k1_vs_reference_cuda(capi, ref_ops, cfg):
    c = k1_case(*cfg)
    S, L, H, B = int(c["S"]), int(c["L"]), int(c["H"]), int(c["B"])
    N = int(c["total_kv"])
    i = lambda k: cu(c[k], torch.int32)
    ekv = torch.zeros(N, dtype=torch.int32, device="cuda")
    elog = torch.full((N,), -7, dtype=torch.int32, device="cuda")
    cnt = torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
    rem = i("hanging_token_count")
    ref_ops.schedule_cache_evictions(ekv, elog, cnt, rem, i("evicted_kv_offsets"), i("sorted_indices"),
                                     i("seq_block_offsets"), i("layer_by_block"), i("head_by_block"),
                                     i("virtual_block_num_by_block"), i("budgets"), i("context_lens"),
                                     i("hanging_token_count"), i("kv_position"), i("last_position"),
                                     i("protected"), B, False, None, N, MAX_INT, True)
    torch.cuda.synchronize()
    got = run_k1(capi, c, truncate=True)
    np.testing.assert_array_equal(got[0], ekv.cpu().numpy())
    np.testing.assert_array_equal(got[1], elog.cpu().numpy())
    np.testing.assert_array_equal(got[2], cnt.cpu().numpy())
    np.testing.assert_array_equal(got[3], rem.cpu().numpy())


def test_k4_k5_vs_reference_cuda(capi, ref_ops, moves_mode):
    c = random_moves_state(61, 4, 4, 8, 16, 128, 40)
    S, L, H = c["evicted_kv_count"].shape
    N, B = int(c["total_kv"]), 16
    args = [cu(c["evicted_logical_indices"], torch.int32), cu(c["evicted_kv_count"], torch.int32),
            cu(c["evicted_kv_offsets"], torch.int32), cu(c["block_tables"], torch.int32),
            cu(c["context_lens_lsh"], torch.int32)]
    rm, rc = torch.zeros((N, 2), dtype=torch.int32, device="cuda"), torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
    ref_ops.schedule_t1_cache_moves(rm, rc, *args, B)
    gm, gc = torch.zeros_like(rm), torch.zeros_like(rc)
    capi.schedule_t1_cache_moves(gm, gc, *args, B)
    assert torch.equal(gm, rm) and torch.equal(gc, rc)
    state = lambda: (torch_cache(c["k_cache"]), torch_cache(c["v_cache"]), cu(c["kv_metrics"]),
                     cu(c["kv_position"], torch.int32))
    r, g = state(), state()
    ref_ops.execute_cache_moves(*r, rm, rc, args[2], 1, 16)
    capi.execute_cache_moves(*g, gm, gc, args[2])
    for a, b in zip(r, g):
        assert torch.equal(a.view(torch.int16) if a.dtype == torch.float16 else a,
                           b.view(torch.int16) if b.dtype == torch.float16 else b)


def test_k2_k3_vs_reference_cuda(capi, ref_ops):
    """truncate_cache_evictions (K2) and count_block_evictions (K3) kernel-vs-kernel against the reference's
    own CUDA kernels (csrc/kvcompress_eviction_kernels.cu:158-221) on identical device tensors."""
    rng = np.random.RandomState(131)
    for (S, L, H, B) in ((5, 4, 8, 16), (3, 32, 8, 16), (2, 2, 4, 4)):
        nblk = rng.randint(1, 9, size=(S, L, H))
        hanging = rng.randint(1, B + 1, size=(S, L, H)).astype(np.int32)
        alloc = (nblk * B).reshape(-1)
        offs = np.concatenate([[0], np.cumsum(alloc)[:-1]]).astype(np.int32).reshape(S, L, H)
        total = int(alloc.sum()) + 333
        cnt = np.minimum(rng.randint(0, 9 * B, size=(S, L, H)), nblk * B).astype(np.int32)
        ekv = rng.randint(0, 1000, size=total).astype(np.int32)
        elog = rng.randint(0, 1000, size=total).astype(np.int32)
        r = [cu(ekv), cu(elog), cu(cnt)]
        g = [cu(ekv), cu(elog), cu(cnt)]
        ref_ops.truncate_cache_evictions(r[0], r[1], r[2], cu(offs), cu(hanging), B, total, MAX_INT)
        capi.truncate_cache_evictions(g[0], g[1], g[2], cu(offs), cu(hanging), B, total, MAX_INT)
        torch.cuda.synchronize()
        for a, b in zip(r, g):
            assert torch.equal(a, b)
        logical = np.full(total, MAX_INT, np.int32)
        for o, n in zip(offs.reshape(-1), nblk.reshape(-1)):
            e = rng.randint(0, n + 1)
            logical[o:o + e * B] = rng.randint(0, n * B, size=e * B)
        rl, rc = cu(logical), torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
        gl, gc = cu(logical), torch.zeros((S, L, H), dtype=torch.int32, device="cuda")
        ref_ops.count_block_evictions(rc, rl, cu(offs), cu(hanging), B, MAX_INT)
        capi.count_block_evictions(gc, gl, cu(offs), cu(hanging), B, MAX_INT)
        torch.cuda.synchronize()
        assert torch.equal(rc, gc) and torch.equal(rl, gl)


@pytest.mark.parametrize("lens", [[53, 16, 7, 130, 32], [300, 53, 16, 7, 130, 32, 411, 1]])
@pytest.mark.parametrize("bf16", [False, True])
def test_k6_prefill_blocks_vs_reference_cuda(capi, ref_ops, O, bf16, lens):
    """K6 on a prefill batch: every head's prompt tokens fill randomly placed blocks in order, so most (16 tokens,
    head) groups are whole blocks (the warp of the block's first token writes all 4 KB) next to
    partial tail blocks, prompts that do not start on a multiple of 16 in the batch, a padding token and a
    re-directed slot inside what would otherwise be whole blocks.  Against the reference's CUDA kernel and the C
    oracle, on strided views of one QKV tensor, in both launch geometries (batches of >= 512 tokens take one warp per
    16-token window)."""
    rng = np.random.RandomState(99)
    H, D, B, x = 8, 128, 16, 8
    T = sum(lens)   # 238 tokens: one warp per (token, head); 950 tokens: one warp per (16-token window, head)
    dt = torch.bfloat16 if bf16 else torch.float16
    qkv = (torch.rand((T, (4 * H + 2 * H) * D), device="cuda") - 0.5).to(dt)
    key = qkv[:, 4 * H * D: 5 * H * D].view(T, H, D)
    val = qkv[:, 5 * H * D:].view(T, H, D)
    nblk = [(n + B - 1) // B for n in lens]
    NB = H * sum(nblk) + 5
    perm = rng.permutation(NB)
    slots = np.empty((T, H), np.int64)
    t0 = b0 = 0
    for n, nb in zip(lens, nblk):
        blocks = perm[b0:b0 + H * nb].reshape(H, nb)
        tok = np.arange(n)
        slots[t0:t0 + n] = (blocks[:, tok // B] * B + tok % B).T
        t0 += n
        b0 += H * nb
    slots[60, 2] = -1                       # padding token inside a whole block of head 2
    slots[100, 5] = perm[-1] * B + 3        # one token of head 5 goes elsewhere: its block is not a whole one
    slots[20, :] = -1
    kc = (torch.rand((NB, D // x, B, x), device="cuda") - 0.5).to(dt)
    vc = (torch.rand((NB, D, B), device="cuda") - 0.5).to(dt)
    m = torch.rand((NB, B), device="cuda")
    bias = torch.rand(H, device="cuda")
    sl = torch.from_numpy(slots).cuda()
    r = (kc.clone(), vc.clone(), m.clone())
    g = (kc.clone(), vc.clone(), m.clone())
    ref_ops.kvcompress_reshape_and_cache(key, val, r[0], r[1], r[2], sl.view(-1), bias, "auto", 1.0, 1.0)
    capi.reshape_and_cache(key, val, g[0], g[1], g[2], sl, bias)
    torch.cuda.synchronize()
    for a_, b_ in zip(r, g):
        assert torch.equal(a_.view(torch.int16) if a_.dtype != torch.float32 else a_,
                           b_.view(torch.int16) if b_.dtype != torch.float32 else b_)
    bits = lambda t_: t_.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
    wk, wv, wm = bits(kc), bits(vc), m.cpu().numpy().copy()
    O.reshape_and_cache_kvc(bits(key), bits(val), wk, wv, wm, slots, bias.cpu().numpy())
    np.testing.assert_array_equal(bits(g[0]), wk)
    np.testing.assert_array_equal(bits(g[1]), wv)
    np.testing.assert_array_equal(g[2].cpu().numpy(), wm)


@pytest.mark.parametrize("bf16", [False, True])
def test_k6_vs_reference_cuda(capi, ref_ops, bf16):
    """kvcompress_reshape_and_cache (K6) kernel-vs-kernel against the reference's CUDA kernel
    (csrc/kvcompress_cache_kernels.cu:27-89), strided key / value views of one QKV tensor as the model passes them."""
    rng = np.random.RandomState(151)
    T, H, D, B, x, NB = 41, 8, 128, 16, 8, 80
    dt = torch.bfloat16 if bf16 else torch.float16
    qkv = (torch.rand((T, (4 * H + 2 * H) * D), device="cuda") - 0.5).to(dt)
    key = qkv[:, 4 * H * D: 5 * H * D].view(T, H, D)
    val = qkv[:, 5 * H * D:].view(T, H, D)
    kc = (torch.rand((NB, D // x, B, x), device="cuda") - 0.5).to(dt)
    vc = (torch.rand((NB, D, B), device="cuda") - 0.5).to(dt)
    m = torch.rand((NB, B), device="cuda")
    bias = torch.rand(H, device="cuda")
    slots = torch.from_numpy(rng.choice(NB * B, T * H, replace=False).astype(np.int64).reshape(T, H)).cuda()
    slots[5, :] = -1
    slots[7, 3] = -1
    r = (kc.clone(), vc.clone(), m.clone())
    g = (kc.clone(), vc.clone(), m.clone())
    ref_ops.kvcompress_reshape_and_cache(key, val, r[0], r[1], r[2], slots.view(-1), bias, "auto", 1.0, 1.0)
    capi.reshape_and_cache(key, val, g[0], g[1], g[2], slots, bias)
    torch.cuda.synchronize()
    assert torch.equal(r[0].view(torch.int16), g[0].view(torch.int16))
    assert torch.equal(r[1].view(torch.int16), g[1].view(torch.int16))
    assert torch.equal(r[2], g[2])


@pytest.mark.parametrize("version", ["v1", "v2"])
@pytest.mark.parametrize("cfg", [(41, 5, 32, 8, 160, False, False, 0), (42, 3, 64, 8, 200, True, False, 0),
                                 (43, 2, 8, 8, 1500, False, True, 0), (47, 3, 32, 8, 2100, False, False, 0),
                                 (48, 2, 64, 8, 1300, True, False, 0)])   # buffer 0: the reference's v2 reduce writes
                                 # unrecorded garbage into buffered slots (:563-569), see DESIGN.md section 4
def test_attention_vs_reference_cuda(capi, ref_ops, cfg, version, v2_mode):
    if version != "v2" and v2_mode != 1:
        pytest.skip("v2 partition-kernel sweep")
    c = attn_case(*cfg)
    S, Hq, H, D, B = (int(c[k]) for k in ("S", "Hq", "H", "D", "B"))
    Q = Hq // H
    bf = bool(c["is_bf16"])
    q, kc, vc = cu16(c["query"], bf), cu16(c["key_cache"], bf), cu16(c["value_cache"], bf)
    NB = kc.shape[0]
    bt, cl = cu(c["block_tables"], torch.int32), cu(c["context_lens"], torch.int32)
    kp, lp, bl = (cu(c[k], torch.int32) for k in ("kv_position", "last_position", "kv_metric_buffer_len"))
    slopes = cu(c["alibi_slopes"]) if c["alibi_slopes"].size else None
    max_ctx = int(c["context_lens"].max())
    rout = torch.zeros((S, Hq, D), dtype=q.dtype, device="cuda")
    rmet = torch.full((NB, B, Q), -1.0, dtype=torch.float32, device="cuda")
    if version == "v1":
        ref_ops.kvcompress_paged_attention_v1(rout, rmet, q, kc, vc, H, float(c["scale"]), bt, cl, kp, lp, bl,
                                              B, max_ctx, slopes, "auto", 1.0, 1.0, True)
    else:
        P = (max_ctx + 511) // 512
        es = torch.empty((S, Hq, P), dtype=torch.float32, device="cuda")
        ml = torch.empty_like(es)
        tmp = torch.empty((S, Hq, P, D), dtype=q.dtype, device="cuda")
        tmpm = torch.zeros_like(rmet)
        ref_ops.kvcompress_paged_attention_v2(rout, rmet, es, ml, tmp, tmpm, q, kc, vc, H, float(c["scale"]),
                                              bt, cl, kp, lp, bl, B, max_ctx, slopes, "auto", 1.0, 1.0, True)
    torch.cuda.synchronize()
    out, metric = run_attention(capi, c, version)
    np.testing.assert_allclose(out, rout.float().cpu().numpy(), atol=1e-3, rtol=1e-2)
    rm = rmet.cpu().numpy()
    assert ((metric == -1.0) == (rm == -1.0)).all()
    np.testing.assert_allclose(metric, rm, atol=2e-5, rtol=5e-3)


# ---------------------------------------------------------------------------
# D. fused selection (additive op) == sort -> K1+K2 -> re-sort chain of the reference
# ---------------------------------------------------------------------------
@pytest.fixture(params=[1, 2], ids=["three_kernels", "fused_per_sequence"])
def select_mode(request, capi):
    """kvc_set_select_mode: 1 = rank / threshold / emit launches, 2 = the one-CTA-per-sequence kernel with the
    pruned sort wherever the shape allows it (block size 16, <= 512 slots per head)."""
    capi.set_select_mode(request.param)
    yield request.param
    capi.set_select_mode(0)


def selection_case(seed, S, L, H, B, max_blocks, prot, frac):
    """State in engine layout (block tables + paged metrics) plus the K1 view of it."""
    from vllm_kvcompress_b200 import workload as W
    g = torch.Generator().manual_seed(seed)
    ctx = torch.randint(1, max_blocks * B + 1, (L, S, H), generator=g, dtype=torch.int32)
    st = W.make_state(S, L, H, B, 16, ctx, dtype=torch.float16, device="cpu", seed=seed, fill_cache=False)
    # positions: a shuffled subset of [0, 2*ctx) per head, so the protected window bites unevenly
    pos = torch.full((st.NB * B,), -1, dtype=torch.int32)
    for l in range(L):
        for s in range(S):
            for h in range(H):
                n = int(ctx[l, s, h])
                slots = (st.block_tables[l, s, h, torch.arange(n) // B].long() * B + torch.arange(n) % B)
                pos[slots] = torch.randperm(2 * n, generator=g)[:n].to(torch.int32)
    st.kv_position = pos.view(st.NB, B)
    last = torch.tensor([int(2 * ctx[:, s].max()) for s in range(S)], dtype=torch.int32)
    protected = torch.full((S,), prot, dtype=torch.int32)
    nblk = (ctx + B - 1) // B
    budgets = (nblk.sum(dim=(0, 2)).float() * frac).to(torch.int32)
    return st, ctx, last, protected, budgets


@pytest.mark.parametrize("cfg", [
    (71, 3, 2, 2, 4, 5, 3, 0.5),
    (72, 4, 4, 8, 16, 10, 16, 0.4),      # <= 256 slots per head: warp-per-head path
    (79, 4, 4, 8, 16, 14, 16, 0.4),      # 193..256 slots per head
    (77, 3, 4, 8, 16, 28, 32, 0.6),      # 257..512 slots per head: warp path with 16 keys per lane
    (73, 2, 4, 8, 16, 40, 32, 0.7),      # > 512 slots per head: CTA-per-head sort path
    (78, 2, 1, 3, 16, 380, 32, 0.95),    # post-prefill-like: ~6000 slots per head, CR ~ 20
    (74, 3, 3, 4, 16, 9, 0, 2.0),        # budget larger than what exists
    (75, 2, 2, 2, 1, 100, 7, 0.5),
    (76, 2, 32, 8, 16, 9, 32, 0.12),     # Llama-3-8B head count, steady-state-like budget
])
def test_select_evictions_equals_reference_chain(capi, O, cfg, select_mode):
    from vllm_kvcompress_b200 import workload as W
    seed, S, L, H, B = cfg[:5]
    st, ctx, last, protected, budgets = selection_case(*cfg)
    hanging = W.hanging_token_counts(ctx, B).transpose(0, 1).contiguous()
    offs = W.evicted_kv_offsets(ctx, B)
    N = int(((ctx + B - 1) // B).sum()) * B
    # reference chain on the CPU: sort_seq_metrics -> K1+K2 (oracle) -> host re-sort -> K3-style counts
    sm = W.sort_seq_metrics(st)
    n = lambda t: t.numpy()
    ekv, elog = np.zeros(N, np.int32), np.zeros(N, np.int32)
    cnt = np.zeros((S, L, H), np.int32)
    O.schedule_cache_evictions(ekv, elog, cnt, n(hanging).copy(), n(offs), n(sm.sorted_indices),
                               n(sm.seq_block_offsets), n(sm.layer_by_block), n(sm.head_by_block),
                               n(sm.logical_block_num_by_block), n(budgets), n(ctx.transpose(0, 1).contiguous()),
                               n(hanging), n(sm.token_positions), n(last), n(protected), B, False, None, N,
                               MAX_INT, True)
    want_logical = O.sort_evictions_by_head(ekv, elog)
    hn = n(hanging)
    want_blocks = np.where(cnt >= hn, (cnt - hn) // B + 1, 0)
    # fused op on the GPU
    out_logical = torch.full((N,), -5, dtype=torch.int32, device="cuda")
    out_cnt = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
    out_blk = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
    capi.select_evictions(out_logical, out_cnt, out_blk, st.kv_metrics.cuda(), st.kv_position.cuda(),
                          st.block_tables.cuda(), ctx.cuda(), hanging.cuda(), offs.cuda(), budgets.cuda(),
                          last.cuda(), protected.cuda(), B, MAX_INT)
    np.testing.assert_array_equal(out_cnt.cpu().numpy(), cnt)
    np.testing.assert_array_equal(out_blk.cpu().numpy(), want_blocks)
    np.testing.assert_array_equal(out_logical.cpu().numpy(), want_logical)


SELV2 = load_cases("selection_v2.npz")


@pytest.mark.parametrize("quantize", [0, 8, 3])
@pytest.mark.parametrize("cfg", [(91, 3, 2, 4, 16, 120, 32), (92, 2, 3, 2, 16, 400, 16), (93, 4, 1, 3, 4, 300, 5)])
def test_select_evictions_long_context_radix_equals_sort(capi, cfg, quantize):
    """Long contexts: the sort-free radix path (default: one cooperative launch; mode 3: seven launches) against the
    CTA-per-head sort path (select mode 1) on the same inputs, bit for bit, over budgets of 0 / small / typical / more than exists, with metrics quantised to a
    handful of values so that the tie rules (completion keys by (key, head, block), slots by (key, logical
    index)) decide almost every grant."""
    from vllm_kvcompress_b200 import workload as W
    seed, S, L, H, B, max_blocks, prot = cfg
    st, ctx, last, protected, _ = selection_case(seed, S, L, H, B, max_blocks, prot, 0.5)
    if quantize:
        st.kv_metrics = (st.kv_metrics * quantize).floor() / quantize
    hanging = W.hanging_token_counts(ctx, B).transpose(0, 1).contiguous()
    offs = W.evicted_kv_offsets(ctx, B)
    nblk = (ctx + B - 1) // B
    N = int(nblk.sum()) * B
    total = nblk.sum(dim=(0, 2))
    for frac in (0.0, 0.02, 0.6, 0.97, 1.5):
        budgets = (total.float() * frac).to(torch.int32)
        budgets[0] = int(total[0] * (1.0 - frac)) if frac <= 1 else budgets[0]   # sequences with different budgets
        outs = []
        for mode in (1, 3, 0):
            capi.set_select_mode(mode)
            logical = torch.full((N,), -5, dtype=torch.int32, device="cuda")
            kvc = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
            blk = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
            capi.select_evictions(logical, kvc, blk, cu(st.kv_metrics), cu(st.kv_position), cu(st.block_tables),
                                  cu(ctx), cu(hanging), cu(offs), cu(budgets), cu(last), cu(protected), B, MAX_INT)
            torch.cuda.synchronize()
            outs.append((logical.cpu(), kvc.cpu(), blk.cpu()))
        capi.set_select_mode(0)
        for other in outs[1:]:
            for a, b in zip(outs[0], other):
                assert torch.equal(a, b), (cfg, quantize, frac)
        assert int(outs[0][2].sum()) == int(torch.minimum(budgets, outs[0][2].sum(dim=(1, 2))).sum())


@pytest.mark.parametrize("name", sorted(SELV2))
def test_select_evictions_v2_golden(capi, name, select_mode):
    """bit-exact against CompressionMetrics.schedule_evictions (metrics.py:441-847) run on CPU,
    incl. head bias, use_average and attention sinks."""
    c = SELV2[name]
    one = lambda v: torch.tensor([int(v)], dtype=torch.int32, device="cuda")
    N = c["out_evicted_logical_indices"].shape[0]
    out_logical = torch.full((N,), -5, dtype=torch.int32, device="cuda")
    out_cnt = torch.full(c["out_evicted_kv_count"].shape, -5, dtype=torch.int32, device="cuda")
    out_blk = torch.full_like(out_cnt, -5)
    has_bias = bool(int(c["has_bias"]))
    capi.select_evictions(out_logical, out_cnt, out_blk, cu(c["kv_metrics"]), cu(c["kv_position"]),
                          cu(c["block_tables"]), cu(c["context_lens"]), cu(c["hanging_token_count"]),
                          cu(c["evicted_kv_offsets"]), one(c["budget"]), one(c["last_position"]), one(c["protected"]),
                          int(c["B"]), MAX_INT, cu(c["head_bias"]) if has_bias else None,
                          cu(c["position_bins"]) if has_bias else None, float(c["bias_weight"]),
                          bool(c["use_average"]), int(c["num_sinks"]))
    np.testing.assert_array_equal(out_cnt.cpu().numpy(), c["out_evicted_kv_count"])
    np.testing.assert_array_equal(out_blk.cpu().numpy(), c["out_evicted_block_count"])
    np.testing.assert_array_equal(out_logical.cpu().numpy(), c["out_evicted_logical_indices"])


@pytest.mark.parametrize("cfg", [
    # seed, S, L, H, B, max_blocks, prot, frac, bins, weight, use_average, sinks
    (91, 3, 4, 8, 16, 10, 16, 0.5, [0, 50, 120], 0.7, False, 3),     # warp path, 8 keys per lane
    (92, 2, 4, 8, 16, 28, 32, 0.6, None, 0.0, True, 0),             # warp path, 16 keys per lane
    (93, 2, 2, 4, 16, 70, 32, 0.7, [0, 300], 1.5, True, 5),         # CTA sort path
    (94, 2, 1, 2, 16, 400, 32, 0.9, None, 0.0, False, 0),           # CTA sort path, 6400 slots per head
])
def test_select_evictions_options_vs_oracle(capi, O, cfg, select_mode):
    from vllm_kvcompress_b200 import workload as W
    seed, S, L, H, B, max_blocks, prot, frac, bins, weight, use_avg, sinks = cfg
    st, ctx, last, protected, budgets = selection_case(seed, S, L, H, B, max_blocks, prot, frac)
    hanging = W.hanging_token_counts(ctx, B).transpose(0, 1).contiguous()
    offs = W.evicted_kv_offsets(ctx, B)
    g = torch.Generator().manual_seed(seed)
    bias = torch.rand((L, H, len(bins)), generator=g) if bins is not None else None
    pbins = torch.tensor(bins, dtype=torch.int32) if bins is not None else None
    n = lambda t: None if t is None else t.numpy()
    want_logical, want_cnt, want_blk = O.select_evictions(
        n(st.kv_metrics), n(st.kv_position), n(st.block_tables), n(ctx), n(hanging), n(offs), n(budgets), n(last),
        n(protected), B, MAX_INT, n(bias), n(pbins), weight, use_avg, sinks)
    N = want_logical.shape[0]
    out_logical = torch.full((N,), -5, dtype=torch.int32, device="cuda")
    out_cnt = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
    out_blk = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
    dev = lambda t: None if t is None else t.cuda()
    capi.select_evictions(out_logical, out_cnt, out_blk, st.kv_metrics.cuda(), st.kv_position.cuda(),
                          st.block_tables.cuda(), ctx.cuda(), hanging.cuda(), offs.cuda(), budgets.cuda(),
                          last.cuda(), protected.cuda(), B, MAX_INT, dev(bias), dev(pbins), weight, use_avg, sinks)
    np.testing.assert_array_equal(out_cnt.cpu().numpy(), want_cnt)
    np.testing.assert_array_equal(out_blk.cpu().numpy(), want_blk)
    np.testing.assert_array_equal(out_logical.cpu().numpy(), want_logical)


@pytest.mark.parametrize("quantize", [False, True])
@pytest.mark.parametrize("cfg", [
    # S, L, H, base ctx, budget per head, protected
    (100, 4, 8, 128, 1.0, 32),     # steady state of continual compression: every head one token past a block
    (97, 2, 8, 368, 0.9, 32),      # 16 keys per lane (70B-like contexts)
    (96, 3, 4, 64, 2.5, 8),        # budget beyond one block per head
])
def test_select_evictions_fused_steady_state(capi, O, cfg, quantize):
    """Decode-time shape (many sequences, heads mostly at 16k+1 tokens, budget ~ one block per head): the
    per-sequence kernel with its pruned sort against the three-kernel path and the numpy oracle, bit-exact,
    also with heavily tied metrics (quantize: 8 distinct values) where the tie rules decide."""
    from vllm_kvcompress_b200 import workload as W
    S, L, H, base, per_head, prot = cfg
    B = 16
    g = torch.Generator().manual_seed(S + L)
    ctx = torch.full((L, S, H), base + 1, dtype=torch.int32)
    r = torch.rand((L, S, H), generator=g)
    ctx[r < 0.10] += torch.randint(1, 15, (int((r < 0.10).sum()),), generator=g, dtype=torch.int32)
    ctx[r > 0.95] -= 17
    st = W.make_state(S, L, H, B, 16, ctx, dtype=torch.float16, device="cpu", seed=7, fill_cache=False)
    if quantize:
        st.kv_metrics = (st.kv_metrics * 8).floor() / 8
    heads = L * S * H
    n = int(ctx.max())
    pos = torch.full((st.NB * B,), -1, dtype=torch.int32)
    bt = st.block_tables.view(heads, -1)
    for i in range(heads):
        c = int(ctx.view(-1)[i])
        slots = bt[i, torch.arange(c) // B].long() * B + torch.arange(c) % B
        pos[slots] = torch.cat([torch.randperm(5000, generator=g)[: c - prot // 2].sort().values,
                                5000 + torch.arange(prot // 2)]).to(torch.int32)
    st.kv_position = pos.view(st.NB, B)
    last = torch.full((S,), 5000 + prot // 2 - 1, dtype=torch.int32)
    protected = torch.full((S,), prot, dtype=torch.int32)
    budgets = torch.full((S,), int(per_head * L * H), dtype=torch.int32)
    budgets[::7] = 0
    budgets[3::11] = 5
    hanging = W.hanging_token_counts(ctx, B).transpose(0, 1).contiguous()
    offs = W.evicted_kv_offsets(ctx, B)
    N = int(((ctx + B - 1) // B).sum()) * B
    outs = {}
    for mode in (1, 2):
        capi.set_select_mode(mode)
        out_logical = torch.full((N,), -5, dtype=torch.int32, device="cuda")
        out_cnt = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
        out_blk = torch.full((S, L, H), -5, dtype=torch.int32, device="cuda")
        capi.select_evictions(out_logical, out_cnt, out_blk, st.kv_metrics.cuda(), st.kv_position.cuda(),
                              st.block_tables.cuda(), ctx.cuda(), hanging.cuda(), offs.cuda(), budgets.cuda(),
                              last.cuda(), protected.cuda(), B, MAX_INT)
        outs[mode] = [t.cpu().numpy() for t in (out_logical, out_cnt, out_blk)]
    capi.set_select_mode(0)
    for a, b in zip(outs[1], outs[2]):
        np.testing.assert_array_equal(a, b)
    nn = lambda t: t.numpy()
    want_logical, want_cnt, want_blk = O.select_evictions(
        nn(st.kv_metrics), nn(st.kv_position), nn(st.block_tables), nn(ctx), nn(hanging), nn(offs), nn(budgets),
        nn(last), nn(protected), B, MAX_INT, None, None, 0.0, False, 0)
    np.testing.assert_array_equal(outs[2][1], want_cnt)
    np.testing.assert_array_equal(outs[2][2], want_blk)
    np.testing.assert_array_equal(outs[2][0], want_logical)
    assert int(want_cnt.sum()) > 0


def test_block_state_ops_match_reference_formulas(capi):
    """prepare_compression / free_and_append_slots vs the reference's torch formulas
    (block.py:184-210, 275-335; scheduler.py:131-141, 274-280) evaluated on the CPU."""
    from vllm_kvcompress_b200 import _custom_ops as ops
    from vllm_kvcompress_b200 import workload as W
    g = torch.Generator().manual_seed(5)
    S, L, H, B, MB = 37, 5, 8, 16, 12
    ctx = torch.randint(1, MB * B - 20, (L, S, H), generator=g, dtype=torch.int32)
    st = W.make_state(S, L, H, B, 16, ctx, device="cpu", seed=5, fill_cache=False, max_blocks=MB,
                      blocks_per_head_alloc=MB, slack=1.0)
    want_hang = W.hanging_token_counts(ctx, B).transpose(0, 1).contiguous()
    want_off = W.evicted_kv_offsets(ctx, B)
    nblk = (ctx + B - 1) // B
    max_blocks = 200
    want_budget = (nblk.sum(dim=(0, 2)) - max_blocks).clamp(min=0).to(torch.int32)
    dctx = ctx.cuda()
    hang = torch.empty((S, L, H), dtype=torch.int32, device="cuda")
    off = torch.empty_like(hang)
    bud = torch.empty(S, dtype=torch.int32, device="cuda")
    ops.prepare_compression(dctx, hang, off, bud, B, max_blocks)
    assert torch.equal(hang.cpu(), want_hang) and torch.equal(off.cpu(), want_off) and torch.equal(bud.cpu(), want_budget)
    # free + append
    ev = torch.randint(0, 3, (S, L, H), generator=g, dtype=torch.int32)
    ev = torch.minimum(ev, ((ctx.transpose(0, 1) - 1) // B).to(torch.int32))
    freed = (ev * B - (B - want_hang)).clamp(min=0)
    ctx_after = ctx - freed.transpose(0, 1)
    blk = torch.gather(st.block_tables, 3, (ctx_after // B).long().unsqueeze(-1)).squeeze(-1)
    want_slot = blk.long() * B + (ctx_after % B).long()
    seq_len = torch.randint(100, 9000, (S,), generator=g, dtype=torch.int32)
    want_pos = st.kv_position.clone()
    want_pos.view(-1)[want_slot.flatten()] = seq_len.view(1, -1, 1).expand_as(ctx).flatten()
    slot = torch.empty((L, S, H), dtype=torch.int64, device="cuda")
    pos = st.kv_position.cuda()
    ops.free_and_append_slots(dctx, st.block_tables.cuda(), ev.cuda(), hang, seq_len.cuda(), slot, pos, B, True)
    assert torch.equal(dctx.cpu(), ctx_after + 1)
    assert torch.equal(slot.cpu(), want_slot) and torch.equal(pos.cpu(), want_pos)


@pytest.mark.parametrize("cfg", [(81, 6, 32, 8, 160, False), (82, 3, 64, 8, 300, True), (83, 4, 8, 8, 33, False),
                                 # every head's new token opens a fresh block (ctx = 16k + 1): the kernel does
                                 # not fetch that block at all and serves it from the step's k/v registers
                                 (84, 6, 32, 8, 160, False, True), (85, 2, 64, 8, 49, True, True)])
def test_fused_step_equals_cache_write_then_attention(capi, cfg):
    """kvcompress_paged_attention_fused_step == reshape_and_cache_kvc followed by the fused-metrics
    attention (bit-exact caches, identical recording rule, outputs within the attention tolerance)."""
    from vllm_kvcompress_b200 import _custom_ops as ops
    seed, S, Hq, H, max_ctx, bf = cfg[:6]
    c = attn_case(seed, S, Hq, H, max_ctx, bf, False, 0, min_ctx=1)
    D, B = 128, 16
    if len(cfg) > 6:
        c["context_lens"] = ((c["context_lens"] - 1) // B * B + 1).astype(np.int32)
    q, kc, vc = cu16(c["query"], bf), cu16(c["key_cache"], bf), cu16(c["value_cache"], bf)
    NB = kc.shape[0]
    bt, cl = cu(c["block_tables"], torch.int32), cu(c["context_lens"], torch.int32)
    kp, lp, bl = (cu(c[k], torch.int32) for k in ("kv_position", "last_position", "kv_metric_buffer_len"))
    g = torch.Generator().manual_seed(seed)
    newk = ((torch.rand((S, H, D), generator=g) - 0.5) * 0.3).to(q.dtype).cuda()
    newv = ((torch.rand((S, H, D), generator=g) - 0.5) * 0.3).to(q.dtype).cuda()
    # the new token is the LAST token of every head (context_lens already counts it)
    last_idx = (cl - 1).long()
    slot = (torch.gather(bt, 2, (last_idx // B).unsqueeze(-1)).squeeze(-1).long() * B + last_idx % B).contiguous()
    slot[0, 0] = -1 if S > 3 else slot[0, 0]   # a head without a new token (padding)
    bias = torch.rand(H, generator=g).cuda()
    metrics0 = torch.rand((NB, B), generator=g).cuda()
    max_c = int(c["context_lens"].max())
    # reference order: cache write, then attention with fused metrics
    k1, v1, m1 = kc.clone(), vc.clone(), metrics0.clone()
    out1 = torch.zeros((S, Hq, D), dtype=q.dtype, device="cuda")
    ops.reshape_and_cache_kvc(newk, newv, k1, v1, m1, slot.flatten(), bias, "auto", 1.0, 1.0)
    ops.paged_attention_kvc_fused_metrics(out1, m1, q, k1, v1, H, float(c["scale"]), bt, cl, kp, lp, bl, B, max_c)
    k2, v2, m2 = kc.clone(), vc.clone(), metrics0.clone()
    out2 = torch.zeros_like(out1)
    ops.paged_attention_kvc_fused_step(out2, m2, q, newk, newv, k2, v2, slot.flatten(), bias, H, float(c["scale"]),
                                       bt, cl, kp, lp, bl, B, max_c)
    torch.cuda.synchronize()
    assert torch.equal(k1.view(torch.int16), k2.view(torch.int16))
    assert torch.equal(v1.view(torch.int16), v2.view(torch.int16))
    torch.testing.assert_close(out2.float(), out1.float(), atol=1e-3, rtol=1e-2)
    torch.testing.assert_close(m2, m1, atol=1e-6, rtol=1e-4)


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("fullrow", ["0", "1"])
def test_execute_moves_many_heads_small_counts_vs_oracle(capi, O, deferred, fullrow):
    """>= 4096 heads (the launch splits: heads with <= 1 move go to the light kernel, the rest to the run
    kernel) with 0..5 moves per head, bit-exact against the C oracle.  deferred: every other head's newest
    token is parked in pend_kv instead of the cache (garbage at its cache slot); the op must leave the cache
    as if the token had been stored first -- except at the source slot of a parked token that a move consumed:
    that slot is never written (its block is freed right after the compression), so it is left out."""
    import subprocess, sys, textwrap
    # KVC_MOVES_FULLROW is read once per process: run the body in a child so that both settings are covered
    code = textwrap.dedent(f"""
        import numpy as np, torch, sys
        sys.path.insert(0, {ROOT!r})
        from oracle import kvc_oracle as O
        from vllm_kvcompress_b200 import capi
        O.build()
        rs = np.random.RandomState(5)
        S, L, H, B, D = 6, 12, 64, 16, 128          # 4608 heads
        heads = S * L * H
        blocks_per_head = 3
        NB = heads * blocks_per_head + 7
        perm = rs.permutation(NB)[: heads * blocks_per_head].reshape(heads, blocks_per_head)
        kc = rs.randint(0, 65536, size=(NB, D // 8, B, 8)).astype(np.uint16)
        vc = rs.randint(0, 65536, size=(NB, D, B)).astype(np.uint16)
        km = rs.rand(NB, B).astype(np.float32)
        kp = rs.randint(0, 6000, size=(NB, B)).astype(np.int32)
        cnt = rs.choice([0, 1, 1, 1, 2, 3, 5], size=heads).astype(np.int32)
        offs = (np.arange(heads) * blocks_per_head * B).astype(np.int32)
        moves = np.zeros((heads * blocks_per_head * B, 2), np.int32)
        # per head (vectorised): destinations in the first two blocks, sources in the last block (descending,
        # like K4), the newest token first in 70 % of the heads
        ctx_tail = rs.randint(1, B + 1, size=heads)                 # tokens in the last block
        cnt = np.minimum(cnt, ctx_tail).astype(np.int32)
        order = np.argsort(rs.rand(heads, 2 * B), axis=1)[:, :5]    # 5 distinct destinations per head
        order = np.where(np.arange(5)[None] < cnt[:, None], order, 10 ** 6)
        dst = np.sort(order, axis=1)
        lo = np.maximum(cnt - 1, 0)
        top = np.where(rs.rand(heads) < 0.7, ctx_tail - 1, lo + (rs.rand(heads) * (ctx_tail - lo)).astype(np.int64))
        gidx = np.arange(heads)
        phys = lambda i: perm[gidx, i // B] * B + i % B
        for k in range(5):
            sel = cnt > k
            d = np.where(sel, dst[:, k], 0)
            sidx = np.where(sel, 2 * B + top - k, 0)
            pd, ps_ = phys(d), phys(sidx)
            moves[offs[sel] + k, 0] = pd[sel]
            moves[offs[sel] + k, 1] = ps_[sel]
        pend_slot = np.where(gidx % 2 == 0, phys(2 * B + ctx_tail - 1), -1).astype(np.int64)   # newest token
        consumed = pend_slot[(gidx % 2 == 0) & (cnt > 0) & (top == ctx_tail - 1)]
        # oracle on the canonical cache
        wk, wv, wm, wp = kc.copy(), vc.copy(), km.copy(), kp.copy()
        O.execute_cache_moves(wk, wv, wm, wp, moves, cnt.reshape(S, L, H), offs.reshape(S, L, H))
        dev = "cuda"
        t16 = lambda a: torch.as_tensor(a.view(np.int16)).to(dev).view(torch.float16)
        gk, gv = t16(kc), t16(vc)
        gm, gp = torch.as_tensor(km).to(dev), torch.as_tensor(kp).to(dev)
        gmoves, gcnt, goffs = (torch.as_tensor(moves).to(dev), torch.as_tensor(cnt.reshape(S, L, H)).to(dev),
                               torch.as_tensor(offs.reshape(S, L, H)).to(dev))
        if {deferred!r}:
            # pend arrays are [L, S, H]; head g = (s * L + l) * H + h
            ps = torch.as_tensor(pend_slot.reshape(S, L, H)).permute(1, 0, 2).contiguous().to(dev)
            pkv = torch.zeros((L, S, H, 2 * D), dtype=torch.float16, device=dev)
            sl = ps.clamp(min=0)
            kflat = gk.permute(0, 2, 1, 3).reshape(NB * B, D)      # [slot, D]
            vflat = gv.permute(0, 2, 1).reshape(NB * B, D)
            pkv[..., :D] = kflat[sl]
            pkv[..., D:] = vflat[sl]
            # scribble over the cache copy of every parked token: the op must not depend on it
            live = ps.reshape(-1)[ps.reshape(-1) >= 0]
            kview = gk.view(torch.int16)
            vview = gv.view(torch.int16)
            blk, o = live // B, live % B
            kview[blk, :, o, :] = 0x7e00
            vview[blk, :, o] = 0x7e00
            capi.execute_cache_moves_deferred(gk, gv, gm, gp, gmoves, gcnt, goffs, pkv, ps)
        else:
            capi.execute_cache_moves(gk, gv, gm, gp, gmoves, gcnt, goffs)
        torch.cuda.synchronize()
        b16 = lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)
        rk, rv = b16(gk), b16(gv)
        if {deferred!r}:
            assert len(consumed) > 500
            c = np.asarray(consumed)
            assert (rk[c // B, :, c % B, :] == 0x7e00).all() and (rv[c // B, :, c % B] == 0x7e00).all()
            rk[c // B, :, c % B, :] = wk[c // B, :, c % B, :]
            rv[c // B, :, c % B] = wv[c // B, :, c % B]
        assert np.array_equal(rk, wk), "k_cache"
        assert np.array_equal(rv, wv), "v_cache"
        assert np.array_equal(gm.cpu().numpy(), wm), "metrics"
        assert np.array_equal(gp.cpu().numpy(), wp), "positions"
        print("ok")
    """)
    env = dict(os.environ, KVC_MOVES_FULLROW=fullrow)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_k4_few_heads_long_lists_vs_oracle(capi, O):
    """schedule_t1_cache_moves at the post-prefill shape (few heads, block tables of more than 512 slots: one CTA
    per head over the gaps of the evicted tail): counts and (dst, src) pairs bit-exact against the C oracle for
    eviction patterns that make the gaps degenerate -- nothing evicted, everything evicted, only the oldest / only
    the newest tokens evicted (one gap holds every survivor), every other token, and random sets at CR 64 ... 1.3."""
    rs = np.random.RandomState(77)
    S, L, H, B, MB = 2, 3, 8, 16, 130            # 48 heads, up to 2080 slots each
    ctx = rs.randint(600, MB * B + 1, size=(L, S, H)).astype(np.int32)
    ctx[0, 0, 0] = MB * B                        # a full table
    ctx[1, 1, 3] = 1                             # and a single token
    nblk = (ctx + B - 1) // B
    NB = int(nblk.sum()) + 5
    perm = rs.permutation(NB).astype(np.int32)
    bt = np.zeros((L, S, H, MB), np.int32)
    k = 0
    for idx in np.ndindex(L, S, H):
        n = int(nblk[idx])
        bt[idx][:n] = perm[k:k + n]
        k += n
    ctx_slh = ctx.transpose(1, 0, 2)
    alloc = (nblk.transpose(1, 0, 2) * B).reshape(-1)
    offs = np.concatenate([[0], np.cumsum(alloc)[:-1]]).astype(np.int32).reshape(S, L, H)
    N = int(alloc.sum())
    cnt = np.zeros((S, L, H), np.int32)
    elog = np.full(N, MAX_INT, np.int32)
    for i, idx in enumerate(np.ndindex(S, L, H)):
        n = int(ctx_slh[idx])
        kind = i % 8
        if kind == 0:
            ev = np.arange(0)                                 # nothing
        elif kind == 1:
            ev = np.arange(n)                                 # everything
        elif kind == 2:
            ev = np.arange(n - n // 3)                        # the oldest two thirds: survivors are all in the tail
        elif kind == 3:
            ev = np.arange(n // 4, n)                         # the newest three quarters: no move at all
        elif kind == 4:
            ev = np.arange(0, n, 2)                           # every other token
        else:
            keep = max(1, int(n / [64, 8, 1.3][kind - 5]))
            ev = np.sort(rs.choice(n, size=n - keep, replace=False))
        cnt[idx] = len(ev)
        elog[offs[idx]:offs[idx] + len(ev)] = ev
    wm = np.zeros((N, 2), np.int32)
    wc = np.zeros((S, L, H), np.int32)
    O.schedule_cache_moves(wm, wc, elog, cnt, offs, bt, ctx, B)
    moves = torch.zeros((N, 2), dtype=torch.int32, device="cuda")
    mcnt = torch.full((S, L, H), -3, dtype=torch.int32, device="cuda")
    capi.schedule_t1_cache_moves(moves, mcnt, cu(elog), cu(cnt), cu(offs), cu(bt), cu(ctx), B)
    got_c, got_m = mcnt.cpu().numpy(), moves.cpu().numpy()
    np.testing.assert_array_equal(got_c, wc)
    valid = np.zeros(N, bool)
    for o, c in zip(offs.reshape(-1), wc.reshape(-1)):
        valid[o:o + c] = True
    np.testing.assert_array_equal(got_m[valid], wm[valid])
    assert (got_m[~valid] == 0).all()
    assert int(wc.sum()) > 5000 and int((wc == 0).sum()) >= 12


def test_k4_many_heads_short_lists_vs_oracle(capi, O):
    """schedule_t1_cache_moves at decode-time shape (>= 4096 heads: 32 heads per warp, heads with at most one
    evicted slot settled by their own lane, longer lists by the whole warp): counts and (dst, src) pairs
    bit-exact against the C oracle, incl. evicting the last token itself (no move) and lists of 2..20."""
    rs = np.random.RandomState(9)
    S, L, H, B, MB = 6, 12, 64, 16, 6            # 4608 heads
    heads = S * L * H
    ctx = rs.randint(1, MB * B + 1, size=(L, S, H)).astype(np.int32)
    nblk = (ctx + B - 1) // B
    NB = int(nblk.sum()) + 5
    perm = rs.permutation(NB).astype(np.int32)
    bt = np.zeros((L, S, H, MB), np.int32)
    k = 0
    for idx in np.ndindex(L, S, H):
        n = int(nblk[idx])
        bt[idx][:n] = perm[k:k + n]
        k += n
    ctx_slh = ctx.transpose(1, 0, 2)
    alloc = (nblk.transpose(1, 0, 2) * B).reshape(-1)
    offs = np.concatenate([[0], np.cumsum(alloc)[:-1]]).astype(np.int32).reshape(S, L, H)
    N = int(alloc.sum())
    cnt = np.zeros((S, L, H), np.int32)
    elog = np.full(N, MAX_INT, np.int32)
    for idx in np.ndindex(S, L, H):
        n = int(ctx_slh[idx])
        c = int(rs.choice([0, 1, 1, 1, 1, 2, 3, 7, 20]))
        c = min(c, n)
        if c == 1 and rs.rand() < 0.15:
            ev = np.array([n - 1])                # the last token itself: nothing to move
        else:
            ev = np.sort(rs.choice(n, size=c, replace=False))
        cnt[idx] = c
        elog[offs[idx]:offs[idx] + c] = ev
    wm = np.zeros((N, 2), np.int32)
    wc = np.zeros((S, L, H), np.int32)
    O.schedule_cache_moves(wm, wc, elog, cnt, offs, bt, ctx, B)
    moves = torch.zeros((N, 2), dtype=torch.int32, device="cuda")
    mcnt = torch.full((S, L, H), -3, dtype=torch.int32, device="cuda")
    capi.schedule_t1_cache_moves(moves, mcnt, cu(elog), cu(cnt), cu(offs), cu(bt), cu(ctx), B)
    got_c, got_m = mcnt.cpu().numpy(), moves.cpu().numpy()
    np.testing.assert_array_equal(got_c, wc)
    valid = np.zeros(N, bool)
    for o, c in zip(offs.reshape(-1), wc.reshape(-1)):
        valid[o:o + c] = True
    np.testing.assert_array_equal(got_m[valid], wm[valid])
    assert (got_m[~valid] == 0).all()
    assert int((wc == 1).sum()) > 1000 and int((wc > 1).sum()) > 500
