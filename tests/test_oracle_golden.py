"""CPU: the C oracle (oracle/kvc_oracle.c) against golden vectors produced by the
reference's own Python oracles (tests/golden/make_golden.py).  Integer/byte work
must be bit-exact; attention within the reference's own tolerance
(tests/kernels/test_kvcompress_attention.py:356-357: atol 1e-3, rtol 1e-5)."""
import numpy as np
import pytest

from oracle import kvc_oracle as O
from tests.golden_util import load_cases

MAX_INT = O.MAX_INT

EVICT = load_cases("evictions.npz")
COUNT = load_cases("count_blocks.npz")
MOVES = load_cases("moves.npz")
ATTN = load_cases("attention.npz")
CACHE = load_cases("cache_write.npz")


def run_oracle_evictions(c, budgets=None, truncate=None):
    S, L, H, B = int(c["S"]), int(c["L"]), int(c["H"]), int(c["B"])
    N = int(c["total_kv"])
    ekv = np.zeros(N, np.int32)
    elog = np.full(N, -7, np.int32)
    cnt = np.full((S, L, H), 99, np.int32)  # kernel must zero it itself (:80-83)
    rem = c["hanging_token_count"].astype(np.int32).copy()
    O.schedule_cache_evictions(
        ekv, elog, cnt, rem, c["evicted_kv_offsets"], c["sorted_indices"], c["seq_block_offsets"],
        c["layer_by_block"], c["head_by_block"], c["virtual_block_num_by_block"],
        c["budgets"] if budgets is None else budgets, c["context_lens"], c["hanging_token_count"],
        c["kv_position"], c["last_position"], c["protected"], B, False, None, N, MAX_INT,
        bool(c["truncate"]) if truncate is None else truncate)
    return ekv, elog, cnt, rem


@pytest.mark.parametrize("name", sorted(EVICT))
def test_schedule_cache_evictions_matches_reference_python(name):
    c = EVICT[name]
    ekv, elog, cnt, _ = run_oracle_evictions(c)
    np.testing.assert_array_equal(cnt, c["out_evicted_kv_count"])
    if int(c["truncate"]):
        # fully determined outputs
        np.testing.assert_array_equal(elog, c["out_evicted_logical_indices"])
        np.testing.assert_array_equal(ekv, c["out_evicted_kv_indices"])
    else:
        # only the written prefix of every head segment is defined; the Python
        # reference pre-fills evicted_kv_indices with MAX_INT (:831), the kernel
        # leaves the rest untouched.
        offs = c["evicted_kv_offsets"].reshape(-1)
        for g, (o, n) in enumerate(zip(offs, cnt.reshape(-1))):
            np.testing.assert_array_equal(elog[o:o + n], c["out_evicted_logical_indices"][o:o + n])
            np.testing.assert_array_equal(ekv[o:o + n], c["out_evicted_kv_indices"][o:o + n])
        written = np.zeros(len(elog), bool)
        for o, n in zip(offs, cnt.reshape(-1)):
            written[o:o + n] = True
        assert (elog[~written] == -7).all() and (ekv[~written] == 0).all()


def test_zero_budget_is_a_noop_with_zero_counts():
    """K1 early return (csrc/kvcompress_eviction_kernels.cu:80-87); the reference's
    Python oracle cannot express it (see make_golden.py)."""
    c = EVICT["c1_full_p0_e8"]
    ekv, elog, cnt, rem = run_oracle_evictions(c, budgets=np.zeros(1, np.int32), truncate=False)
    assert (cnt == 0).all() and (elog == -7).all() and (ekv == 0).all()
    np.testing.assert_array_equal(rem, c["hanging_token_count"])
    # with truncate every slot becomes (head id, NULL)
    ekv, elog, cnt, _ = run_oracle_evictions(c, budgets=np.zeros(1, np.int32), truncate=True)
    assert (cnt == 0).all() and (elog == MAX_INT).all()
    np.testing.assert_array_equal(ekv, np.repeat(np.arange(4, dtype=np.int32), 256))


def test_eviction_semantic_invariants_config1():
    """Size-independent properties on BASELINE configs[0] (4 heads x 256 KV):
    freed-block total equals the budget, nothing protected is evicted, per-head
    evictions are the head's lowest-metric candidates."""
    for name, c in EVICT.items():
        if not name.startswith("c1_"):
            continue
        ekv, elog, cnt, _ = run_oracle_evictions(c)
        B = int(c["B"])
        hang = c["hanging_token_count"].reshape(-1)
        cntf = cnt.reshape(-1)
        freed = np.where(cntf >= hang, (cntf - hang) // B + 1, 0).sum()
        assert freed == int(c["budgets"][0])
        offs = c["evicted_kv_offsets"].reshape(-1)
        maxpos = int(c["last_position"][0]) - int(c["protected"][0])
        for h, (o, n) in enumerate(zip(offs, cntf)):
            idx = elog[o:o + n]
            assert (idx <= maxpos).all()  # kv_position == logical index in config 1
            assert len(np.unique(idx)) == n
            # metric of evicted <= metric of every non-evicted candidate of the head
            blocks = np.where(c["head_by_block"] == h)[0]
            met = {}
            for b in blocks:
                for o_in in range(B):
                    v = int(c["virtual_block_num_by_block"][b]) * B + o_in
                    if v < int(c["context_lens"].reshape(-1)[h]) and v <= maxpos:
                        met[v] = float(c["metrics"][b, o_in])
            if n:
                worst = max(met[int(v)] for v in idx)
                rest = [m for v, m in met.items() if v not in set(idx.tolist())]
                assert not rest or worst < min(rest)


@pytest.mark.parametrize("name", sorted(COUNT))
def test_count_block_evictions(name):
    c = COUNT[name]
    logical = c["in_evicted_logical_indices"].astype(np.int32).copy()
    out = np.zeros(c["out_evicted_block_count"].shape, np.int32)
    O.count_block_evictions(out, logical, c["evicted_kv_offsets"], c["hanging_token_count"],
                            int(c["B"]), MAX_INT)
    np.testing.assert_array_equal(out, c["out_evicted_block_count"])
    np.testing.assert_array_equal(logical, c["out_evicted_logical_indices"])


@pytest.mark.parametrize("name", sorted(MOVES))
def test_schedule_and_execute_moves(name):
    c = MOVES[name]
    B = int(c["B"])
    moves = np.zeros((int(c["total_kv"]), 2), np.int32)
    mcnt = np.zeros(c["evicted_kv_count"].shape, np.int32)
    O.schedule_cache_moves(moves, mcnt, c["evicted_logical_indices"], c["evicted_kv_count"],
                           c["evicted_kv_offsets"], c["block_tables"], c["context_lens_lsh"], B)
    np.testing.assert_array_equal(mcnt, c["out_cache_moves_count"])
    np.testing.assert_array_equal(moves, c["out_cache_moves_idx"])
    kc, vc = c["k_cache"].copy(), c["v_cache"].copy()
    km, kp = c["kv_metrics"].copy(), c["kv_position"].copy()
    O.execute_cache_moves(kc, vc, km, kp, moves, mcnt, c["evicted_kv_offsets"])
    np.testing.assert_array_equal(kc, c["out_k_cache"])
    np.testing.assert_array_equal(vc, c["out_v_cache"])
    np.testing.assert_array_equal(km, c["out_kv_metrics"])
    np.testing.assert_array_equal(kp, c["out_kv_position"])


def test_moves_closed_form():
    """dst_k = k-th smallest evicted index below new_n, src_k = k-th
    largest surviving index >= new_n, m = |E & [0, new_n)|."""
    for c in MOVES.values():
        B = int(c["B"])
        S, L, H = c["evicted_kv_count"].shape
        for s in range(S):
            for l in range(L):
                for h in range(H):
                    n = int(c["context_lens_lsh"][l, s, h])
                    e = int(c["evicted_kv_count"][s, l, h])
                    o = int(c["evicted_kv_offsets"][s, l, h])
                    E = c["evicted_logical_indices"][o:o + e]
                    new_n = n - e
                    dst = [int(v) for v in E if v < new_n]
                    Eset = set(E.tolist())
                    src = [i for i in range(n - 1, new_n - 1, -1) if i not in Eset]
                    m = int(c["out_cache_moves_count"][s, l, h])
                    assert m == len(dst) == len(src)
                    bt = c["block_tables"][l, s, h]
                    phys = lambda i: int(bt[i // B]) * B + i % B
                    got = c["out_cache_moves_idx"][o:o + m]
                    assert [tuple(r) for r in got] == [(phys(d), phys(sv)) for d, sv in zip(dst, src)]


@pytest.mark.parametrize("name", sorted(ATTN))
def test_paged_attention(name):
    c = ATTN[name]
    S, Hq, H, D, B = (int(c[k]) for k in ("S", "Hq", "H", "D", "B"))
    Q = Hq // H
    out = np.zeros((S, Hq, D), np.float32)
    metric = np.full(c["kv_metric_out"].shape, -1.0, np.float32)
    slopes = c["alibi_slopes"] if c["alibi_slopes"].size else None
    O.paged_attention_kvc(out, metric, c["query"], c["key_cache"], c["value_cache"], H,
                          float(c["scale"]), c["block_tables"], c["context_lens"], c["kv_position"],
                          c["last_position"], c["kv_metric_buffer_len"], slopes,
                          bool(c["is_bf16"]), True)
    np.testing.assert_allclose(out, c["out"], atol=1e-3, rtol=1e-5)
    untouched = c["kv_metric_out"] == -1.0
    assert ((metric == -1.0) == untouched).all()
    # the oracle divides by (sum + 1e-6) like the kernel (:298); torch.softmax does not
    np.testing.assert_allclose(metric[~untouched], c["kv_metric_out"][~untouched], atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("name", sorted(CACHE))
def test_reshape_and_cache(name):
    c = CACHE[name]
    kc, vc, m = c["key_cache"].copy(), c["value_cache"].copy(), c["kv_metrics"].copy()
    O.reshape_and_cache_kvc(c["key"], c["value"], kc, vc, m, c["slot_mapping"], c["kv_metric_head_bias"])
    np.testing.assert_array_equal(kc, c["out_key_cache"])
    np.testing.assert_array_equal(vc, c["out_value_cache"])
    np.testing.assert_array_equal(m, c["out_kv_metrics"])


def test_aggregate_decode():
    rng = np.random.RandomState(0)
    temp = rng.rand(40, 16, 4).astype(np.float32)
    metrics = rng.rand(40, 16).astype(np.float32)
    want = metrics + (temp ** 2).sum(-1)
    O.aggregate_decode(metrics, temp)
    np.testing.assert_allclose(metrics, want, rtol=1e-6)


# ---------------------------------------------------------------------------
# prefill metric initialisation (_naive_kvc_attention + aggregate_prefill)
# ---------------------------------------------------------------------------
PREFILL = load_cases("prefill_metrics.npz")


def oracle_prefill(c):
    return O.prefill_kv_metrics(c["query"], c["key"], c["seq_lens"].tolist(), float(c["scale"]),
                                c["kv_metric_buffer_len"], int(c["n_observed"]), 4096, bool(c["use_l2"]),
                                bool(c["use_average"]), bool(c["use_maxpool"]), bool(c["is_bf16"]))


@pytest.mark.parametrize("name", sorted(PREFILL))
def test_prefill_kv_metrics_matches_reference_python(name):
    """fp32 softmax of fp16/bf16-rounded logits: tolerance rtol 1e-5 / atol 1e-6 against
    _naive_kvc_attention run on CPU tensors (summation order is the only difference)."""
    c = PREFILL[name]
    out = oracle_prefill(c)
    np.testing.assert_allclose(out, c["out_kv_metric"], rtol=1e-5, atol=1e-6)
    m = c["kv_metrics_in"].copy()
    O.aggregate_prefill(m, c["out_kv_metric"], c["slot_mapping"])
    np.testing.assert_allclose(m, c["out_kv_metrics"], rtol=1e-6, atol=1e-7)


def test_prefill_observation_chunks_sum():
    """max_observed_block_size splits the observed queries into chunks whose (separately pooled)
    metrics add up (flash_attn.py:1146-1160): without max-pool the split is invisible."""
    c = PREFILL["prefill0"]
    args = (c["query"], c["key"], c["seq_lens"].tolist(), float(c["scale"]), c["kv_metric_buffer_len"], 32)
    one = O.prefill_kv_metrics(*args, 4096, True, False, False, False)
    two = O.prefill_kv_metrics(*args, 8, True, False, False, False)
    np.testing.assert_allclose(one, two, rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------
# live ("V2") selection incl. head bias / use_average / attention sinks
# ---------------------------------------------------------------------------
SELV2 = load_cases("selection_v2.npz")


def oracle_select_v2(c):
    one = lambda v: np.array([int(v)], np.int32)
    return O.select_evictions(
        c["kv_metrics"], c["kv_position"], c["block_tables"], c["context_lens"], c["hanging_token_count"],
        c["evicted_kv_offsets"], one(c["budget"]), one(c["last_position"]), one(c["protected"]), int(c["B"]),
        MAX_INT, c["head_bias"] if int(c["has_bias"]) else None, c["position_bins"], float(c["bias_weight"]),
        bool(c["use_average"]), int(c["num_sinks"]))


@pytest.mark.parametrize("name", sorted(SELV2))
def test_select_evictions_matches_reference_schedule_evictions(name):
    """bit-exact against CompressionMetrics.schedule_evictions run on CPU tensors."""
    c = SELV2[name]
    logical, kv_count, blk_count = oracle_select_v2(c)
    np.testing.assert_array_equal(kv_count, c["out_evicted_kv_count"])
    np.testing.assert_array_equal(blk_count, c["out_evicted_block_count"])
    np.testing.assert_array_equal(logical, c["out_evicted_logical_indices"])
    assert int(blk_count.sum()) == min(int(c["budget"]), int(blk_count.sum()))


# ---------------------------------------------------------------------------
# block manager / allocator
# ---------------------------------------------------------------------------
BLOCKMGR = load_cases("block_manager.npz")


def replay_block_manager(c, mgr, check):
    """Replays the golden script on `mgr` (oracle or CUDA adapter), calling check(step_index, prefix)."""
    for i in range(int(c["num_steps"])):
        pre = f"s{i:03d}_"
        op = str(c[pre + "op"])
        slots = c[pre + "slots"].tolist()
        if op == "add":
            mgr.add_sequences(slots, c[pre + "lens"].tolist())
        elif op == "append":
            mgr.append_slots(slots, c[pre + "positions"].tolist())
        elif op == "free":
            mgr.free_compressed(slots, c[pre + "counts"])
        elif op == "remove":
            mgr.remove_sequences(slots)
        check(i, pre)


def assert_block_state_equal(c, pre, B, block_tables, context_lens, free_mask, seq_idx, layer_idx, head_idx,
                             logical_idx, token_positions):
    np.testing.assert_array_equal(context_lens, c[pre + "context_lens"])
    np.testing.assert_array_equal(free_mask.astype(bool), c[pre + "free_mask"])
    MB = block_tables.shape[-1]
    alloc = np.arange(MB)[None, None, None] < ((context_lens + B - 1) // B)[..., None]
    np.testing.assert_array_equal(block_tables[alloc], c[pre + "block_tables"][alloc])
    np.testing.assert_array_equal(seq_idx, c[pre + "seq_index_by_block"])
    used = seq_idx >= 0   # metadata of live blocks (freed blocks keep stale values in the reference too)
    np.testing.assert_array_equal(layer_idx[used], c[pre + "layer_index_by_block"][used])
    np.testing.assert_array_equal(head_idx[used], c[pre + "head_index_by_block"][used])
    np.testing.assert_array_equal(logical_idx[used], c[pre + "logical_block_num_by_block"][used])
    np.testing.assert_array_equal(token_positions[used], c[pre + "token_positions"][used])


@pytest.mark.parametrize("name", sorted(BLOCKMGR))
def test_block_manager_oracle_matches_reference_python(name):
    c = BLOCKMGR[name]
    L, H, B, MS, MB, NB = (int(c[k]) for k in ("L", "H", "B", "MS", "MB", "NB"))
    mgr = O.BlockManagerOracle(L, H, B, MS, MB, NB)

    def check(i, pre):
        assert mgr.free_count == int(c[pre + "free_count"])
        assert_block_state_equal(c, pre, B, mgr.block_tables, mgr.context_lens, mgr.free_mask,
                                 mgr.seq_index_by_block, mgr.layer_index_by_block, mgr.head_index_by_block,
                                 mgr.logical_block_num_by_block, mgr.token_positions)

    replay_block_manager(c, mgr, check)
