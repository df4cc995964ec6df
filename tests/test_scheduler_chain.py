"""The reference's OWN scheduler as the oracle of a multi-step chain (tests/golden/scheduler_chain.npz, made by
tests/golden/make_golden.py::gen_scheduler_chain: vllm/kvcompress/scheduler.py CompressionScheduler driving
block_manager.py / block.py / metrics.py on CPU tensors for 24 steps over four sequences with different
policies, a max_kv_per_compression cap and compression_interval = 2).

vllm_kvcompress_b200.scheduler.CompressionScheduler (same method names and results) is replayed over the same
scenario and must reproduce, step by step and bit for bit: which sequences are compressed, the freed-block counts,
the move lists (physical slot ids: the allocator hands out the same blocks as the reference's), and the block
tables / context lengths / free mask / metrics / positions / K / V afterwards.

  * CPU (`-m "not gpu"`): backend = the numpy / C oracle (oracle/kvc_oracle.py).  This pins the oracle's
    select_evictions for MULTI-sequence calls against the live reference, including its `evicted_infs` offset
    behaviour (metrics.py:718) through scheduler.reference_v2_budgets.
  * GPU (`-m gpu`): backend = scheduler.DeviceBackend: the CUDA ops and the device-side allocator.
"""
import os
import types

import numpy as np
import pytest

from tests.golden.make_golden import chain_key, chain_noise

GOLD = os.path.join(os.path.dirname(__file__), "golden", "scheduler_chain.npz")
MAX_INT = 2147483000


class Gold:
    def __init__(self):
        self.d = np.load(GOLD)

    def __call__(self, k):
        return self.d["chain/" + k]

    def step(self, t, k):
        return self.d[f"chain/s{t:03d}_{k}"]


class _Seq:
    def __init__(self, seq_id, n):
        self.seq_id, self.compressed, self._len = seq_id, False, n
        self.data = types.SimpleNamespace(get_len=lambda: self._len)


def make_seqs(g):
    seqs, params = [], []
    for sid, n, rate1000, mct, prot, once in g("seq_specs").tolist():
        seqs.append(_Seq(sid, n))
        params.append(types.SimpleNamespace(target_compression_rate=rate1000 / 1000.0, max_cache_tokens=mct,
                                            protected_window_size=prot, compress_once=bool(once)))
    return seqs, params


def live_slots(block_tables, context_lens, positions, slots_of_seqs, L, H, B):
    out_s, out_k, newest = [], [], []
    flat_pos = positions.reshape(-1)
    for s_idx in slots_of_seqs:
        for l in range(L):
            for h in range(H):
                n = int(context_lens[l, s_idx, h])
                idx = np.arange(n)
                sl = block_tables[l, s_idx, h, idx // B].astype(np.int64) * B + idx % B
                out_s.append(sl)
                out_k.append(chain_key(s_idx, l, h, flat_pos[sl], L, H))
                nw = np.zeros(n, bool)
                nw[n - 1] = True
                newest.append(nw)
    return np.concatenate(out_s), np.concatenate(out_k), np.concatenate(newest)


def write_payload(k_cache, v_cache, slots, keys, B, D, x):
    blk, off = slots // B, slots % B
    for j in range(D // x):
        for e in range(x):
            k_cache[blk, j, off, e] = ((keys >> (3 * (j * x + e))) & 7).astype(np.float16)
    for d in range(D):
        v_cache[blk, d, off] = ((keys >> (3 * d + 1)) & 7).astype(np.float16)


# --------------------------------------------------------------------------------------------------------------
# CPU backend over the oracle
# --------------------------------------------------------------------------------------------------------------
class OracleBackend:
    def __init__(self, g):
        from oracle import kvc_oracle as O
        self.O = O
        L, H, B, MS, MB, NB = (int(g(k)) for k in ("L", "H", "B", "MS", "MB", "NB"))
        self.L, self.H, self.B = L, H, B
        bm = O.BlockManagerOracle(L, H, B, MS, MB, NB)
        bm.block_tables[:] = g("init_block_tables")
        bm.context_lens[:] = g("init_context_lens")
        bm.free_mask[:] = g("init_free_mask")
        bm.token_positions[:] = g("init_token_positions")
        for k in ("seq_index_by_block", "layer_index_by_block", "head_index_by_block", "logical_block_num_by_block"):
            getattr(bm, k)[:] = g("init_" + k)
        self.bm = bm
        self.metrics = g("init_metrics").copy()
        self.k_cache = g("init_k_cache").copy()
        self.v_cache = g("init_v_cache").copy()
        self.max_kv = int(g("max_kv_per_compression"))

    def get_slot_index(self, seq):
        return seq.seq_id

    def sequence_counts(self, slots):
        ctx = self.bm.context_lens[:, slots]
        return ctx.sum(axis=(0, 2)).tolist(), ((ctx + self.B - 1) // self.B).sum(axis=(0, 2)).tolist()

    def finite_completion_blocks(self, slots, last_positions, protected):
        B, out = self.B, []
        pos = self.bm.token_positions.reshape(-1)
        for s_idx, last, prot in zip(slots, last_positions, protected):
            fin = 0
            for l in range(self.L):
                for h in range(self.H):
                    n = int(self.bm.context_lens[l, s_idx, h])
                    idx = np.arange(n)
                    sl = self.bm.block_tables[l, s_idx, h, idx // B].astype(np.int64) * B + idx % B
                    e = int((pos[sl] <= last - prot).sum())
                    r0 = n % B or B
                    fin += min((e - r0) // B + 1 if e >= r0 else 0, (n + B - 1) // B)
            out.append(fin)
        return out

    def compress(self, slots, last_positions, budgets, protected):
        from vllm_kvcompress_b200.scheduler import CacheMoves
        O, B = self.O, self.B
        ctx = np.ascontiguousarray(self.bm.context_lens[:, slots])
        bt = np.ascontiguousarray(self.bm.block_tables[:, slots])
        hanging = np.where(ctx % B == 0, B, ctx % B).transpose(1, 0, 2).astype(np.int32).copy()
        alloc = ((ctx + B - 1) // B * B).transpose(1, 0, 2).reshape(-1)
        offs = (np.cumsum(alloc) - alloc).astype(np.int32).reshape(hanging.shape)
        logical, kv_count, blk_count = O.select_evictions(
            self.metrics, self.bm.token_positions, bt, ctx, hanging, offs, np.asarray(budgets, np.int32),
            np.asarray(last_positions, np.int32), np.asarray(protected, np.int32), B, MAX_INT)
        moves = np.zeros((self.max_kv, 2), np.int32)
        cnt = np.zeros(hanging.shape, np.int32)
        O.schedule_cache_moves(moves, cnt, logical, kv_count, offs, bt, ctx, B)
        return CacheMoves(moves, cnt, offs), blk_count

    def free_compressed(self, slots, evicted_block_count):
        self.bm.free_compressed(slots, evicted_block_count)

    # ---- what the replay needs besides the scheduler interface ----
    def execute_moves(self, mv):
        kc = self.k_cache.view(np.uint16)
        vc = self.v_cache.view(np.uint16)
        self.O.execute_cache_moves(kc, vc, self.metrics, self.bm.token_positions, mv.index, mv.count, mv.offsets)

    def append(self, slots, positions):
        self.bm.append_slots(slots, positions)

    def state(self):
        return dict(block_tables=self.bm.block_tables, context_lens=self.bm.context_lens,
                    free_mask=self.bm.free_mask, metrics=self.metrics, token_positions=self.bm.token_positions,
                    k_cache=self.k_cache, v_cache=self.v_cache)

    def update(self, fn):
        """fn(metrics, k_cache, v_cache, block_tables, context_lens, positions) mutates the host arrays in place"""
        fn(self.metrics, self.k_cache, self.v_cache, self.bm.block_tables, self.bm.context_lens,
           self.bm.token_positions)

    def moves_host(self, mv):
        return mv.index, mv.count, mv.offsets


# --------------------------------------------------------------------------------------------------------------
# the replay
# --------------------------------------------------------------------------------------------------------------
def replay(g, backend, quirk=True):
    from vllm_kvcompress_b200.scheduler import CompressionScheduler, SchedulerConfig
    L, H, B, D, x = (int(g(k)) for k in ("L", "H", "B", "D", "x"))
    cfg = SchedulerConfig(num_layers=L, num_kv_heads=H, block_size=B,
                          max_kv_per_compression=int(g("max_kv_per_compression")),
                          compression_interval=int(g("compression_interval")), reference_v2_batch_quirk=quirk)
    sched = CompressionScheduler(cfg, backend)
    seqs, params = make_seqs(g)
    slots_all = [s.seq_id for s in seqs]
    compressions = 0
    for t in range(int(g("num_steps"))):
        out = sched.schedule_compression(seqs, params)
        tag = f"step {t}"
        assert int(g.step(t, "compressed")) == (out is not None), tag
        if out is not None:
            compressions += 1
            ids = sorted(out.freed_block_count.keys())
            assert ids == g.step(t, "seq_ids").tolist(), tag
            index, count, offsets = backend.moves_host(out.cache_moves)
            freed = np.stack([np.asarray(backend.to_host(out.freed_block_count[i])
                                         if hasattr(backend, "to_host") else out.freed_block_count[i])
                              for i in ids])
            np.testing.assert_array_equal(freed, g.step(t, "freed_block_count"), err_msg=tag)
            np.testing.assert_array_equal(count, g.step(t, "moves_count"), err_msg=tag)
            np.testing.assert_array_equal(offsets, g.step(t, "moves_offsets"), err_msg=tag)
            flat = [index[o:o + c] for o, c in zip(offsets.reshape(-1), count.reshape(-1))]
            got = np.concatenate(flat) if flat else np.zeros((0, 2), np.int32)
            np.testing.assert_array_equal(got, g.step(t, "moves"), err_msg=tag)
            backend.execute_moves(out.cache_moves)
        for s in seqs:
            s._len += 1
        backend.append(slots_all, [s._len - 1 for s in seqs])

        def model(metrics, k_cache, v_cache, block_tables, context_lens, positions):
            sl, keys, newest = live_slots(block_tables, context_lens, positions, slots_all, L, H, B)
            metrics.reshape(-1)[sl[newest]] = 0.0                  # kvcompress_reshape_and_cache: metric := bias
            write_payload(k_cache, v_cache, sl[newest], keys[newest], B, D, x)
            metrics.reshape(-1)[sl] += chain_noise(keys, t)          # aggregate_decode

        backend.update(model)
        st = backend.state()
        np.testing.assert_array_equal(st["context_lens"], g.step(t, "context_lens"), err_msg=tag)
        np.testing.assert_array_equal(st["free_mask"], g.step(t, "free_mask"), err_msg=tag)
        sl, _, _ = live_slots(st["block_tables"], st["context_lens"], st["token_positions"], slots_all, L, H, B)
        gsl, _, _ = live_slots(g.step(t, "block_tables"), g.step(t, "context_lens"), g.step(t, "token_positions"),
                               slots_all, L, H, B)
        np.testing.assert_array_equal(sl, gsl, err_msg=tag + " live block-table entries")
        for name in ("metrics", "token_positions"):
            np.testing.assert_array_equal(st[name].reshape(-1)[sl], g.step(t, name).reshape(-1)[sl],
                                          err_msg=f"{tag} {name}")
        blk, off = sl // B, sl % B
        np.testing.assert_array_equal(st["k_cache"][blk, :, off, :], g.step(t, "k_cache")[blk, :, off, :], err_msg=tag)
        np.testing.assert_array_equal(st["v_cache"][blk, :, off], g.step(t, "v_cache")[blk, :, off], err_msg=tag)
    assert compressions >= 6


def test_reference_v2_budgets_single_sequence_is_identity_up_to_finite_rows():
    from vllm_kvcompress_b200.scheduler import reference_v2_budgets
    assert reference_v2_budgets([10], [40], [25]) == [10]
    assert reference_v2_budgets([30], [40], [25]) == [25]         # cannot evict +inf rows
    # second sequence loses the first one's 15 inf rows; third one goes negative and cuts into nothing it owns
    assert reference_v2_budgets([10, 20, 5], [40, 50, 30], [25, 30, 30]) == [10, 5, 0]
    # a very negative U_i reaches back into the previous sequence's kept rows
    assert reference_v2_budgets([10, 0], [12, 30], [2, 30]) == [min(2, 12 - 10), 0]


def test_policy_matches_reference_formulas():
    """_schedule_seq_evictions against hand-computed values of scheduler.py:100-181"""
    from vllm_kvcompress_b200.scheduler import CompressionScheduler, SchedulerConfig
    s = CompressionScheduler(SchedulerConfig(num_layers=2, num_kv_heads=4, block_size=16), backend=None)
    seq = _Seq(0, 200)
    # max_cache_tokens 50 -> 64 per head -> 64 * 8 / 16 = 32 blocks kept
    assert s._schedule_seq_evictions(seq, 1600, 104, 1.0, 50, 16, False) == ((104 - 32) * 16, 104 - 32)
    assert seq.compressed
    assert s._schedule_seq_evictions(seq, 1600, 104, 1.0, 50, 16, True) == (0, 0)        # compress_once
    # rate 0.5, protected 32 -> 32: compressible (200 - 32) * 8 = 1344 -> target 672 + 256 = 928 -> evict 672 KVs
    assert s._schedule_seq_evictions(_Seq(1, 200), 1600, 104, 0.5, -1, 32, False) == (672, 42)
    with pytest.raises(RuntimeError):
        s._schedule_seq_evictions(_Seq(2, 200), 1600, 104, 0.5, 64, 32, False)


def test_scheduler_chain_vs_reference_scheduler_cpu_oracle():
    g = Gold()
    replay(g, OracleBackend(g))


def test_live_reference_batch_selection_differs_from_k1_without_the_quirk_budgets():
    """Pins the documented difference: with the budgets the policy computes (what the reference's CUDA kernel K1
    would be given), the multi-sequence calls of the live V2 selection free fewer blocks for every sequence but the
    first (metrics.py:718); scheduler.reference_v2_budgets is exactly that difference."""
    g = Gold()
    with pytest.raises(AssertionError):
        replay(g, OracleBackend(g), quirk=False)


# --------------------------------------------------------------------------------------------------------------
# GPU backend
# --------------------------------------------------------------------------------------------------------------
class GpuReplayBackend:
    """scheduler.DeviceBackend + the few extra hooks the replay needs (host copies of the state)"""

    def __init__(self, g):
        import torch
        from vllm_kvcompress_b200 import capi
        from vllm_kvcompress_b200.scheduler import DeviceBackend
        self.torch = torch
        dev = torch.device("cuda", 0)
        cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.B = int(g("B"))
        self.block_tables = cu(g("init_block_tables"))
        self.context_lens = cu(g("init_context_lens"))
        self.kv_position = cu(g("init_token_positions"))
        self.metrics = cu(g("init_metrics"))
        self.k_cache = cu(g("init_k_cache"))
        self.v_cache = cu(g("init_v_cache"))
        meta = tuple(cu(g("init_" + k)) for k in ("seq_index_by_block", "layer_index_by_block", "head_index_by_block",
                                                  "logical_block_num_by_block"))
        self.alloc = capi.DeviceBlockAllocator(int(g("NB")), self.block_tables, self.context_lens, self.kv_position,
                                               self.B, metadata=meta)
        fm = cu(g("init_free_mask"))
        self.alloc.free_mask.copy_(fm)
        self.alloc.free_count.fill_(int(fm.sum()))
        self.inner = DeviceBackend(self.alloc, self.metrics, self.kv_position, {i: i for i in range(int(g("MS")))},
                                   int(g("max_kv_per_compression")))
        self.L, self.H = int(g("L")), int(g("H"))

    def __getattr__(self, name):      # the scheduler interface
        return getattr(self.inner, name)

    def to_host(self, t):
        return t.cpu().numpy()

    def moves_host(self, mv):
        return mv.index.cpu().numpy(), mv.count.cpu().numpy(), mv.offsets.cpu().numpy()

    def execute_moves(self, mv):
        from vllm_kvcompress_b200 import _custom_ops as ops
        ops.execute_cache_moves(self.k_cache, self.v_cache, self.metrics, self.kv_position, mv.index, mv.count,
                                mv.offsets, 1, 16)

    def append(self, slots, positions):
        t = self.torch
        dev = self.metrics.device
        out = t.empty((self.L, len(slots), self.H), dtype=t.int64, device=dev)
        self.alloc.append_slots(t.tensor(slots, dtype=t.int32, device=dev),
                                t.tensor(positions, dtype=t.int32, device=dev), out)

    def update(self, fn):
        h = [x.cpu().numpy() for x in (self.metrics, self.k_cache, self.v_cache, self.block_tables, self.context_lens,
                                       self.kv_position)]
        fn(*h)
        self.metrics.copy_(self.torch.from_numpy(h[0]))
        self.k_cache.copy_(self.torch.from_numpy(h[1]))
        self.v_cache.copy_(self.torch.from_numpy(h[2]))

    def state(self):
        n = lambda x: x.cpu().numpy()
        assert int(self.alloc.status.item()) == 0
        return dict(block_tables=n(self.block_tables), context_lens=n(self.context_lens),
                    free_mask=n(self.alloc.free_mask), metrics=n(self.metrics), token_positions=n(self.kv_position),
                    k_cache=n(self.k_cache), v_cache=n(self.v_cache))


@pytest.mark.gpu
def test_scheduler_chain_vs_reference_scheduler_cuda():
    g = Gold()
    replay(g, GpuReplayBackend(g))
