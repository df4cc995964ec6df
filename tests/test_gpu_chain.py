"""GPU: multi-step CHAIN parity of the decode hot path (the reference's own test is such a chain:
tests/kernels/test_kvcompress_eviction.py:387-1226 runs evictions -> moves -> execute on one state).

Three implementations of the same step start from the same logical cache and are stepped side by side:

  prod  vllm_kvcompress_b200.engine.DecodeHotPath in its PRODUCTION configuration: fused selection, device-side
        allocator, deferred append + deferred move, fused attention step, optionally as CUDA-graph replays
  ref   the same driver bound to the REFERENCE's own CUDA kernels (oracle/ref_ops.py over oracle/_ref/_kvc_ref.so =
        the fork's three .cu files compiled unmodified): torch.sort chain -> K1/K2 -> host re-sort -> K4 -> K5 ->
        per layer K6 + paged attention v1 with temp metrics -> aggregate_decode
  cpu   oracle/cpu_decode_step.CpuDecodeHotPath (C oracle + numpy sorts)

Every step: context lengths, eviction counts, freed-block counts and move counts bit-equal; the move lists equal
as (dst, src) LOGICAL indices (prod allocates blocks with the device allocator, ref/cpu use static slabs, so the
physical ids differ by design); attention outputs within the reference's own tolerance (atol 1e-3,
test_kvcompress_attention.py:356-357).  At checkpoints, after flush_deferred: K / V / positions of every live slot
bit-equal, metrics within float tolerance (three different softmax implementations feed them)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B = 16


def _logical(bt, ctx_max_blocks, t):
    """gather a per-block tensor through block tables [heads, MB] -> [heads, nblk, ...]"""
    return t[bt[:, :ctx_max_blocks].long()]


def _phys_to_logical(moves, counts, offs, block_tables_lsh, S, L, H):
    """moves [N,2] physical slot ids -> list per head (s,l,h order) of [cnt,2] logical indices"""
    out = []
    bt = block_tables_lsh  # [L,S,H,MB] numpy
    for s in range(S):
        for l in range(L):
            for h in range(H):
                c, o = int(counts[s, l, h]), int(offs[s, l, h])
                mv = moves[o:o + c].astype(np.int64)
                row = bt[l, s, h]
                # entries right of the live part are stale and may repeat a live id: first occurrence wins
                inv = {}
                for i, b in enumerate(row):
                    inv.setdefault(int(b), i)
                lg = np.array([[inv[int(d) // B] * B + int(d) % B, inv[int(sr) // B] * B + int(sr) % B]
                               for d, sr in mv], dtype=np.int64).reshape(-1, 2)
                out.append(lg)
    return out


def _make_cpu(cfg_kw, ref_eng):
    """CpuDecodeHotPath over a copy of ref_eng's initial state (static slabs, same physical ids)."""
    from oracle.cpu_decode_step import CpuDecodeHotPath
    from vllm_kvcompress_b200.engine import HotPathConfig
    ccfg = HotPathConfig(allocator="static", **cfg_kw)
    cpu = CpuDecodeHotPath(ccfg, seed=cfg_kw.get("seed", 0))
    st = ref_eng.st
    assert np.array_equal(cpu.block_tables, st.block_tables.cpu().numpy())
    n = lambda t: t.detach().cpu().numpy().copy()
    cpu.k_cache = st.k_cache.view(torch.int16).cpu().numpy().view(np.uint16).copy()
    cpu.v_cache = st.v_cache.view(torch.int16).cpu().numpy().view(np.uint16).copy()
    cpu.kv_metrics = n(st.kv_metrics)
    cpu.kv_position = n(st.kv_position)
    cpu.context_lens = n(st.context_lens)
    cpu.seq_len = n(ref_eng.seq_len)
    return cpu


CASES = [
    # (shape kwargs, dtype, graph replay for prod)
    (dict(num_seqs=3, num_layers=2, num_kv_heads=8, queries_per_kv=4, max_cache_tokens=64, blocks_per_head=8,
          protected_window=16, seed=21), torch.float16, False),
    (dict(num_seqs=3, num_layers=2, num_kv_heads=8, queries_per_kv=4, max_cache_tokens=64, blocks_per_head=8,
          protected_window=16, seed=22), torch.float16, True),
    # Llama-3-70B group size, bf16, longer per-head contexts (logits leave shared memory: Q * ctx * 4 > 3584 B)
    (dict(num_seqs=2, num_layers=3, num_kv_heads=8, queries_per_kv=8, max_cache_tokens=128, blocks_per_head=14,
          protected_window=32, seed=23), torch.bfloat16, False),
    # one-CTA-per-sequence selection + 32-heads-per-warp K4 / K5 launches (>= 96 sequences, >= 4096 heads)
    (dict(num_seqs=96, num_layers=6, num_kv_heads=8, queries_per_kv=4, max_cache_tokens=64, blocks_per_head=8,
          protected_window=16, seed=24), torch.float16, True),
]


@pytest.mark.parametrize("pdl", ["model", "safe", "independent"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_decode_chain_vs_reference_kernels_and_cpu_oracle(case, pdl):
    from oracle import ref_ops
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig
    if not ref_ops.available():
        pytest.skip("oracle/_ref/_kvc_ref.so not built")
    kw, dtype, graph = CASES[case]
    if pdl != "model" and case not in (0, 2):
        pytest.skip("launch-mode sweep runs on two shapes")
    big = kw["num_seqs"] >= 96
    steps = 12 if big else 32
    dev = torch.device("cuda", 0)
    prod = DecodeHotPath(HotPathConfig(dtype=dtype, attention_pdl=pdl, **kw), dev)       # production defaults
    assert prod.deferred and prod.pend_dst is not None and prod.alloc is not None and prod.cfg.selection == "fused"
    ref = DecodeHotPath(HotPathConfig(dtype=dtype, allocator="static", selection="chain", fused_metrics=False,
                                      fused_cache_write=False, deferred_append=False, **kw), dev,
                        ops_module=ref_ops.RefOps())
    cfg = prod.cfg
    S, L, H, Q, D = cfg.num_seqs, cfg.num_layers, cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size
    heads = L * S * H
    nb0 = cfg.max_cache_tokens // B
    bt_p = lambda: prod.st.block_tables.view(heads, -1)
    bt_r = lambda: ref.st.block_tables.view(heads, -1)
    # same LOGICAL contents in both layouts
    for name in ("k_cache", "v_cache", "kv_metrics", "kv_position"):
        getattr(prod.st, name)[bt_p()[:, :nb0].long()] = _logical(bt_r(), nb0, getattr(ref.st, name))
    use_cpu = dtype == torch.float16 and not big        # the C oracle step is fp16 and single-digit sequences
    cpu = _make_cpu(kw, ref) if use_cpu else None

    g = torch.Generator(device=dev).manual_seed(100 + case)
    nvar = 5
    qkv_all = (torch.rand((nvar, L, S, (H * Q + 2 * H) * D), generator=g, device=dev) - 0.5).to(dtype)
    qkv = qkv_all[0].clone()
    out_p = torch.zeros((L, S, H * Q * D), device=dev, dtype=dtype)
    out_r = torch.zeros_like(out_p)
    out_c = np.zeros((L, S, H * Q * D), np.float32)
    if graph:
        ts = (prod.st.k_cache, prod.st.v_cache, prod.st.kv_metrics, prod.st.kv_position, prod.st.context_lens,
              prod.st.block_tables, prod._pos2, prod.alloc.free_mask, prod.alloc.free_count, prod.pend_kv,
              prod.pend_slot, prod.pend_dst)
        saved = [t.clone() for t in ts]
        assert prod.capture(qkv, out_p)
        [t.copy_(v) for t, v in zip(ts, saved)]

    n = lambda t: t.detach().cpu().numpy()
    total_moves = 0
    for i in range(steps):
        qkv.copy_(qkv_all[i % nvar])
        ctx_before = n(ref.st.context_lens).copy()
        bt_before = n(prod.st.block_tables).copy()   # the step frees and re-allocates table entries
        if graph:
            prod.replay()
        else:
            prod.step(qkv, out_p)
        ref.step(qkv, out_r)
        torch.cuda.synchronize()
        tag = f"case {case} step {i}"
        # ---- schedules ----
        assert torch.equal(prod.st.context_lens, ref.st.context_lens), tag
        assert torch.equal(prod.evicted_kv_count, ref.evicted_kv_count), tag
        assert torch.equal(prod.evicted_block_count, ref.evicted_block_count), tag
        assert torch.equal(prod.moves_count, ref.moves_count), tag
        total_moves += int(ref.moves_count.sum())
        if not big or i % 4 == 3:
            alloc = ((ctx_before + B - 1) // B * B).transpose(1, 0, 2).reshape(-1)
            offs = (np.cumsum(alloc) - alloc).astype(np.int64).reshape(S, L, H)
            lp = _phys_to_logical(n(prod.cache_moves), n(prod.moves_count), offs, bt_before, S, L, H)
            lr = _phys_to_logical(n(ref.cache_moves), n(ref.moves_count), offs, n(ref.st.block_tables), S, L, H)
            for a_, b_ in zip(lp, lr):
                assert np.array_equal(a_, b_), tag
        # ---- attention outputs ----
        assert torch.allclose(out_p.float(), out_r.float(), atol=1e-3, rtol=1e-2), tag
        if cpu is not None:
            cpu.step(qkv.view(torch.int16).cpu().numpy().view(np.uint16), out_c)
            assert np.array_equal(cpu.context_lens, n(ref.st.context_lens)), tag
            assert np.allclose(out_c, n(out_r.float()), atol=1e-3, rtol=1e-2), tag
            assert np.allclose(out_c, n(out_p.float()), atol=1e-3, rtol=1e-2), tag
        # ---- cache contents at checkpoints ----
        if i % 8 == 7 or i == steps - 1:
            prod.flush_pending()
            torch.cuda.synchronize()
            nblk = int((int(ref.st.context_lens.max()) + B - 1) // B)
            ctx = ref.st.context_lens.view(-1)
            live = (torch.arange(nblk * B, device=dev)[None] < ctx[:, None]).view(heads, nblk, B)
            pos_p = _logical(bt_p(), nblk, prod.st.kv_position)[live]
            pos_r = _logical(bt_r(), nblk, ref.st.kv_position)[live]
            assert torch.equal(pos_p, pos_r), tag
            m_p = _logical(bt_p(), nblk, prod.st.kv_metrics)[live]
            m_r = _logical(bt_r(), nblk, ref.st.kv_metrics)[live]
            assert torch.allclose(m_p, m_r, rtol=2e-3, atol=1e-6), tag
            k_p = _logical(bt_p(), nblk, prod.st.k_cache).permute(0, 1, 3, 2, 4)[live]   # [tokens, D/x, x]
            k_r = _logical(bt_r(), nblk, ref.st.k_cache).permute(0, 1, 3, 2, 4)[live]
            assert torch.equal(k_p, k_r), tag
            v_p = _logical(bt_p(), nblk, prod.st.v_cache).permute(0, 1, 3, 2)[live]      # [tokens, D]
            v_r = _logical(bt_r(), nblk, ref.st.v_cache).permute(0, 1, 3, 2)[live]
            assert torch.equal(v_p, v_r), tag
            if cpu is not None:
                bt_c = torch.from_numpy(cpu.block_tables).view(heads, -1)
                livec = live.cpu()
                assert np.array_equal(_logical(bt_c, nblk, torch.from_numpy(cpu.kv_position))[livec].numpy(),
                                      n(pos_r)), tag
                mc = _logical(bt_c, nblk, torch.from_numpy(cpu.kv_metrics))[livec].numpy()
                assert np.allclose(mc, n(m_r), rtol=2e-3, atol=1e-6), tag
                kc = torch.from_numpy(cpu.k_cache.view(np.int16)).view(-1, D // 8, B, 8)
                assert np.array_equal(_logical(bt_c, nblk, kc).permute(0, 1, 3, 2, 4)[livec].numpy(),
                                      n(k_r.view(torch.int16))), tag
    prod.check_status()
    assert total_moves > 0


def test_full_scale_step_vs_reference_kernels():
    """BASELINE config 2 at FULL scale (256 sequences x 32 layers x 8 kv heads = 65 536 heads, 9.4 M slots per
    compression call): four production steps against the reference's own CUDA kernels on the same logical cache.
    Eviction / freed-block / move counts and context lengths bit-equal for every head, the move lists of a sample of
    heads equal as logical pairs, outputs within 1e-3."""
    from oracle import ref_ops
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig
    if not ref_ops.available():
        pytest.skip("oracle/_ref/_kvc_ref.so not built")
    dev = torch.device("cuda", 0)
    kw = dict(num_seqs=256, num_layers=32, seed=31)       # HotPathConfig defaults = the benchmark's shape
    prod = DecodeHotPath(HotPathConfig(**kw), dev)
    ref = DecodeHotPath(HotPathConfig(allocator="static", selection="chain", fused_metrics=False,
                                      fused_cache_write=False, deferred_append=False, **kw), dev,
                        ops_module=ref_ops.RefOps())
    cfg = prod.cfg
    S, L, H, Q, D = cfg.num_seqs, cfg.num_layers, cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size
    heads = L * S * H
    nb0 = cfg.max_cache_tokens // B
    bt_p = prod.st.block_tables.view(heads, -1)
    bt_r = ref.st.block_tables.view(heads, -1)
    for name in ("k_cache", "v_cache", "kv_metrics", "kv_position"):
        for h0 in range(0, heads, 8192):
            getattr(prod.st, name)[bt_p[h0:h0 + 8192, :nb0].long()] = getattr(ref.st, name)[bt_r[h0:h0 + 8192, :nb0].long()]
    g = torch.Generator(device=dev).manual_seed(77)
    qkv = (torch.rand((L, S, (H * Q + 2 * H) * D), generator=g, device=dev) - 0.5).to(cfg.dtype)
    out_p = torch.zeros((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    out_r = torch.zeros_like(out_p)
    n = lambda t: t.detach().cpu().numpy()
    rng = np.random.RandomState(5)
    total_moves = 0
    for i in range(4):    # step 0 has nothing to evict (every head exactly at max_cache_tokens)
        ctx_before = n(ref.st.context_lens).copy()
        bt_before = n(prod.st.block_tables).copy()
        prod.step(qkv, out_p)
        ref.step(qkv, out_r)
        torch.cuda.synchronize()
        tag = f"full scale step {i}"
        assert torch.equal(prod.st.context_lens, ref.st.context_lens), tag
        assert torch.equal(prod.evicted_kv_count, ref.evicted_kv_count), tag
        assert torch.equal(prod.evicted_block_count, ref.evicted_block_count), tag
        assert torch.equal(prod.moves_count, ref.moves_count), tag
        total_moves += int(ref.moves_count.sum())
        assert torch.allclose(out_p.float(), out_r.float(), atol=1e-3, rtol=1e-2), tag
        # move lists of 300 random heads, as logical (dst, src) pairs
        alloc = ((ctx_before + B - 1) // B * B).transpose(1, 0, 2).reshape(-1)
        offs = (np.cumsum(alloc) - alloc).astype(np.int64).reshape(S, L, H)
        mp, mr = n(prod.cache_moves), n(ref.cache_moves)
        cnt = n(ref.moves_count)
        bt_ref = n(ref.st.block_tables)
        for _ in range(300):
            s_, l_, h_ = rng.randint(S), rng.randint(L), rng.randint(H)
            c, o = int(cnt[s_, l_, h_]), int(offs[s_, l_, h_])
            def logical(mv, row):
                inv = {}
                for j, b in enumerate(row):
                    inv.setdefault(int(b), j)
                return [(inv[int(d) // B] * B + int(d) % B, inv[int(sr) // B] * B + int(sr) % B) for d, sr in mv]
            assert logical(mp[o:o + c], bt_before[l_, s_, h_]) == logical(mr[o:o + c], bt_ref[l_, s_, h_]), tag
    assert total_moves > 100000
    prod.check_status()
