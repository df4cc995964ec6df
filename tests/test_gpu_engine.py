"""GPU: the decode hot-path driver (engine.DecodeHotPath) end to end.  The device-side allocator
(blocks freed and re-allocated every step, block tables rewritten in place) must be invisible to the
results: starting from the same LOGICAL cache contents, a replica on the device allocator and a replica
on static per-head slabs produce bit-identical attention outputs, context lengths and per-slot metrics
step after step, also through the CUDA-graph replay."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def pend(e):
    if not e.deferred:
        return ()
    return (e.pend_kv, e.pend_slot) + ((e.pend_dst,) if e.pend_dst is not None else ())


def logical_view(eng, t, nblk):
    """[heads, nblk, ...] gather of a per-block tensor through the engine's block tables"""
    bt = eng.st.block_tables.view(-1, eng.st.block_tables.shape[-1])[:, :nblk].long()
    return t[bt]


@pytest.mark.parametrize("graph", [False, True])
def test_device_allocator_is_invisible_to_results(graph):
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig
    dev = torch.device("cuda", 0)
    kw = dict(num_seqs=5, num_layers=3, max_cache_tokens=64, blocks_per_head=8, protected_window=16, seed=3)
    a = DecodeHotPath(HotPathConfig(allocator="static", **kw), dev)
    b = DecodeHotPath(HotPathConfig(allocator="device", **kw), dev)
    nb0 = kw["max_cache_tokens"] // 16
    bt_b = b.st.block_tables.view(-1, b.st.block_tables.shape[-1])[:, :nb0].long()
    for name in ("k_cache", "v_cache", "kv_metrics", "kv_position"):   # same logical contents in both layouts
        getattr(b.st, name)[bt_b] = logical_view(a, getattr(a.st, name), nb0)
    cfg = a.cfg
    H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
    g = torch.Generator(device=dev).manual_seed(5)
    steps = 40
    qkv_all = (torch.rand((steps, L, S, (H * Q + 2 * H) * D), generator=g, device=dev) - 0.5).to(cfg.dtype)
    qkv = qkv_all[0].clone()
    out_a = torch.zeros((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    out_b = torch.zeros_like(out_a)
    if graph:
        tensors = lambda e: (e.st.k_cache, e.st.v_cache, e.st.kv_metrics, e.st.kv_position, e.st.context_lens,
                             e.st.block_tables, e._pos2) + pend(e)
        state = lambda e: [t.clone() for t in tensors(e)]
        restore = lambda e, s: [t.copy_(v) for t, v in zip(tensors(e), s)]
        sa, sb = state(a), state(b)
        fm = b.alloc.free_mask.clone(), b.alloc.free_count.clone()
        assert a.capture(qkv, out_a) and b.capture(qkv, out_b)   # capture runs the step: rewind the state
        restore(a, sa), restore(b, sb)
        b.alloc.free_mask.copy_(fm[0]), b.alloc.free_count.copy_(fm[1])
    grew = 0
    for i in range(steps):
        qkv.copy_(qkv_all[i])
        if graph:
            a.replay(), b.replay()
        else:
            a.step(qkv, out_a), b.step(qkv, out_b)
        torch.cuda.synchronize()
        assert torch.equal(a.st.context_lens, b.st.context_lens), f"step {i}"
        assert torch.equal(out_a, out_b), f"step {i}"
        nblk = int((a.st.context_lens.max() + 15) // 16)
        ctx = a.st.context_lens.view(-1)
        live = (torch.arange(nblk * 16, device=dev)[None] < ctx[:, None]).view(-1, nblk, 16)
        for name in ("kv_metrics", "kv_position"):
            va = logical_view(a, getattr(a.st, name), nblk)
            vb = logical_view(b, getattr(b.st, name), nblk)
            assert torch.equal(va[live], vb[live]), f"{name} at step {i}"
        grew += int((b.st.context_lens.view(-1) > kw["max_cache_tokens"]).any())
    assert int(b.alloc.status.item()) == 0
    # allocator bookkeeping is consistent with the tables: every live block is marked used exactly once
    nblk_b = (b.st.context_lens.view(-1) + 15) // 16
    bt = b.st.block_tables.view(-1, b.st.block_tables.shape[-1])
    mask = torch.arange(bt.shape[1], device=dev)[None] < nblk_b[:, None]
    live_blocks = bt[mask].long()
    assert live_blocks.unique().numel() == live_blocks.numel()
    assert not b.alloc.free_mask[live_blocks].any()
    assert int(b.alloc.free_count.item()) == int(b.alloc.free_mask.sum()) == b.st.NB - live_blocks.numel()


@pytest.mark.parametrize("shape", [dict(num_seqs=2, num_layers=2), dict(num_seqs=5, num_layers=3),
                                   dict(num_seqs=96, num_layers=6)])
@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("deferred_move", [False, True])
def test_deferred_append_is_invisible_to_results(shape, graph, deferred_move):
    """Deferred append (new tokens that open a fresh block are parked in a compact buffer and stored by the
    next execute_cache_moves_deferred) against the plain fused step on the same allocator: identical
    attention outputs, tables and context lengths every step, and after flush_pending() bit-identical
    K / V / metrics / positions in every live slot.  Shapes cover gridDim.y > 1 (few heads: several warps
    share a head's move list) and the decode-time launches of the benchmark (>= 4096 heads: 32 heads per warp in
    K4 / K5; >= 96 sequences: the one-CTA-per-sequence selection)."""
    from vllm_kvcompress_b200.engine import DecodeHotPath, HotPathConfig
    dev = torch.device("cuda", 0)
    kw = dict(max_cache_tokens=64, blocks_per_head=8, protected_window=16, seed=11, **shape)
    a = DecodeHotPath(HotPathConfig(deferred_append=False, **kw), dev)
    b = DecodeHotPath(HotPathConfig(deferred_append=True, deferred_move=deferred_move, **kw), dev)
    assert b.deferred and not a.deferred and (b.pend_dst is not None) == deferred_move
    for name in ("k_cache", "v_cache", "kv_metrics", "kv_position"):
        assert torch.equal(getattr(a.st, name), getattr(b.st, name))
    cfg = a.cfg
    H, Q, D, L, S = cfg.num_kv_heads, cfg.queries_per_kv, cfg.head_size, cfg.num_layers, cfg.num_seqs
    g = torch.Generator(device=dev).manual_seed(6)
    steps = 36
    nvar = 4
    qkv_all = (torch.rand((nvar, L, S, (H * Q + 2 * H) * D), generator=g, device=dev) - 0.5).to(cfg.dtype)
    qkv = qkv_all[0].clone()
    out_a = torch.zeros((L, S, H * Q * D), device=dev, dtype=cfg.dtype)
    out_b = torch.zeros_like(out_a)
    if graph:
        def tensors(e):
            t = (e.st.k_cache, e.st.v_cache, e.st.kv_metrics, e.st.kv_position, e.st.context_lens,
                 e.st.block_tables, e._pos2, e.alloc.free_mask, e.alloc.free_count)
            return t + pend(e)
        sa, sb = [t.clone() for t in tensors(a)], [t.clone() for t in tensors(b)]
        assert a.capture(qkv, out_a) and b.capture(qkv, out_b)
        [t.copy_(v) for t, v in zip(tensors(a), sa)]
        [t.copy_(v) for t, v in zip(tensors(b), sb)]

    def live_equal(tag):
        nblk = (a.st.context_lens.view(-1) + 15) // 16
        bt = a.st.block_tables.view(-1, a.st.block_tables.shape[-1])
        mask = torch.arange(bt.shape[1], device=dev)[None] < nblk[:, None]
        blocks = bt[mask].long()
        tok = (torch.arange(bt.shape[1] * 16, device=dev)[None] < a.st.context_lens.view(-1)[:, None])
        tok = tok.view(-1, bt.shape[1], 16)[mask]                                    # [live blocks, 16]
        assert torch.equal(a.st.kv_metrics[blocks][tok], b.st.kv_metrics[blocks][tok]), tag
        assert torch.equal(a.st.kv_position[blocks][tok], b.st.kv_position[blocks][tok]), tag
        ka = a.st.k_cache[blocks].permute(0, 2, 1, 3)[tok]                           # [tokens, D/x, x]
        kb = b.st.k_cache[blocks].permute(0, 2, 1, 3)[tok]
        assert torch.equal(ka, kb), tag
        va = a.st.v_cache[blocks].permute(0, 2, 1)[tok]                              # [tokens, D]
        vb = b.st.v_cache[blocks].permute(0, 2, 1)[tok]
        assert torch.equal(va, vb), tag

    parked = 0
    for i in range(steps):
        qkv.copy_(qkv_all[i % nvar])
        if graph:
            a.replay(), b.replay()
        else:
            a.step(qkv, out_a), b.step(qkv, out_b)
        torch.cuda.synchronize()
        assert torch.equal(a.st.context_lens, b.st.context_lens), f"step {i}"
        assert torch.equal(a.st.block_tables, b.st.block_tables), f"step {i}"
        assert torch.equal(out_a, out_b), f"step {i}"
        parked += int((b.pend_slot >= 0).sum())
        if i % 7 == 6 or i == steps - 1:
            b.flush_pending()
            torch.cuda.synchronize()
            live_equal(f"after flush at step {i}")
    assert parked > 0
    # a flush right after a compression (deferred moves outstanding) must land the tokens at their destinations
    if deferred_move and not graph:
        a.step(qkv, out_a), b.step(qkv, out_b)
        a.compress(), b.compress()
        assert int((b.pend_dst >= 0).sum()) > 0
        b.flush_pending()
        torch.cuda.synchronize()
        assert int((b.pend_dst >= 0).sum()) == 0 and int((b.pend_slot >= 0).sum()) == 0
        # compare at the post-compression lengths (free_compressed has not run: shrink them by hand)
        for e in (a, b):
            ev = e.evicted_kv_count.permute(1, 0, 2)
            e.st.context_lens -= ev
        live_equal("flush between compression and attention")
