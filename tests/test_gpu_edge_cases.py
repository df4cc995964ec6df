"""GPU edge cases the reference's tests exercise implicitly: empty batches, heads with no work,
zero-length contexts, schedules with no moves."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
MAX_INT = 2147483000


@pytest.fixture(scope="module")
def capi():
    from vllm_kvcompress_b200 import capi as c
    return c


def i32(*shape, fill=0):
    return torch.full(shape, fill, dtype=torch.int32, device="cuda")


def test_empty_batch_is_a_noop(capi):
    z = i32(0)
    z3 = torch.zeros((0, 2, 2), dtype=torch.int32, device="cuda")
    capi.count_block_evictions(z3, i32(4), z3, z3, 16, MAX_INT)
    capi.truncate_cache_evictions(i32(4), i32(4), z3, z3, z3, 16, 4, MAX_INT)
    capi.schedule_t1_cache_moves(torch.zeros((4, 2), dtype=torch.int32, device="cuda"), z3, i32(4), z3, z3,
                                 torch.zeros((2, 0, 2, 3), dtype=torch.int32, device="cuda"),
                                 torch.zeros((2, 0, 2), dtype=torch.int32, device="cuda"), 16)
    torch.cuda.synchronize()


def test_execute_moves_with_zero_counts_leaves_cache_untouched(capi):
    g = torch.Generator().manual_seed(0)
    NB, S, L, H = 8, 2, 2, 2
    kc = torch.rand((NB, 16, 16, 8), generator=g).half().cuda()
    vc = torch.rand((NB, 128, 16), generator=g).half().cuda()
    m = torch.rand((NB, 16), generator=g).cuda()
    pos = torch.randint(0, 99, (NB, 16), generator=g, dtype=torch.int32).cuda()
    ref = [t.clone() for t in (kc, vc, m, pos)]
    moves = torch.randint(0, NB * 16, (64, 2), generator=g, dtype=torch.int32).cuda()
    capi.execute_cache_moves(kc, vc, m, pos, moves, i32(S, L, H), i32(S, L, H))
    torch.cuda.synchronize()
    for a, b in zip(ref, (kc, vc, m, pos)):
        assert torch.equal(a, b)


def test_single_partial_run_moves_only_owned_bytes(capi):
    """one move into a block: exactly one token column of the destination block changes"""
    g = torch.Generator().manual_seed(1)
    NB = 4
    kc = torch.rand((NB, 16, 16, 8), generator=g).half().cuda()
    vc = torch.rand((NB, 128, 16), generator=g).half().cuda()
    m = torch.rand((NB, 16), generator=g).cuda()
    pos = torch.arange(NB * 16, dtype=torch.int32).view(NB, 16).cuda()
    k0, v0 = kc.clone(), vc.clone()
    moves = torch.tensor([[2 * 16 + 5, 3 * 16 + 9]], dtype=torch.int32, device="cuda")  # dst (blk 2, off 5) <- src (blk 3, off 9)
    capi.execute_cache_moves(kc, vc, m, pos, moves, i32(1, 1, 1, fill=1), i32(1, 1, 1))
    torch.cuda.synchronize()
    k0[2, :, 5, :] = k0[3, :, 9, :]
    v0[2, :, 5] = v0[3, :, 9]
    assert torch.equal(kc, k0) and torch.equal(vc, v0)
    assert int(pos[2, 5]) == 3 * 16 + 9


@pytest.mark.parametrize("version", ["v1", "fused"])
def test_attention_zero_context_head_outputs_zeros(capi, version):
    g = torch.Generator().manual_seed(2)
    S, Hq, H, D, B, NB = 2, 8, 2, 128, 16, 6
    q = (torch.rand((S, Hq, D), generator=g) - 0.5).half().cuda()
    kc = (torch.rand((NB, 16, B, 8), generator=g) - 0.5).half().cuda()
    vc = (torch.rand((NB, D, B), generator=g) - 0.5).half().cuda()
    bt = torch.tensor([[[0, 1], [2, 0]], [[3, 0], [4, 5]]], dtype=torch.int32, device="cuda")
    cl = torch.tensor([[20, 0], [0, 17]], dtype=torch.int32, device="cuda")   # two empty heads
    out = torch.full((S, Hq, D), 7.0, dtype=torch.float16, device="cuda")
    kp = torch.zeros((NB, B), dtype=torch.int32, device="cuda")
    z = torch.zeros(S, dtype=torch.int32, device="cuda")
    if version == "v1":
        met = torch.full((NB, B, 4), -1.0, device="cuda")
        capi.paged_attention_v1(out, met, q, kc, vc, H, D ** -0.5, bt, cl, kp, z, z, B, 20, None, True)
    else:
        met = torch.zeros((NB, B), device="cuda")
        capi.paged_attention_fused_metrics(out, met, q, kc, vc, H, D ** -0.5, bt, cl, kp, z, z, B, 20, None)
    torch.cuda.synchronize()
    assert (out[0, 4:] == 0).all() and (out[1, :4] == 0).all()       # empty heads -> zeros
    assert torch.isfinite(out.float()).all() and (out[0, :4] != 7.0).any()


def test_k4_evicting_everything_or_nothing(capi):
    """c = 0 -> no moves; c = ctx -> no moves (nothing survives)"""
    B, MB = 16, 4
    bt = torch.arange(2 * MB, dtype=torch.int32, device="cuda").view(1, 1, 2, MB)
    ctx = torch.tensor([[[40, 40]]], dtype=torch.int32, device="cuda")
    offs = torch.tensor([[[0, 48]]], dtype=torch.int32, device="cuda")
    logical = torch.full((96,), MAX_INT, dtype=torch.int32, device="cuda")
    logical[48:88] = torch.arange(40, dtype=torch.int32, device="cuda")
    cnt = torch.tensor([[[0, 40]]], dtype=torch.int32, device="cuda")
    moves = torch.full((96, 2), -1, dtype=torch.int32, device="cuda")
    mcnt = i32(1, 1, 2, fill=-1)
    capi.schedule_t1_cache_moves(moves, mcnt, logical, cnt, offs, bt, ctx, B)
    torch.cuda.synchronize()
    assert mcnt.flatten().tolist() == [0, 0] and (moves == -1).all()
