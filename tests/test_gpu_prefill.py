"""GPU parity of the tcgen05 prefill-metric kernels (kvc_prefill.cu) through the C ABI and the
torch op surface, against (1) golden vectors from the reference's _naive_kvc_attention /
aggregate_prefill and (2) the C oracle on larger seeded inputs.
Tolerance (written here, floating point): rtol 2e-4 / atol 1e-6 on the metrics -- fp32 softmax of
logits rounded to the input type like the reference; the differences are __expf and summation order."""
import numpy as np
import pytest
import torch

from tests.golden_util import load_cases

pytestmark = pytest.mark.gpu

PREFILL = load_cases("prefill_metrics.npz")
RTOL, ATOL = 2e-4, 1e-6


@pytest.fixture(scope="module")
def capi():
    from vllm_kvcompress_b200 import capi as c
    return c


@pytest.fixture(scope="module")
def O():
    from oracle import kvc_oracle
    kvc_oracle.build()
    return kvc_oracle


def cu16(a, is_bf16):
    t = torch.as_tensor(np.ascontiguousarray(a).view(np.int16)).cuda()
    return t.view(torch.bfloat16 if is_bf16 else torch.float16)


def starts_of(lens):
    return torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32).cuda()


def run_capi(capi, c, fused=False, want_out=True):
    bf = bool(c["is_bf16"])
    q, k = cu16(c["query"], bf), cu16(c["key"], bf)
    lens = c["seq_lens"].tolist()
    T, Hq = q.shape[0], q.shape[1]
    out = torch.full((T, Hq), -3.0, dtype=torch.float32, device="cuda") if want_out else None
    metrics = torch.as_tensor(c["kv_metrics_in"]).cuda().clone() if fused else None
    slots = torch.as_tensor(c["slot_mapping"]).cuda() if fused else None
    capi.prefill_kv_metrics(out, metrics, slots, q, k, starts_of(lens),
                            torch.as_tensor(c["kv_metric_buffer_len"]).cuda(), max(lens), float(c["scale"]),
                            int(c["n_observed"]), bool(c["use_l2"]), bool(c["use_average"]), bool(c["use_maxpool"]))
    torch.cuda.synchronize()
    return out, metrics


@pytest.mark.parametrize("name", sorted(PREFILL))
def test_prefill_metrics_golden(capi, name):
    c = PREFILL[name]
    out, _ = run_capi(capi, c)
    np.testing.assert_allclose(out.cpu().numpy(), c["out_kv_metric"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", sorted(PREFILL))
def test_prefill_metrics_fused_aggregate_golden(capi, name):
    c = PREFILL[name]
    out, metrics = run_capi(capi, c, fused=True)
    np.testing.assert_allclose(metrics.cpu().numpy(), c["out_kv_metrics"], rtol=RTOL, atol=ATOL)
    # aggregate-only mode (no [T, Hq] output) gives the same cache metrics
    _, metrics2 = run_capi(capi, c, fused=True, want_out=False)
    if int(c["n_observed"]) <= 32:
        assert torch.equal(metrics, metrics2)
    else:  # several warps add into one query head's column sums with atomics: order-dependent last bits
        torch.testing.assert_close(metrics, metrics2, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", sorted(PREFILL))
def test_aggregate_prefill_golden(capi, name):
    c = PREFILL[name]
    m = torch.as_tensor(c["kv_metrics_in"]).cuda().clone()
    capi.aggregate_prefill(m, torch.as_tensor(c["out_kv_metric"]).cuda(), torch.as_tensor(c["slot_mapping"]).cuda())
    np.testing.assert_allclose(m.cpu().numpy(), c["out_kv_metrics"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("cfg", [
    # seed, prompt lens, Hq, H, bf16, n_observed, buffer, l2, avg, pool
    (1, [6000], 4, 1, False, 32, [0], True, False, True),          # BASELINE prompt length, 12 key splits
    (2, [2049, 511, 3], 8, 2, True, 32, [5, 0, 1], True, False, True),
    (3, [777, 1300], 16, 2, False, 48, [0, 100], True, True, True),  # Q=8, 64-row heads, atomics
    (4, [4097], 2, 2, False, 128, [0], False, False, False),        # one query head per tile
    (5, [1500, 300], 16, 2, True, 500, [0, 20], True, False, True),   # 500 observed queries: 4 row chunks per head, Q=8
    (6, [900], 4, 4, False, 4096, [0], True, True, True),             # the reference's whole observation chunk
])
def test_prefill_metrics_random_vs_oracle(capi, O, cfg):
    seed, lens, Hq, H, bf, n_obs, buf, l2, avg, pool = cfg
    g = torch.Generator().manual_seed(seed)
    T, D = sum(lens), 128
    dt = torch.bfloat16 if bf else torch.float16
    q = torch.randn((T, Hq, D), generator=g).to(dt)
    k = torch.randn((T, H, D), generator=g).to(dt)
    bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)
    want = O.prefill_kv_metrics(bits(q), bits(k), lens, D ** -0.5, np.array(buf, np.int32), n_obs, 4096, l2, avg,
                                pool, bf)
    out = torch.empty((T, Hq), dtype=torch.float32, device="cuda")
    capi.prefill_kv_metrics(out, None, None, q.cuda(), k.cuda(), starts_of(lens),
                            torch.tensor(buf, dtype=torch.int32).cuda(), max(lens), D ** -0.5, n_obs, l2, avg, pool)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    # a logit that lands on a rounding tie of the input type can round the other way when the fp32
    # dot product is summed in a different order; allow that on a handful of entries
    close = np.isclose(got, want, rtol=RTOL, atol=ATOL)
    assert close.mean() > 0.999, f"{(~close).sum()} of {close.size} entries differ"
    np.testing.assert_allclose(got, want, rtol=2e-2, atol=1e-5)


@pytest.mark.gpu
def test_prefill_metrics_negative_scale_vs_oracle(capi, O):
    """A negative softmax scale turns the row maximum of the logits into the minimum of the accumulators: the
    kernel tracks the extremum in accumulator units (the scale is folded into the exponent constant)."""
    g = torch.Generator().manual_seed(11)
    lens, Hq, H, D, n_obs = [700, 45], 8, 2, 128, 32
    T = sum(lens)
    q = torch.randn((T, Hq, D), generator=g).to(torch.float16)
    k = torch.randn((T, H, D), generator=g).to(torch.float16)
    bits = lambda t: t.view(torch.int16).numpy().view(np.uint16)
    buf = np.array([0, 3], np.int32)
    want = O.prefill_kv_metrics(bits(q), bits(k), lens, -0.05, buf, n_obs, 4096, True, False, True, False)
    out = torch.empty((T, Hq), dtype=torch.float32, device="cuda")
    capi.prefill_kv_metrics(out, None, None, q.cuda(), k.cuda(), starts_of(lens), torch.as_tensor(buf).cuda(),
                            max(lens), -0.05, n_obs, True, False, True)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    close = np.isclose(got, want, rtol=RTOL, atol=ATOL)
    assert close.mean() > 0.999, f"{(~close).sum()} of {close.size} entries differ"
    np.testing.assert_allclose(got, want, rtol=2e-2, atol=1e-5)


def test_prefill_torch_op_surface_and_errors(capi):
    from vllm_kvcompress_b200 import _custom_ops as ops
    c = PREFILL["prefill0"]
    q, k = cu16(c["query"], False), cu16(c["key"], False)
    lens = c["seq_lens"].tolist()
    buf = torch.as_tensor(c["kv_metric_buffer_len"]).cuda()
    out = ops.prefill_kv_metrics(q, k, lens, float(c["scale"]), buf, n_observed=32)
    np.testing.assert_allclose(out.cpu().numpy(), c["out_kv_metric"], rtol=RTOL, atol=ATOL)
    m = torch.as_tensor(c["kv_metrics_in"]).cuda().clone()
    ops.aggregate_prefill(m, out, torch.as_tensor(c["slot_mapping"]).cuda())
    np.testing.assert_allclose(m.cpu().numpy(), c["out_kv_metrics"], rtol=RTOL, atol=ATOL)
    # strided query (a slice of a fused qkv projection) is accepted: token stride != Hq * D
    qkv = torch.zeros((q.shape[0], 3 * q.shape[1] * 128), dtype=q.dtype, device="cuda")
    qkv[:, : q.shape[1] * 128] = q.reshape(q.shape[0], -1)
    out2 = ops.prefill_kv_metrics(qkv[:, : q.shape[1] * 128].view(q.shape[0], q.shape[1], 128), k, lens,
                                  float(c["scale"]), buf, n_observed=32)
    assert torch.equal(out, out2)
    # query AND key as views of one fused projection output (the key's base pointer and token stride both differ
    # from a dense [T, H, D] tensor: the TMA tensor maps are built from the strides)
    Hq, H = q.shape[1], k.shape[1]
    fused = torch.randn((q.shape[0], (Hq + 2 * H) * 128), device="cuda").to(q.dtype)
    fused[:, : Hq * 128] = q.reshape(q.shape[0], -1)
    fused[:, Hq * 128: (Hq + H) * 128] = k.reshape(k.shape[0], -1)
    out3 = ops.prefill_kv_metrics(fused[:, : Hq * 128].view(-1, Hq, 128),
                                  fused[:, Hq * 128: (Hq + H) * 128].view(-1, H, 128), lens, float(c["scale"]), buf,
                                  n_observed=32)
    assert torch.equal(out, out3)
    with pytest.raises(RuntimeError, match="Unsupported head size"):
        ops.prefill_kv_metrics(q[:, :, :64].contiguous(), k[:, :, :64].contiguous(), lens, 1.0, buf)
    with pytest.raises(RuntimeError, match="Unsupported prefill_observed_queries"):
        ops.prefill_kv_metrics(q, k, lens, 1.0, buf, n_observed=5000)
    # empty prefill batch is a no-op
    e = ops.prefill_kv_metrics(q[:0], k[:0], [], 1.0, buf[:0])
    assert e.shape == (0, q.shape[1])
