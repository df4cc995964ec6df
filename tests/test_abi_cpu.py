"""CPU: the C-ABI library loads and exports every symbol include/kvc_b200.h declares, the
torch extension registers the reference's operator schemas, and argument validation that
needs no device behaves like the reference (RuntimeError / status codes).  No kernels run."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from vllm_kvcompress_b200 import native
    return native


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kvc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"KVC_API\s+[\w\s\*]+?\b(kvc_\w+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    names = declared_symbols()
    assert len(names) >= 13, names
    lib = built.load_lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kvc_b200.h but not exported"
    assert lib.kvc_abi_version() == 1
    assert lib.kvc_arch() == b"sm_100a"


def test_no_undeclared_exports(built):
    """the only dynamic symbols of the library are the declared C ABI"""
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", built.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_symbols()


def test_workspace_query_and_status_codes(built):
    lib = built.load_lib()
    n = lib.kvc_schedule_cache_evictions_workspace_bytes(3, 32, 8, 288000, 16)
    assert n > 2 * (288000 * 16 // 1024) * 256 * 4
    # unsupported block size is rejected before any device work (reference :474-490)
    rc = lib.kvc_count_block_evictions(None, None, None, None, 1, 1, 1, 0, 8, 0, None)
    assert rc == -1 and b"Unsupported block size: 8" in lib.kvc_last_error()
    rc = lib.kvc_count_block_evictions(None, None, None, None, 1, 1, 1, 0, 16, 0, None)
    assert rc == -2  # null tensors
    rc = lib.kvc_execute_cache_moves(None, None, None, None, None, None, None, 1, 1, 1,
                                     ctypes.c_int64(1), 64, 8, 16, 0, 1, 16, None)
    assert rc == -1 and b"Unsupported head size: 64" in lib.kvc_last_error()


# schemas of the reference extension: csrc/torch_bindings.cpp:53-80, 354-362, 374-417
REFERENCE_SCHEMAS = {
    ("_C", "kvcompress_paged_attention_v1"):
        "(Tensor! out, Tensor! kv_metric_out, Tensor query, Tensor key_cache, Tensor value_cache, "
        "int num_kv_heads, float scale, Tensor block_tables, Tensor context_lens, Tensor kv_position, "
        "Tensor last_position, Tensor kv_metric_buffer_len, int block_size, int max_context_len, "
        "Tensor? alibi_slopes, str kv_cache_dtype, float k_scale, float v_scale, bool record_kv_metrics) -> ()",
    ("_C", "kvcompress_paged_attention_v2"):
        "(Tensor! out, Tensor! kv_metric_out, Tensor exp_sums, Tensor max_logits, Tensor tmp_out, "
        "Tensor tmp_kv_metric_out, Tensor query, Tensor key_cache, Tensor value_cache, int num_kv_heads, "
        "float scale, Tensor block_tables, Tensor context_lens, Tensor kv_position, Tensor last_position, "
        "Tensor kv_metric_buffer_len, int block_size, int max_context_len, Tensor? alibi_slopes, "
        "str kv_cache_dtype, float k_scale, float v_scale, bool record_kv_metrics) -> ()",
    ("_C_cache_ops", "kvcompress_reshape_and_cache"):
        "(Tensor key, Tensor value, Tensor! key_cache, Tensor! value_cache, Tensor kv_metrics, "
        "Tensor slot_mapping, Tensor kv_metric_head_bias, str kv_cache_dtype, float k_scale, float v_scale) -> ()",
    ("_C_kvc_ops", "schedule_cache_evictions"):
        "(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices, Tensor! evicted_kv_count, "
        "Tensor! remaining_kv_count, Tensor evicted_kv_offsets, Tensor sorted_indices, "
        "Tensor seq_block_offsets, Tensor layer_by_block, Tensor head_by_block, "
        "Tensor virtual_block_num_by_block, Tensor evicted_blocks_per_seq, Tensor context_lens, "
        "Tensor hanging_token_count, Tensor kv_position, Tensor last_position, Tensor protected_window_size, "
        "int block_size, bool evict_evenly_per_layer, Tensor? control_layers, int max_evicted_kv, "
        "int null_eviction_index, bool truncate) -> ()",
    ("_C_kvc_ops", "truncate_cache_evictions"):
        "(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices, Tensor! evicted_kv_count, "
        "Tensor evicted_kv_offsets, Tensor hanging_token_count, int block_size, int max_evicted_kv, "
        "int null_eviction_index) -> ()",
    ("_C_kvc_ops", "count_block_evictions"):
        "(Tensor! evicted_block_count, Tensor! evicted_logical_indices, Tensor evicted_kv_offsets, "
        "Tensor hanging_token_count, int block_size, int null_value) -> ()",
    ("_C_kvc_ops", "schedule_t1_cache_moves"):
        "(Tensor! cache_moves_idx, Tensor! cache_moves_count, Tensor evicted_logical_indices, "
        "Tensor evicted_kv_count, Tensor evicted_kv_offsets, Tensor block_tables, Tensor context_lens, "
        "int block_size) -> ()",
    ("_C_kvc_ops", "execute_cache_moves"):
        "(Tensor! k_cache, Tensor! v_cache, Tensor! kv_metrics, Tensor! kv_position, "
        "Tensor cache_moves_idx, Tensor cache_moves_count, Tensor evicted_kv_offsets, int blocks_per_head, "
        "int threads_per_head) -> ()",
}


@pytest.mark.parametrize("key", sorted(REFERENCE_SCHEMAS))
def test_torch_op_schemas_match_reference(built, key):
    built.load_torch_ops()
    ns, name = key
    op = getattr(getattr(torch.ops, ns), name).default
    want = torch._C.parse_schema(f"{ns}::{name}{REFERENCE_SCHEMAS[key]}")
    got = op._schema

    def sig(schema):
        return [(a.name, str(a.type), bool(a.alias_info and a.alias_info.is_write)) for a in schema.arguments]

    assert sig(got) == sig(want)
    assert len(got.returns) == 0


def test_misnamespaced_eviction_op_also_resolves(built):
    """vllm/_custom_ops.py:982 calls torch.ops._C.schedule_cache_evictions"""
    built.load_torch_ops()
    assert str(torch.ops._C.schedule_cache_evictions.default._schema).startswith("_C::schedule_cache_evictions(")


def test_ops_have_no_cpu_fallback(built):
    built.load_torch_ops()
    t = torch.zeros((1, 1, 1), dtype=torch.int32)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops._C_kvc_ops.count_block_evictions(t, t.flatten(), t, t, 16, 0)


def test_python_shim_mirrors_reference_names(built):
    from vllm_kvcompress_b200 import _custom_ops as ops
    import inspect
    want = {  # name -> positional parameter names of the reference wrapper (vllm/_custom_ops.py)
        "schedule_cache_evictions": ["out_evicted_kv_indices", "out_evicted_logical_indices", "out_evicted_kv_count",
                                     "out_evicted_kv_offsets", "sorted_indices", "seq_block_offsets", "layer_by_block",
                                     "head_by_block", "logical_block_num_by_block", "evicted_blocks_per_seq",
                                     "context_lens", "hanging_token_count", "kv_position", "last_position",
                                     "protected_window_size", "block_size", "max_evicted_kv", "null_eviction_index",
                                     "truncate", "evict_evenly_per_layer", "control_layers"],
        "count_block_evictions": ["evicted_block_count", "evicted_logical_indices", "evicted_kv_offsets",
                                  "hanging_token_count", "block_size", "null_value", "evicted_blocks_per_seq"],
        "schedule_cache_moves": ["out_cache_moves_indices", "out_cache_moves_count", "evicted_logical_indices",
                                 "evicted_kv_count", "evicted_kv_offsets", "block_tables", "context_lens", "block_size"],
        "execute_cache_moves": ["k_cache", "v_cache", "kv_metrics", "kv_position", "cache_moves_indices",
                                "cache_moves_count", "evicted_kv_offsets", "blocks_per_head", "threads_per_head"],
        "reshape_and_cache_kvc": ["key", "value", "key_cache", "value_cache", "kv_metrics", "slot_mapping",
                                  "kv_metric_head_bias", "kv_cache_dtype", "k_scale", "v_scale"],
    }
    for name, params in want.items():
        assert list(inspect.signature(getattr(ops, name)).parameters) == params, name
    assert ops.MAX_INT == 2147483000
    with pytest.raises(RuntimeError, match="cannot evict evenly"):
        t = torch.zeros(1, dtype=torch.int32)
        ops.schedule_cache_evictions(t, t, t, t, t, t, t, t, t, t, t, t, t, t, t, 16, 1, 0, True,
                                     evict_evenly_per_layer=True)


ADDITIVE_OPS = {
    # op -> (namespace, arguments that the Python shim passes positionally, in order)
    "kvcompress_paged_attention_fused_step_deferred": ("_C", [
        "out", "kv_metrics", "query", "key", "value", "key_cache", "value_cache", "slot_mapping",
        "kv_metric_head_bias", "pend_kv", "pend_slot", "pend_dst", "num_kv_heads", "scale", "block_tables",
        "context_lens", "kv_position", "last_position", "kv_metric_buffer_len", "block_size", "max_context_len",
        "alibi_slopes", "kv_cache_dtype"]),
    "execute_cache_moves_deferred": ("_C_kvc_ops", [
        "k_cache", "v_cache", "kv_metrics", "kv_position", "cache_moves_idx", "cache_moves_count",
        "evicted_kv_offsets", "pend_kv", "pend_slot", "pend_dst"]),
    "flush_deferred": ("_C_kvc_ops", ["k_cache", "v_cache", "pend_kv", "pend_slot", "pend_dst"]),
}


@pytest.mark.parametrize("name", sorted(ADDITIVE_OPS))
def test_additive_deferred_ops_are_registered(built, name):
    """The deferred append / move ops are additive (new names, the reference schemas stay untouched) and the
    shim's positional call order matches the dispatcher's argument list."""
    built.load_torch_ops()
    ns, args = ADDITIVE_OPS[name]
    schema = getattr(getattr(torch.ops, ns), name).default._schema
    assert [a.name for a in schema.arguments] == args
    assert len(schema.returns) == 0
    assert (ns, name) not in REFERENCE_SCHEMAS


def test_deferred_ops_fail_loudly_on_cpu(built):
    built.load_torch_ops()
    k = torch.zeros((1, 16, 16, 8), dtype=torch.float16)
    v = torch.zeros((1, 128, 16), dtype=torch.float16)
    pend = torch.zeros((1, 256), dtype=torch.float16)
    slot = torch.full((1,), -1, dtype=torch.int64)
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops._C_kvc_ops.flush_deferred(k, v, pend, slot, None)


def test_select_mode_switch_round_trips(built):
    lib = built.load_lib()
    assert lib.kvc_get_select_mode() == 0
    assert lib.kvc_set_select_mode(3) == 0 and lib.kvc_get_select_mode() == 3
    assert lib.kvc_set_select_mode(2) == 0 and lib.kvc_get_select_mode() == 2
    assert lib.kvc_set_select_mode(7) != 0 and lib.kvc_get_select_mode() == 2   # rejected, unchanged
    assert lib.kvc_set_select_mode(0) == 0


def test_moves_mode_switch_round_trips(built):
    lib = built.load_lib()
    assert lib.kvc_get_moves_mode() == 0
    assert lib.kvc_set_moves_mode(1) == 0 and lib.kvc_get_moves_mode() == 1
    assert lib.kvc_set_moves_mode(5) != 0 and lib.kvc_get_moves_mode() == 1   # rejected, unchanged
    assert lib.kvc_set_moves_mode(0) == 0


def test_attention_pdl_modes_round_trip(built):
    lib = built.load_lib()
    keep = lib.kvc_get_attention_pdl()
    for m in (0, 1, 2, 3):
        assert lib.kvc_set_attention_pdl(m) == 0 and lib.kvc_get_attention_pdl() == m
    assert lib.kvc_set_attention_pdl(4) != 0 and lib.kvc_get_attention_pdl() == 3
    assert lib.kvc_set_attention_pdl(keep) == 0
