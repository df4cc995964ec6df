// TEST INFRASTRUCTURE ONLY (never imported by the product path).
//
// Registers the *reference's own* CUDA kernels (compiled unmodified from the
// sources where they lie under /root/reference/csrc, see Makefile.ref) under a
// separate torch library namespace `_kvc_ref`, so that GPU parity tests can run
// kernel-vs-kernel on identical device tensors in one process:
//   reference csrc/kvcompress_eviction_kernels.cu      (K1-K5)
//   reference csrc/kvcompress_cache_kernels.cu         (K6)
//   reference csrc/attention/kvcompress_attention_kernels.cu (K7/K8)
// Schemas restate csrc/torch_bindings.cpp:52-80, 353-362, 371-418.
#include <torch/all.h>
#include <torch/library.h>
#include <Python.h>

#include "kvcompress_eviction.h"  // reference header, found via -I/root/reference/csrc
#include "kvcompress_cache.h"     // reference header

// reference csrc/ops.h:18-37, 67-90 (declared here to avoid ops.h's unrelated includes)
void kvcompress_paged_attention_v1(
    torch::Tensor& out, torch::Tensor& kv_metric_out, torch::Tensor& query,
    torch::Tensor& key_cache, torch::Tensor& value_cache, int64_t num_kv_heads,
    double scale, torch::Tensor& block_tables, torch::Tensor& context_lens,
    torch::Tensor& kv_position, torch::Tensor& last_position,
    torch::Tensor& kv_metric_buffer_len, int64_t block_size,
    int64_t max_context_len, const c10::optional<torch::Tensor>& alibi_slopes,
    const std::string& kv_cache_dtype, double k_scale, double v_scale,
    bool record_kv_metrics);
void kvcompress_paged_attention_v2(
    torch::Tensor& out, torch::Tensor& kv_metric_out, torch::Tensor& exp_sums,
    torch::Tensor& max_logits, torch::Tensor& tmp_out,
    torch::Tensor& tmp_kv_metric_out, torch::Tensor& query,
    torch::Tensor& key_cache, torch::Tensor& value_cache, int64_t num_kv_heads,
    double scale, torch::Tensor& block_tables, torch::Tensor& context_lens,
    torch::Tensor& kv_position, torch::Tensor& last_position,
    torch::Tensor& kv_metric_buffer_len, int64_t block_size,
    int64_t max_context_len, const c10::optional<torch::Tensor>& alibi_slopes,
    const std::string& kv_cache_dtype, double k_scale, double v_scale,
    bool record_kv_metrics);

TORCH_LIBRARY(_kvc_ref, m) {
  m.def("kvcompress_paged_attention_v1(Tensor! out, Tensor! kv_metric_out, Tensor query,"
        " Tensor key_cache, Tensor value_cache, int num_kv_heads, float scale,"
        " Tensor block_tables, Tensor context_lens, Tensor kv_position, Tensor last_position,"
        " Tensor kv_metric_buffer_len, int block_size, int max_context_len, Tensor? alibi_slopes,"
        " str kv_cache_dtype, float k_scale, float v_scale, bool record_kv_metrics) -> ()");
  m.impl("kvcompress_paged_attention_v1", torch::kCUDA, &kvcompress_paged_attention_v1);
  m.def("kvcompress_paged_attention_v2(Tensor! out, Tensor! kv_metric_out, Tensor exp_sums,"
        " Tensor max_logits, Tensor tmp_out, Tensor tmp_kv_metric_out, Tensor query,"
        " Tensor key_cache, Tensor value_cache, int num_kv_heads, float scale,"
        " Tensor block_tables, Tensor context_lens, Tensor kv_position, Tensor last_position,"
        " Tensor kv_metric_buffer_len, int block_size, int max_context_len, Tensor? alibi_slopes,"
        " str kv_cache_dtype, float k_scale, float v_scale, bool record_kv_metrics) -> ()");
  m.impl("kvcompress_paged_attention_v2", torch::kCUDA, &kvcompress_paged_attention_v2);
  m.def("kvcompress_reshape_and_cache(Tensor key, Tensor value, Tensor! key_cache,"
        " Tensor! value_cache, Tensor kv_metrics, Tensor slot_mapping, Tensor kv_metric_head_bias,"
        " str kv_cache_dtype, float k_scale, float v_scale) -> ()");
  m.impl("kvcompress_reshape_and_cache", torch::kCUDA, &kvcompress_reshape_and_cache);
  m.def("schedule_cache_evictions(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices,"
        " Tensor! evicted_kv_count, Tensor! remaining_kv_count, Tensor evicted_kv_offsets,"
        " Tensor sorted_indices, Tensor seq_block_offsets, Tensor layer_by_block, Tensor head_by_block,"
        " Tensor virtual_block_num_by_block, Tensor evicted_blocks_per_seq, Tensor context_lens,"
        " Tensor hanging_token_count, Tensor kv_position, Tensor last_position,"
        " Tensor protected_window_size, int block_size, bool evict_evenly_per_layer,"
        " Tensor? control_layers, int max_evicted_kv, int null_eviction_index, bool truncate) -> ()");
  m.impl("schedule_cache_evictions", torch::kCUDA, &schedule_cache_evictions);
  m.def("truncate_cache_evictions(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices,"
        " Tensor! evicted_kv_count, Tensor evicted_kv_offsets, Tensor hanging_token_count,"
        " int block_size, int max_evicted_kv, int null_eviction_index) -> ()");
  m.impl("truncate_cache_evictions", torch::kCUDA, &truncate_cache_evictions);
  m.def("count_block_evictions(Tensor! evicted_block_count, Tensor! evicted_logical_indices,"
        " Tensor evicted_kv_offsets, Tensor hanging_token_count, int block_size, int null_value) -> ()");
  m.impl("count_block_evictions", torch::kCUDA, &count_block_evictions);
  m.def("schedule_t1_cache_moves(Tensor! cache_moves_idx, Tensor! cache_moves_count,"
        " Tensor evicted_logical_indices, Tensor evicted_kv_count, Tensor evicted_kv_offsets,"
        " Tensor block_tables, Tensor context_lens, int block_size) -> ()");
  m.impl("schedule_t1_cache_moves", torch::kCUDA, &schedule_t1_cache_moves);
  m.def("execute_cache_moves(Tensor! k_cache, Tensor! v_cache, Tensor! kv_metrics,"
        " Tensor! kv_position, Tensor cache_moves_idx, Tensor cache_moves_count,"
        " Tensor evicted_kv_offsets, int blocks_per_head, int threads_per_head) -> ()");
  m.impl("execute_cache_moves", torch::kCUDA, &execute_cache_moves);
}
