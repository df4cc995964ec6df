// Torch dispatcher registration for libkvc_b200: the reference's own operator surface.
//
// Same namespaces, op names and schemas as the reference extension
// (csrc/torch_bindings.cpp:52-80 `_C`, :353-362 `_C_cache_ops`, :371-418 `_C_kvc_ops`;
// C++ signatures csrc/kvcompress_eviction.h:5-78, csrc/kvcompress_cache.h:5-15,
// csrc/ops.h:18-37, 67-90), so vllm/_custom_ops.py of the fork resolves
// torch.ops._C.kvcompress_paged_attention_v1 etc. against this library unchanged.
// Every op validates dtype / contiguity / device (the reference reinterpret_casts
// blindly, csrc/kvcompress_eviction_kernels.cu:527-542), forwards raw device pointers
// to the C ABI (include/kvc_b200.h) on the current CUDA stream, and turns a non-zero
// status into the RuntimeError the reference's TORCH_CHECK would raise.
// CUDA dispatch key only: there is no CPU implementation and no fallback.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/all.h>
#include <torch/library.h>
#include <Python.h>

#include "kvc_b200.h"

namespace {

inline kvc_stream_t cur_stream() {
  return reinterpret_cast<kvc_stream_t>(at::cuda::getCurrentCUDAStream().stream());
}

inline void check_status(int rc, const char* op) {
  TORCH_CHECK(rc == KVC_OK, op, ": ", kvc_last_error());
}

inline const int* i32(const torch::Tensor& t, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == at::kInt, name, " must be int32, got ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  return t.data_ptr<int>();
}
inline int* i32_mut(torch::Tensor& t, const char* name) { return const_cast<int*>(i32(t, name)); }

inline int cache_dtype(const torch::Tensor& t) {
  switch (t.scalar_type()) {
    case at::kHalf: return KVC_F16;
    case at::kBFloat16: return KVC_BF16;
    case at::kFloat: return KVC_F32;
    case at::kFloat8_e5m2:
    case at::kByte: return KVC_U8;
    default: TORCH_CHECK(false, "Unsupported data type: ", t.scalar_type());
  }
  return -1;
}

// ---- _C_kvc_ops ------------------------------------------------------------
void schedule_cache_evictions(
    torch::Tensor& evicted_kv_indices, torch::Tensor& evicted_logical_indices,
    torch::Tensor& evicted_kv_count, torch::Tensor& remaining_kv_count,
    torch::Tensor& evicted_kv_offsets, torch::Tensor& sorted_indices,
    torch::Tensor& seq_block_offsets, torch::Tensor& layer_by_block, torch::Tensor& head_by_block,
    torch::Tensor& virtual_block_num_by_block, torch::Tensor& evicted_blocks_per_seq,
    torch::Tensor& context_lens, torch::Tensor& hanging_token_count, torch::Tensor& kv_position,
    torch::Tensor& last_position, torch::Tensor& protected_window_size, int64_t block_size,
    bool evict_evenly_per_layer, const c10::optional<torch::Tensor>& control_layers,
    int64_t max_evicted_kv, int64_t null_eviction_index, bool truncate) {
  TORCH_CHECK(evicted_kv_count.dim() == 3, "evicted_kv_count must be [num_seqs, num_layers, num_kv_heads]");
  const int S = evicted_kv_count.size(0), L = evicted_kv_count.size(1), H = evicted_kv_count.size(2);
  const int total_blocks = layer_by_block.size(0);
  const at::cuda::OptionalCUDAGuard guard(device_of(evicted_kv_indices));
  const size_t ws_bytes = kvc_schedule_cache_evictions_workspace_bytes(S, L, H, total_blocks, (int)block_size);
  auto ws = torch::empty({(int64_t)ws_bytes}, evicted_kv_indices.options().dtype(at::kByte));
  const int* ctl = nullptr;
  int nctl = 0;
  if (control_layers) {
    ctl = i32(control_layers.value(), "control_layers");
    nctl = control_layers.value().numel();
  }
  check_status(
      kvc_schedule_cache_evictions(
          i32_mut(evicted_kv_indices, "evicted_kv_indices"),
          i32_mut(evicted_logical_indices, "evicted_logical_indices"),
          i32_mut(evicted_kv_count, "evicted_kv_count"), i32_mut(remaining_kv_count, "remaining_kv_count"),
          i32(evicted_kv_offsets, "evicted_kv_offsets"), i32(sorted_indices, "sorted_indices"),
          i32(seq_block_offsets, "seq_block_offsets"), i32(layer_by_block, "layer_by_block"),
          i32(head_by_block, "head_by_block"), i32(virtual_block_num_by_block, "virtual_block_num_by_block"),
          i32(evicted_blocks_per_seq, "evicted_blocks_per_seq"), i32(context_lens, "context_lens"),
          i32(hanging_token_count, "hanging_token_count"), i32(kv_position, "kv_position"),
          i32(last_position, "last_position"), i32(protected_window_size, "protected_window_size"), S, L, H,
          total_blocks, (int)block_size, evict_evenly_per_layer ? 1 : 0, ctl, nctl, (int)max_evicted_kv,
          (int)null_eviction_index, truncate ? 1 : 0, ws.data_ptr(), ws_bytes, cur_stream()),
      "schedule_cache_evictions");
}

void truncate_cache_evictions(torch::Tensor& evicted_kv_indices, torch::Tensor& evicted_logical_indices,
                              torch::Tensor& evicted_kv_count, torch::Tensor& evicted_kv_offsets,
                              torch::Tensor& hanging_token_count, int64_t block_size,
                              int64_t max_evicted_kv, int64_t null_eviction_index) {
  TORCH_CHECK(evicted_kv_count.dim() == 3, "evicted_kv_count must be [num_seqs, num_layers, num_kv_heads]");
  const at::cuda::OptionalCUDAGuard guard(device_of(evicted_logical_indices));
  check_status(kvc_truncate_cache_evictions(
                   i32_mut(evicted_kv_indices, "evicted_kv_indices"),
                   i32_mut(evicted_logical_indices, "evicted_logical_indices"),
                   i32_mut(evicted_kv_count, "evicted_kv_count"), i32(evicted_kv_offsets, "evicted_kv_offsets"),
                   i32(hanging_token_count, "hanging_token_count"), evicted_kv_count.size(0),
                   evicted_kv_count.size(1), evicted_kv_count.size(2), (int)block_size, (int)max_evicted_kv,
                   (int)null_eviction_index, cur_stream()),
               "truncate_cache_evictions");
}

void count_block_evictions(torch::Tensor& evicted_block_count, torch::Tensor& evicted_logical_indices,
                           torch::Tensor& evicted_kv_offsets, torch::Tensor& hanging_token_count,
                           int64_t block_size, int64_t null_value) {
  TORCH_CHECK(evicted_block_count.dim() == 3, "evicted_block_count must be [num_seqs, num_layers, num_kv_heads]");
  const at::cuda::OptionalCUDAGuard guard(device_of(evicted_logical_indices));
  check_status(kvc_count_block_evictions(
                   i32_mut(evicted_block_count, "evicted_block_count"),
                   i32_mut(evicted_logical_indices, "evicted_logical_indices"),
                   i32(evicted_kv_offsets, "evicted_kv_offsets"), i32(hanging_token_count, "hanging_token_count"),
                   evicted_block_count.size(0), evicted_block_count.size(1), evicted_block_count.size(2),
                   (int)evicted_logical_indices.numel(), (int)block_size, (int)null_value, cur_stream()),
               "count_block_evictions");
}

void schedule_t1_cache_moves(torch::Tensor& cache_moves_idx, torch::Tensor& cache_moves_count,
                             torch::Tensor& evicted_logical_indices, torch::Tensor& evicted_kv_count,
                             torch::Tensor& evicted_kv_offsets, torch::Tensor& block_tables,
                             torch::Tensor& context_lens, int64_t block_size) {
  TORCH_CHECK(evicted_kv_count.dim() == 3, "evicted_kv_count must be [num_seqs, num_layers, num_kv_heads]");
  TORCH_CHECK(block_tables.dim() == 4, "block_tables must be [num_layers, num_seqs, num_kv_heads, max_blocks]");
  const at::cuda::OptionalCUDAGuard guard(device_of(evicted_logical_indices));
  check_status(kvc_schedule_t1_cache_moves(
                   i32_mut(cache_moves_idx, "cache_moves_idx"), i32_mut(cache_moves_count, "cache_moves_count"),
                   i32(evicted_logical_indices, "evicted_logical_indices"),
                   i32(evicted_kv_count, "evicted_kv_count"), i32(evicted_kv_offsets, "evicted_kv_offsets"),
                   i32(block_tables, "block_tables"), i32(context_lens, "context_lens"),
                   evicted_kv_count.size(0), evicted_kv_count.size(1), evicted_kv_count.size(2),
                   block_tables.size(3), (int)block_size, cur_stream()),
               "schedule_t1_cache_moves");
}

void execute_cache_moves(torch::Tensor& k_cache, torch::Tensor& v_cache, torch::Tensor& kv_metrics,
                         torch::Tensor& kv_position, torch::Tensor& cache_moves_idx,
                         torch::Tensor& cache_moves_count, torch::Tensor& evicted_kv_offsets,
                         int64_t blocks_per_head, int64_t threads_per_head) {
  TORCH_CHECK(k_cache.dim() == 4 && v_cache.dim() == 3, "k_cache [NB, D/x, B, x], v_cache [NB, D, B] expected");
  TORCH_CHECK(k_cache.is_cuda() && v_cache.is_cuda() && k_cache.is_contiguous() && v_cache.is_contiguous(),
              "k_cache / v_cache must be contiguous CUDA tensors");
  TORCH_CHECK(k_cache.scalar_type() == v_cache.scalar_type(), "k_cache / v_cache dtype mismatch");
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous() && kv_metrics.is_cuda(),
              "kv_metrics must be a contiguous float32 CUDA tensor");
  TORCH_CHECK(cache_moves_count.dim() == 3, "cache_moves_count must be [num_seqs, num_layers, num_kv_heads]");
  const at::cuda::OptionalCUDAGuard guard(device_of(cache_moves_idx));
  check_status(kvc_execute_cache_moves(
                   k_cache.data_ptr(), v_cache.data_ptr(), kv_metrics.data_ptr<float>(),
                   i32_mut(kv_position, "kv_position"), i32(cache_moves_idx, "cache_moves_idx"),
                   i32(cache_moves_count, "cache_moves_count"), i32(evicted_kv_offsets, "evicted_kv_offsets"),
                   cache_moves_count.size(0), cache_moves_count.size(1), cache_moves_count.size(2),
                   v_cache.size(0), v_cache.size(1), k_cache.size(3), v_cache.size(2), cache_dtype(k_cache),
                   (int)blocks_per_head, (int)threads_per_head, cur_stream()),
               "execute_cache_moves");
}

// additive op: see include/kvc_b200.h kvc_execute_cache_moves_deferred
void execute_cache_moves_deferred(torch::Tensor& k_cache, torch::Tensor& v_cache, torch::Tensor& kv_metrics,
                                  torch::Tensor& kv_position, torch::Tensor& cache_moves_idx,
                                  torch::Tensor& cache_moves_count, torch::Tensor& evicted_kv_offsets,
                                  torch::Tensor& pend_kv, torch::Tensor& pend_slot,
                                  const c10::optional<torch::Tensor>& pend_dst) {
  TORCH_CHECK(k_cache.dim() == 4 && v_cache.dim() == 3, "k_cache [NB, D/x, B, x], v_cache [NB, D, B] expected");
  TORCH_CHECK(k_cache.is_cuda() && v_cache.is_cuda() && k_cache.is_contiguous() && v_cache.is_contiguous(),
              "k_cache / v_cache must be contiguous CUDA tensors");
  TORCH_CHECK(k_cache.scalar_type() == v_cache.scalar_type(), "k_cache / v_cache dtype mismatch");
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous() && kv_metrics.is_cuda(),
              "kv_metrics must be a contiguous float32 CUDA tensor");
  TORCH_CHECK(cache_moves_count.dim() == 3, "cache_moves_count must be [num_seqs, num_layers, num_kv_heads]");
  const int64_t heads = cache_moves_count.numel();
  TORCH_CHECK(pend_kv.scalar_type() == k_cache.scalar_type() && pend_kv.is_contiguous() && pend_kv.is_cuda() &&
                  pend_kv.numel() == heads * 2 * v_cache.size(1),
              "pend_kv must be contiguous [num_layers, num_seqs, num_kv_heads, 2 * head_size] of the cache dtype");
  TORCH_CHECK(pend_slot.scalar_type() == at::kLong && pend_slot.is_contiguous() && pend_slot.is_cuda() &&
                  pend_slot.numel() == heads,
              "pend_slot must be contiguous int64 [num_layers, num_seqs, num_kv_heads]");
  int64_t* pend_dst_ptr = nullptr;
  if (pend_dst) {
    const auto& t = pend_dst.value();
    TORCH_CHECK(t.scalar_type() == at::kLong && t.is_contiguous() && t.is_cuda() && t.numel() == heads,
                "pend_dst must be contiguous int64 [num_layers, num_seqs, num_kv_heads]");
    pend_dst_ptr = t.data_ptr<int64_t>();
  }
  const at::cuda::OptionalCUDAGuard guard(device_of(cache_moves_idx));
  check_status(kvc_execute_cache_moves_deferred(
                   k_cache.data_ptr(), v_cache.data_ptr(), kv_metrics.data_ptr<float>(),
                   i32_mut(kv_position, "kv_position"), i32(cache_moves_idx, "cache_moves_idx"),
                   i32(cache_moves_count, "cache_moves_count"), i32(evicted_kv_offsets, "evicted_kv_offsets"),
                   pend_kv.data_ptr(), pend_slot.data_ptr<int64_t>(), pend_dst_ptr, cache_moves_count.size(0),
                   cache_moves_count.size(1), cache_moves_count.size(2), v_cache.size(0), v_cache.size(1),
                   k_cache.size(3), v_cache.size(2), cache_dtype(k_cache), cur_stream()),
               "execute_cache_moves_deferred");
}

// additive op: see include/kvc_b200.h kvc_flush_deferred
void flush_deferred(torch::Tensor& k_cache, torch::Tensor& v_cache, torch::Tensor& pend_kv, torch::Tensor& pend_slot,
                    const c10::optional<torch::Tensor>& pend_dst) {
  TORCH_CHECK(k_cache.dim() == 4 && v_cache.dim() == 3 && k_cache.is_cuda() && k_cache.is_contiguous() &&
                  v_cache.is_contiguous(), "k_cache [NB, D/x, B, x], v_cache [NB, D, B] contiguous CUDA tensors expected");
  const int64_t heads = pend_slot.numel();
  TORCH_CHECK(pend_kv.scalar_type() == k_cache.scalar_type() && pend_kv.is_contiguous() &&
                  pend_kv.numel() == heads * 2 * v_cache.size(1), "pend_kv must be contiguous [heads, 2 * head_size]");
  TORCH_CHECK(pend_slot.scalar_type() == at::kLong && pend_slot.is_contiguous(), "pend_slot must be contiguous int64");
  int64_t* pend_dst_ptr = nullptr;
  if (pend_dst) {
    const auto& t = pend_dst.value();
    TORCH_CHECK(t.scalar_type() == at::kLong && t.is_contiguous() && t.numel() == heads, "pend_dst must match pend_slot");
    pend_dst_ptr = t.data_ptr<int64_t>();
  }
  const at::cuda::OptionalCUDAGuard guard(device_of(k_cache));
  check_status(kvc_flush_deferred(k_cache.data_ptr(), v_cache.data_ptr(), pend_kv.data_ptr(),
                                  pend_slot.data_ptr<int64_t>(), pend_dst_ptr, heads, v_cache.size(1), k_cache.size(3),
                                  v_cache.size(2), cache_dtype(k_cache), cur_stream()),
               "flush_deferred");
}

// additive op: see include/kvc_b200.h kvc_select_evictions
void select_evictions(torch::Tensor& evicted_logical_indices, torch::Tensor& evicted_kv_count,
                      torch::Tensor& evicted_block_count, torch::Tensor& kv_metrics,
                      torch::Tensor& kv_position, torch::Tensor& block_tables, torch::Tensor& context_lens,
                      torch::Tensor& hanging_token_count, torch::Tensor& evicted_kv_offsets,
                      torch::Tensor& evicted_blocks_per_seq, torch::Tensor& last_position,
                      torch::Tensor& protected_window_size, int64_t block_size, int64_t null_eviction_index,
                      const c10::optional<torch::Tensor>& head_bias, const c10::optional<torch::Tensor>& position_bins,
                      double bias_weight, bool use_average, int64_t num_sinks) {
  const float* bias_ptr = nullptr;
  const int* bins_ptr = nullptr;
  int num_bins = 0;
  if (head_bias) {
    const auto& b = head_bias.value();
    TORCH_CHECK(position_bins.has_value(), "head_bias needs position_bins");
    TORCH_CHECK(b.is_cuda() && b.scalar_type() == at::kFloat && b.is_contiguous() && b.dim() == 3,
                "head_bias must be a contiguous float32 [num_layers, num_kv_heads, num_bins] CUDA tensor");
    TORCH_CHECK(b.size(0) == evicted_kv_count.size(1) && b.size(1) == evicted_kv_count.size(2) &&
                    b.size(2) == position_bins.value().numel(),
                "head_bias shape does not match [num_layers, num_kv_heads, len(position_bins)]");
    bias_ptr = b.data_ptr<float>();
    bins_ptr = i32(position_bins.value(), "position_bins");
    num_bins = b.size(2);
  }
  TORCH_CHECK(evicted_kv_count.dim() == 3, "evicted_kv_count must be [num_seqs, num_layers, num_kv_heads]");
  TORCH_CHECK(block_tables.dim() == 4, "block_tables must be [num_layers, num_seqs, num_kv_heads, max_blocks]");
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous() && kv_metrics.is_cuda(),
              "kv_metrics must be a contiguous float32 CUDA tensor");
  const at::cuda::OptionalCUDAGuard guard(device_of(evicted_logical_indices));
  const int64_t max_kv = evicted_logical_indices.numel();
  const size_t ws_bytes = kvc_select_evictions_workspace_bytes_ex(max_kv, (int)block_size, (int)evicted_kv_count.size(0),
                                                                  (int)block_tables.size(3));
  auto ws = torch::empty({(int64_t)ws_bytes}, evicted_logical_indices.options().dtype(at::kByte));
  check_status(kvc_select_evictions_ex(
                   i32_mut(evicted_logical_indices, "evicted_logical_indices"),
                   i32_mut(evicted_kv_count, "evicted_kv_count"), i32_mut(evicted_block_count, "evicted_block_count"),
                   kv_metrics.data_ptr<float>(), i32(kv_position, "kv_position"), i32(block_tables, "block_tables"),
                   i32(context_lens, "context_lens"), i32(hanging_token_count, "hanging_token_count"),
                   i32(evicted_kv_offsets, "evicted_kv_offsets"), i32(evicted_blocks_per_seq, "evicted_blocks_per_seq"),
                   i32(last_position, "last_position"), i32(protected_window_size, "protected_window_size"),
                   evicted_kv_count.size(0), evicted_kv_count.size(1), evicted_kv_count.size(2),
                   block_tables.size(3), (int)block_size, max_kv, (int)null_eviction_index, bias_ptr, bins_ptr,
                   num_bins, (float)bias_weight, use_average ? 1 : 0, (int)num_sinks, ws.data_ptr(), ws_bytes,
                   cur_stream()),
               "select_evictions");
}

// additive ops: see include/kvc_b200.h kvc_prefill_kv_metrics / kvc_aggregate_prefill
void prefill_kv_metrics(const c10::optional<torch::Tensor>& kv_metric_out,
                        const c10::optional<torch::Tensor>& kv_metrics,
                        const c10::optional<torch::Tensor>& slot_mapping, torch::Tensor& query, torch::Tensor& key,
                        torch::Tensor& seq_start_loc, torch::Tensor& kv_metric_buffer_len, int64_t max_prompt_len,
                        double scale, int64_t n_observed, bool use_l2, bool use_average, bool use_maxpool) {
  TORCH_CHECK(query.dim() == 3 && key.dim() == 3, "query [T, Hq, D] and key [T, H, D] expected");
  TORCH_CHECK(query.is_cuda() && key.is_cuda() && query.scalar_type() == key.scalar_type(),
              "query and key must be CUDA tensors of one dtype");
  TORCH_CHECK(query.stride(2) == 1 && key.stride(2) == 1 && query.stride(1) == query.size(2) &&
                  key.stride(1) == key.size(2),
              "query/key must be contiguous within a token");
  TORCH_CHECK(query.size(0) == key.size(0), "query and key must hold the same tokens");
  TORCH_CHECK(kv_metrics.has_value() == slot_mapping.has_value(), "kv_metrics and slot_mapping go together");
  TORCH_CHECK(kv_metric_out.has_value() || kv_metrics.has_value(), "nothing to write");
  const int T = query.size(0), Hq = query.size(1), H = key.size(1), D = query.size(2);
  const int num_seqs = seq_start_loc.numel() - 1;
  float* out_ptr = nullptr;
  float* metrics_ptr = nullptr;
  const int64_t* slots_ptr = nullptr;
  if (kv_metric_out) {
    const auto& t = kv_metric_out.value();
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous() && t.numel() == (int64_t)T * Hq,
                "kv_metric_out must be a contiguous float32 [T, Hq] CUDA tensor");
    out_ptr = t.data_ptr<float>();
  }
  if (kv_metrics) {
    const auto& m = kv_metrics.value();
    const auto& sm = slot_mapping.value();
    TORCH_CHECK(m.is_cuda() && m.scalar_type() == at::kFloat && m.is_contiguous(), "kv_metrics must be contiguous float32");
    TORCH_CHECK(sm.is_cuda() && sm.scalar_type() == at::kLong && sm.is_contiguous() && sm.numel() == (int64_t)T * H,
                "slot_mapping must be a contiguous int64 [T, H] CUDA tensor");
    metrics_ptr = m.data_ptr<float>();
    slots_ptr = sm.data_ptr<int64_t>();
  }
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const size_t ws_bytes = kvc_prefill_kv_metrics_workspace_bytes(T, num_seqs, Hq, H, (int)n_observed, (int)max_prompt_len);
  auto ws = torch::empty({(int64_t)ws_bytes + 1}, query.options().dtype(at::kByte));
  check_status(kvc_prefill_kv_metrics(out_ptr, metrics_ptr, slots_ptr, query.data_ptr(), key.data_ptr(),
                                      i32(seq_start_loc, "seq_start_loc"),
                                      i32(kv_metric_buffer_len, "kv_metric_buffer_len"), T, num_seqs, Hq, H, D,
                                      query.stride(0), key.stride(0), (int)max_prompt_len, (float)scale,
                                      (int)n_observed, use_l2, use_average, use_maxpool, cache_dtype(query),
                                      ws.data_ptr(), ws_bytes, cur_stream()),
               "prefill_kv_metrics");
}

void aggregate_prefill(torch::Tensor& kv_metrics, torch::Tensor& prefill_metrics, torch::Tensor& slot_mapping) {
  TORCH_CHECK(prefill_metrics.dim() == 2 && slot_mapping.dim() == 2 && prefill_metrics.size(0) == slot_mapping.size(0),
              "prefill_metrics [T, Hq] and slot_mapping [T, H] expected");
  TORCH_CHECK(kv_metrics.is_cuda() && kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous(),
              "kv_metrics must be a contiguous float32 CUDA tensor");
  TORCH_CHECK(prefill_metrics.is_cuda() && prefill_metrics.scalar_type() == at::kFloat && prefill_metrics.is_contiguous(),
              "prefill_metrics must be a contiguous float32 CUDA tensor");
  TORCH_CHECK(slot_mapping.is_cuda() && slot_mapping.scalar_type() == at::kLong && slot_mapping.is_contiguous(),
              "slot_mapping must be a contiguous int64 CUDA tensor");
  const at::cuda::OptionalCUDAGuard guard(device_of(kv_metrics));
  check_status(kvc_aggregate_prefill(kv_metrics.data_ptr<float>(), prefill_metrics.data_ptr<float>(),
                                     slot_mapping.data_ptr<int64_t>(), prefill_metrics.size(0), prefill_metrics.size(1),
                                     slot_mapping.size(1), cur_stream()),
               "aggregate_prefill");
}

}  // namespace

// additive ops: device-side block allocator (include/kvc_b200.h kvc_alloc_*).  `state` tensors are the
// reference's own: free_mask [NB] bool, block_tables [L,MS,H,MB], context_lens [L,MS,H], token positions.
namespace {
struct AllocState {
  uint8_t* free_mask;
  int* free_count;
  int* status;
  int64_t num_blocks;
  int L, MS, H, MB;
};
inline AllocState alloc_state(torch::Tensor& free_mask, torch::Tensor& free_count, torch::Tensor& status,
                              torch::Tensor& block_tables, torch::Tensor& context_lens) {
  TORCH_CHECK(free_mask.is_cuda() && free_mask.is_contiguous() &&
                  (free_mask.scalar_type() == at::kBool || free_mask.scalar_type() == at::kByte),
              "free_mask must be a contiguous bool CUDA tensor");
  TORCH_CHECK(block_tables.dim() == 4 && context_lens.dim() == 3, "block_tables [L,S,H,MB], context_lens [L,S,H] expected");
  AllocState a;
  a.free_mask = reinterpret_cast<uint8_t*>(free_mask.data_ptr());
  a.free_count = i32_mut(free_count, "free_count");
  a.status = i32_mut(status, "status");
  a.num_blocks = free_mask.numel();
  a.L = block_tables.size(0); a.MS = block_tables.size(1); a.H = block_tables.size(2); a.MB = block_tables.size(3);
  return a;
}
inline int* opt_i32(const c10::optional<torch::Tensor>& t, const char* name) {
  return t ? const_cast<int*>(i32(t.value(), name)) : nullptr;
}
}  // namespace

void alloc_init(torch::Tensor& free_mask, torch::Tensor& free_count, torch::Tensor& status) {
  const at::cuda::OptionalCUDAGuard guard(device_of(free_mask));
  check_status(kvc_alloc_init(reinterpret_cast<uint8_t*>(free_mask.data_ptr()), i32_mut(free_count, "free_count"),
                              i32_mut(status, "status"), free_mask.numel(), cur_stream()),
               "alloc_init");
}

void alloc_append_slots(torch::Tensor& free_mask, torch::Tensor& free_count, torch::Tensor& status,
                        torch::Tensor& block_tables, torch::Tensor& context_lens,
                        const c10::optional<torch::Tensor>& seq_slots, torch::Tensor& seq_len,
                        torch::Tensor& slot_mapping, torch::Tensor& kv_position,
                        const c10::optional<torch::Tensor>& seq_index_by_block,
                        const c10::optional<torch::Tensor>& layer_index_by_block,
                        const c10::optional<torch::Tensor>& head_index_by_block,
                        const c10::optional<torch::Tensor>& logical_block_num_by_block, int64_t block_size) {
  AllocState a = alloc_state(free_mask, free_count, status, block_tables, context_lens);
  TORCH_CHECK(slot_mapping.scalar_type() == at::kLong && slot_mapping.is_contiguous(), "slot_mapping must be contiguous int64");
  const int n = seq_len.numel();
  const at::cuda::OptionalCUDAGuard guard(device_of(free_mask));
  const size_t ws_bytes = kvc_alloc_workspace_bytes(a.num_blocks, n, a.L, a.H);
  auto ws = torch::empty({(int64_t)ws_bytes + 1}, free_mask.options().dtype(at::kByte));
  check_status(kvc_alloc_append_slots(a.free_mask, a.free_count, a.status, a.num_blocks, i32_mut(block_tables, "block_tables"),
                                      i32_mut(context_lens, "context_lens"), opt_i32(seq_slots, "seq_slots"),
                                      i32(seq_len, "seq_len"), slot_mapping.data_ptr<int64_t>(),
                                      i32_mut(kv_position, "kv_position"), opt_i32(seq_index_by_block, "seq_index_by_block"),
                                      opt_i32(layer_index_by_block, "layer_index_by_block"),
                                      opt_i32(head_index_by_block, "head_index_by_block"),
                                      opt_i32(logical_block_num_by_block, "logical_block_num_by_block"), n, a.MS, a.L, a.H,
                                      a.MB, (int)block_size, ws.data_ptr(), ws_bytes, cur_stream()),
               "alloc_append_slots");
}

void alloc_add_sequences(torch::Tensor& free_mask, torch::Tensor& free_count, torch::Tensor& status,
                         torch::Tensor& block_tables, torch::Tensor& context_lens, torch::Tensor& seq_slots,
                         torch::Tensor& seq_lens, int64_t max_seq_len, torch::Tensor& kv_position,
                         const c10::optional<torch::Tensor>& seq_index_by_block,
                         const c10::optional<torch::Tensor>& layer_index_by_block,
                         const c10::optional<torch::Tensor>& head_index_by_block,
                         const c10::optional<torch::Tensor>& logical_block_num_by_block, int64_t block_size) {
  AllocState a = alloc_state(free_mask, free_count, status, block_tables, context_lens);
  const at::cuda::OptionalCUDAGuard guard(device_of(free_mask));
  const size_t ws_bytes = kvc_alloc_workspace_bytes(a.num_blocks, 0, a.L, a.H);
  auto ws = torch::empty({(int64_t)ws_bytes + 1}, free_mask.options().dtype(at::kByte));
  check_status(kvc_alloc_add_sequences(a.free_mask, a.free_count, a.status, a.num_blocks, i32_mut(block_tables, "block_tables"),
                                       i32_mut(context_lens, "context_lens"), i32(seq_slots, "seq_slots"),
                                       i32(seq_lens, "seq_lens"), i32_mut(kv_position, "kv_position"),
                                       opt_i32(seq_index_by_block, "seq_index_by_block"),
                                       opt_i32(layer_index_by_block, "layer_index_by_block"),
                                       opt_i32(head_index_by_block, "head_index_by_block"),
                                       opt_i32(logical_block_num_by_block, "logical_block_num_by_block"),
                                       (int)seq_lens.numel(), (int)max_seq_len, a.MS, a.L, a.H, a.MB, (int)block_size,
                                       ws.data_ptr(), ws_bytes, cur_stream()),
               "alloc_add_sequences");
}

void alloc_free_compressed(torch::Tensor& free_mask, torch::Tensor& free_count, torch::Tensor& block_tables,
                           torch::Tensor& context_lens, torch::Tensor& evicted_block_count,
                           const c10::optional<torch::Tensor>& seq_slots,
                           const c10::optional<torch::Tensor>& seq_index_by_block, int64_t block_size) {
  TORCH_CHECK(evicted_block_count.dim() == 3, "evicted_block_count must be [num_batch, num_layers, num_kv_heads]");
  TORCH_CHECK(block_tables.dim() == 4, "block_tables [L,S,H,MB] expected");
  const at::cuda::OptionalCUDAGuard guard(device_of(free_mask));
  check_status(kvc_alloc_free_compressed(reinterpret_cast<uint8_t*>(free_mask.data_ptr()), i32_mut(free_count, "free_count"),
                                         i32(block_tables, "block_tables"), i32_mut(context_lens, "context_lens"),
                                         i32(evicted_block_count, "evicted_block_count"), opt_i32(seq_slots, "seq_slots"),
                                         opt_i32(seq_index_by_block, "seq_index_by_block"), evicted_block_count.size(0),
                                         block_tables.size(1), block_tables.size(0), block_tables.size(2),
                                         block_tables.size(3), (int)block_size, cur_stream()),
               "alloc_free_compressed");
}

void alloc_remove_sequences(torch::Tensor& free_mask, torch::Tensor& free_count, torch::Tensor& block_tables,
                            torch::Tensor& context_lens, torch::Tensor& seq_slots,
                            const c10::optional<torch::Tensor>& seq_index_by_block, int64_t block_size) {
  TORCH_CHECK(block_tables.dim() == 4, "block_tables [L,S,H,MB] expected");
  const at::cuda::OptionalCUDAGuard guard(device_of(free_mask));
  check_status(kvc_alloc_remove_sequences(reinterpret_cast<uint8_t*>(free_mask.data_ptr()), i32_mut(free_count, "free_count"),
                                          i32(block_tables, "block_tables"), i32_mut(context_lens, "context_lens"),
                                          i32(seq_slots, "seq_slots"), opt_i32(seq_index_by_block, "seq_index_by_block"),
                                          (int)seq_slots.numel(), block_tables.size(1), block_tables.size(0),
                                          block_tables.size(2), block_tables.size(3), (int)block_size, cur_stream()),
               "alloc_remove_sequences");
}

namespace {
// additive ops: see include/kvc_b200.h kvc_prepare_compression / kvc_free_and_append_slots
void prepare_compression(torch::Tensor& context_lens, torch::Tensor& hanging_token_count,
                         torch::Tensor& evicted_kv_offsets, torch::Tensor& evicted_blocks_per_seq,
                         int64_t block_size, int64_t max_cache_blocks_per_seq) {
  TORCH_CHECK(context_lens.dim() == 3, "context_lens must be [num_layers, num_seqs, num_kv_heads]");
  const int L = context_lens.size(0), S = context_lens.size(1), H = context_lens.size(2);
  const at::cuda::OptionalCUDAGuard guard(device_of(context_lens));
  auto ws = torch::empty({S}, context_lens.options());
  check_status(kvc_prepare_compression(i32(context_lens, "context_lens"), i32_mut(hanging_token_count, "hanging_token_count"),
                                       i32_mut(evicted_kv_offsets, "evicted_kv_offsets"),
                                       i32_mut(evicted_blocks_per_seq, "evicted_blocks_per_seq"), ws.data_ptr<int>(), S, L,
                                       H, (int)block_size, (int)max_cache_blocks_per_seq, cur_stream()),
               "prepare_compression");
}

void free_and_append_slots(torch::Tensor& context_lens, torch::Tensor& block_tables,
                           const c10::optional<torch::Tensor>& evicted_block_count,
                           const c10::optional<torch::Tensor>& hanging_token_count, torch::Tensor& seq_len,
                           torch::Tensor& slot_mapping, torch::Tensor& kv_position, int64_t block_size, bool append,
                           const c10::optional<torch::Tensor>& status) {
  TORCH_CHECK(context_lens.dim() == 3 && block_tables.dim() == 4, "context_lens [L,S,H], block_tables [L,S,H,MB] expected");
  TORCH_CHECK(slot_mapping.scalar_type() == at::kLong && slot_mapping.is_contiguous(), "slot_mapping must be contiguous int64");
  const int L = context_lens.size(0), S = context_lens.size(1), H = context_lens.size(2);
  const at::cuda::OptionalCUDAGuard guard(device_of(context_lens));
  check_status(kvc_free_and_append_slots(
                   i32_mut(context_lens, "context_lens"), i32(block_tables, "block_tables"),
                   evicted_block_count ? i32(evicted_block_count.value(), "evicted_block_count") : nullptr,
                   hanging_token_count ? i32(hanging_token_count.value(), "hanging_token_count") : nullptr,
                   i32(seq_len, "seq_len"), slot_mapping.data_ptr<int64_t>(), i32_mut(kv_position, "kv_position"), S, L, H,
                   block_tables.size(3), (int)block_size, append ? 1 : 0,
                   status ? const_cast<int*>(i32(status.value(), "status")) : nullptr, cur_stream()),
               "free_and_append_slots");
}

// ---- _C_cache_ops ----------------------------------------------------------
void kvcompress_reshape_and_cache(torch::Tensor& key, torch::Tensor& value, torch::Tensor& key_cache,
                                  torch::Tensor& value_cache, torch::Tensor& kv_metrics,
                                  torch::Tensor& slot_mapping, torch::Tensor& kv_metric_head_bias,
                                  const std::string& kv_cache_dtype, const double k_scale,
                                  const double v_scale) {
  (void)k_scale;
  (void)v_scale;
  TORCH_CHECK(kv_cache_dtype == "auto", "Unsupported data type of kv cache: ", kv_cache_dtype,
              " (this build implements kv_cache_dtype='auto')");
  TORCH_CHECK(key.dim() == 3 && value.dim() == 3, "key / value must be [num_tokens, num_heads, head_size]");
  TORCH_CHECK(key.stride(2) == 1 && key.stride(1) == key.size(2) && value.stride(2) == 1 &&
                  value.stride(1) == value.size(2),
              "key / value heads must be dense");
  TORCH_CHECK(key.scalar_type() == key_cache.scalar_type() && value.scalar_type() == value_cache.scalar_type() &&
                  key.scalar_type() == value.scalar_type(),
              "kv_cache_dtype='auto' requires key/value/cache of one dtype");
  TORCH_CHECK(key_cache.is_contiguous() && value_cache.is_contiguous(), "caches must be contiguous");
  TORCH_CHECK(slot_mapping.scalar_type() == at::kLong && slot_mapping.is_contiguous(), "slot_mapping must be contiguous int64");
  TORCH_CHECK(slot_mapping.numel() == key.size(0) * key.size(1), "slot_mapping must hold num_tokens*num_heads slots");
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metric_head_bias.scalar_type() == at::kFloat,
              "kv_metrics / kv_metric_head_bias must be float32");
  const at::cuda::OptionalCUDAGuard guard(device_of(key));
  check_status(kvc_reshape_and_cache(key.data_ptr(), value.data_ptr(), key_cache.data_ptr(),
                                     value_cache.data_ptr(), kv_metrics.data_ptr<float>(),
                                     slot_mapping.data_ptr<int64_t>(), kv_metric_head_bias.data_ptr<float>(),
                                     key.size(0), key.size(1), key.size(2), key_cache.size(2), key_cache.size(3),
                                     key.stride(0), value.stride(0), cache_dtype(key_cache), cur_stream()),
               "kvcompress_reshape_and_cache");
}

// ---- _C (attention) --------------------------------------------------------
struct AttnCommon {
  int S, Hq, D, MB;
  int64_t q_stride, kv_block_stride;
  const float* alibi;
  int dtype;
};

AttnCommon attn_common(torch::Tensor& out, torch::Tensor& query, torch::Tensor& key_cache,
                       torch::Tensor& value_cache, torch::Tensor& block_tables,
                       const c10::optional<torch::Tensor>& alibi_slopes, const std::string& kv_cache_dtype) {
  TORCH_CHECK(kv_cache_dtype == "auto", "Unsupported data type of kv cache: ", kv_cache_dtype,
              " (this build implements kv_cache_dtype='auto')");
  TORCH_CHECK(query.dim() == 3 && query.stride(2) == 1 && query.stride(1) == query.size(2),
              "query must be [num_seqs, num_heads, head_size] with dense heads");
  TORCH_CHECK(out.is_contiguous() && out.scalar_type() == query.scalar_type(), "out must be contiguous, query dtype");
  TORCH_CHECK(key_cache.scalar_type() == query.scalar_type() && value_cache.scalar_type() == query.scalar_type(),
              "kv_cache_dtype='auto' requires cache dtype == query dtype");
  TORCH_CHECK(block_tables.dim() == 3, "block_tables must be [num_seqs, num_kv_heads, max_blocks]");
  AttnCommon c;
  c.S = query.size(0); c.Hq = query.size(1); c.D = query.size(2); c.MB = block_tables.size(2);
  c.q_stride = query.stride(0);
  c.kv_block_stride = key_cache.stride(0);
  c.alibi = nullptr;
  if (alibi_slopes) {
    TORCH_CHECK(alibi_slopes.value().scalar_type() == at::kFloat && alibi_slopes.value().is_contiguous(),
                "alibi_slopes must be contiguous float32");
    c.alibi = alibi_slopes.value().data_ptr<float>();
  }
  c.dtype = cache_dtype(query);
  return c;
}

void kvcompress_paged_attention_v1(
    torch::Tensor& out, torch::Tensor& kv_metric_out, torch::Tensor& query, torch::Tensor& key_cache,
    torch::Tensor& value_cache, int64_t num_kv_heads, double scale, torch::Tensor& block_tables,
    torch::Tensor& context_lens, torch::Tensor& kv_position, torch::Tensor& last_position,
    torch::Tensor& kv_metric_buffer_len, int64_t block_size, int64_t max_context_len,
    const c10::optional<torch::Tensor>& alibi_slopes, const std::string& kv_cache_dtype, double k_scale,
    double v_scale, bool record_kv_metrics) {
  (void)k_scale; (void)v_scale;
  AttnCommon c = attn_common(out, query, key_cache, value_cache, block_tables, alibi_slopes, kv_cache_dtype);
  TORCH_CHECK(kv_metric_out.scalar_type() == at::kFloat && kv_metric_out.is_contiguous(), "kv_metric_out must be contiguous float32");
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  check_status(kvc_paged_attention_v1(
                   out.data_ptr(), kv_metric_out.data_ptr<float>(), query.data_ptr(), key_cache.data_ptr(),
                   value_cache.data_ptr(), c.S, c.Hq, (int)num_kv_heads, c.D, (float)scale,
                   i32(block_tables, "block_tables"), i32(context_lens, "context_lens"),
                   i32(kv_position, "kv_position"), i32(last_position, "last_position"),
                   i32(kv_metric_buffer_len, "kv_metric_buffer_len"), (int)block_size, (int)max_context_len,
                   c.MB, c.q_stride, c.kv_block_stride, c.alibi, c.dtype, record_kv_metrics ? 1 : 0, cur_stream()),
               "kvcompress_paged_attention_v1");
}

void kvcompress_paged_attention_v2(
    torch::Tensor& out, torch::Tensor& kv_metric_out, torch::Tensor& exp_sums, torch::Tensor& max_logits,
    torch::Tensor& tmp_out, torch::Tensor& tmp_kv_metric_out, torch::Tensor& query, torch::Tensor& key_cache,
    torch::Tensor& value_cache, int64_t num_kv_heads, double scale, torch::Tensor& block_tables,
    torch::Tensor& context_lens, torch::Tensor& kv_position, torch::Tensor& last_position,
    torch::Tensor& kv_metric_buffer_len, int64_t block_size, int64_t max_context_len,
    const c10::optional<torch::Tensor>& alibi_slopes, const std::string& kv_cache_dtype, double k_scale,
    double v_scale, bool record_kv_metrics) {
  (void)k_scale; (void)v_scale;
  AttnCommon c = attn_common(out, query, key_cache, value_cache, block_tables, alibi_slopes, kv_cache_dtype);
  TORCH_CHECK(kv_metric_out.scalar_type() == at::kFloat && kv_metric_out.is_contiguous(), "kv_metric_out must be contiguous float32");
  TORCH_CHECK(exp_sums.scalar_type() == at::kFloat && max_logits.scalar_type() == at::kFloat &&
                  exp_sums.is_contiguous() && max_logits.is_contiguous(), "exp_sums / max_logits must be contiguous float32");
  TORCH_CHECK(tmp_out.scalar_type() == query.scalar_type() && tmp_out.is_contiguous(), "tmp_out must be contiguous, query dtype");
  const int64_t parts = (max_context_len + 511) / 512;
  TORCH_CHECK(exp_sums.numel() >= (int64_t)c.S * c.Hq * parts && tmp_out.numel() >= (int64_t)c.S * c.Hq * parts * c.D,
              "v2 temporaries are smaller than [num_seqs, num_heads, ceil(max_context_len/512)]");
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  check_status(kvc_paged_attention_v2(
                   out.data_ptr(), kv_metric_out.data_ptr<float>(), exp_sums.data_ptr<float>(),
                   max_logits.data_ptr<float>(), tmp_out.data_ptr(),
                   tmp_kv_metric_out.defined() && tmp_kv_metric_out.numel() ? tmp_kv_metric_out.data_ptr<float>() : nullptr,
                   query.data_ptr(), key_cache.data_ptr(), value_cache.data_ptr(), c.S, c.Hq, (int)num_kv_heads,
                   c.D, (float)scale, i32(block_tables, "block_tables"), i32(context_lens, "context_lens"),
                   i32(kv_position, "kv_position"), i32(last_position, "last_position"),
                   i32(kv_metric_buffer_len, "kv_metric_buffer_len"), (int)block_size, (int)max_context_len,
                   c.MB, c.q_stride, c.kv_block_stride, c.alibi, c.dtype, record_kv_metrics ? 1 : 0, cur_stream()),
               "kvcompress_paged_attention_v2");
}

// additive op (not in the reference list): see include/kvc_b200.h kvc_paged_attention_fused_metrics
void kvcompress_paged_attention_fused_metrics(
    torch::Tensor& out, torch::Tensor& kv_metrics, torch::Tensor& query, torch::Tensor& key_cache,
    torch::Tensor& value_cache, int64_t num_kv_heads, double scale, torch::Tensor& block_tables,
    torch::Tensor& context_lens, torch::Tensor& kv_position, torch::Tensor& last_position,
    torch::Tensor& kv_metric_buffer_len, int64_t block_size, int64_t max_context_len,
    const c10::optional<torch::Tensor>& alibi_slopes, const std::string& kv_cache_dtype) {
  AttnCommon c = attn_common(out, query, key_cache, value_cache, block_tables, alibi_slopes, kv_cache_dtype);
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous(), "kv_metrics must be contiguous float32");
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const size_t ws_bytes = kvc_paged_attention_fused_metrics_workspace_bytes(c.S, c.Hq, (int)max_context_len);
  auto ws = torch::empty({(int64_t)ws_bytes}, query.options().dtype(at::kByte));
  check_status(kvc_paged_attention_fused_metrics(
                   out.data_ptr(), kv_metrics.data_ptr<float>(), query.data_ptr(), key_cache.data_ptr(),
                   value_cache.data_ptr(), c.S, c.Hq, (int)num_kv_heads, c.D, (float)scale,
                   i32(block_tables, "block_tables"), i32(context_lens, "context_lens"),
                   i32(kv_position, "kv_position"), i32(last_position, "last_position"),
                   i32(kv_metric_buffer_len, "kv_metric_buffer_len"), (int)block_size, (int)max_context_len,
                   c.MB, c.q_stride, c.kv_block_stride, c.alibi, c.dtype, ws_bytes ? ws.data_ptr() : nullptr,
                   ws_bytes, cur_stream()),
               "kvcompress_paged_attention_fused_metrics");
}

// additive op: see include/kvc_b200.h kvc_paged_attention_fused_step
void kvcompress_paged_attention_fused_step(
    torch::Tensor& out, torch::Tensor& kv_metrics, torch::Tensor& query, torch::Tensor& key, torch::Tensor& value,
    torch::Tensor& key_cache, torch::Tensor& value_cache, torch::Tensor& slot_mapping,
    torch::Tensor& kv_metric_head_bias, int64_t num_kv_heads, double scale, torch::Tensor& block_tables,
    torch::Tensor& context_lens, torch::Tensor& kv_position, torch::Tensor& last_position,
    torch::Tensor& kv_metric_buffer_len, int64_t block_size, int64_t max_context_len,
    const c10::optional<torch::Tensor>& alibi_slopes, const std::string& kv_cache_dtype) {
  AttnCommon c = attn_common(out, query, key_cache, value_cache, block_tables, alibi_slopes, kv_cache_dtype);
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous(), "kv_metrics must be contiguous float32");
  TORCH_CHECK(key.dim() == 3 && value.dim() == 3 && key.stride(2) == 1 && key.stride(1) == key.size(2) &&
                  value.stride(2) == 1 && value.stride(1) == value.size(2),
              "key / value must be [num_seqs, num_kv_heads, head_size] with dense heads");
  TORCH_CHECK(key.scalar_type() == query.scalar_type() && value.scalar_type() == query.scalar_type(), "key/value dtype");
  TORCH_CHECK(slot_mapping.scalar_type() == at::kLong && slot_mapping.is_contiguous() &&
                  slot_mapping.numel() == key.size(0) * key.size(1), "slot_mapping must be contiguous int64 [num_seqs*num_kv_heads]");
  TORCH_CHECK(kv_metric_head_bias.scalar_type() == at::kFloat, "kv_metric_head_bias must be float32");
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const size_t ws_bytes = kvc_paged_attention_fused_metrics_workspace_bytes(c.S, c.Hq, (int)max_context_len);
  auto ws = torch::empty({(int64_t)ws_bytes}, query.options().dtype(at::kByte));
  check_status(kvc_paged_attention_fused_step(
                   out.data_ptr(), kv_metrics.data_ptr<float>(), query.data_ptr(), key.data_ptr(), value.data_ptr(),
                   key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr<int64_t>(),
                   kv_metric_head_bias.data_ptr<float>(), c.S, c.Hq, (int)num_kv_heads, c.D, (float)scale,
                   i32(block_tables, "block_tables"), i32(context_lens, "context_lens"), i32(kv_position, "kv_position"),
                   i32(last_position, "last_position"), i32(kv_metric_buffer_len, "kv_metric_buffer_len"),
                   (int)block_size, (int)max_context_len, c.MB, c.q_stride, key.stride(0), value.stride(0),
                   c.kv_block_stride, c.alibi, c.dtype, ws.data_ptr(), ws_bytes, cur_stream()),
               "kvcompress_paged_attention_fused_step");
}


// additive op: see include/kvc_b200.h kvc_paged_attention_fused_step_deferred
void kvcompress_paged_attention_fused_step_deferred(
    torch::Tensor& out, torch::Tensor& kv_metrics, torch::Tensor& query, torch::Tensor& key, torch::Tensor& value,
    torch::Tensor& key_cache, torch::Tensor& value_cache, torch::Tensor& slot_mapping,
    torch::Tensor& kv_metric_head_bias, torch::Tensor& pend_kv, torch::Tensor& pend_slot,
    const c10::optional<torch::Tensor>& pend_dst, int64_t num_kv_heads,
    double scale, torch::Tensor& block_tables, torch::Tensor& context_lens, torch::Tensor& kv_position,
    torch::Tensor& last_position, torch::Tensor& kv_metric_buffer_len, int64_t block_size, int64_t max_context_len,
    const c10::optional<torch::Tensor>& alibi_slopes, const std::string& kv_cache_dtype) {
  AttnCommon c = attn_common(out, query, key_cache, value_cache, block_tables, alibi_slopes, kv_cache_dtype);
  TORCH_CHECK(kv_metrics.scalar_type() == at::kFloat && kv_metrics.is_contiguous(), "kv_metrics must be contiguous float32");
  TORCH_CHECK(key.dim() == 3 && value.dim() == 3 && key.stride(2) == 1 && key.stride(1) == key.size(2) &&
                  value.stride(2) == 1 && value.stride(1) == value.size(2),
              "key / value must be [num_seqs, num_kv_heads, head_size] with dense heads");
  TORCH_CHECK(key.scalar_type() == query.scalar_type() && value.scalar_type() == query.scalar_type(), "key/value dtype");
  TORCH_CHECK(slot_mapping.scalar_type() == at::kLong && slot_mapping.is_contiguous() &&
                  slot_mapping.numel() == key.size(0) * key.size(1), "slot_mapping must be contiguous int64 [num_seqs*num_kv_heads]");
  TORCH_CHECK(kv_metric_head_bias.scalar_type() == at::kFloat, "kv_metric_head_bias must be float32");
  TORCH_CHECK(pend_kv.scalar_type() == key_cache.scalar_type() && pend_kv.is_contiguous() &&
                  pend_kv.numel() == key.size(0) * key.size(1) * 2 * key.size(2),
              "pend_kv must be contiguous [num_seqs, num_kv_heads, 2 * head_size] of the cache dtype");
  TORCH_CHECK(pend_slot.scalar_type() == at::kLong && pend_slot.is_contiguous() &&
                  pend_slot.numel() == key.size(0) * key.size(1), "pend_slot must be contiguous int64 [num_seqs, num_kv_heads]");
  int64_t* pend_dst_ptr = nullptr;
  if (pend_dst) {
    const auto& t = pend_dst.value();
    TORCH_CHECK(t.scalar_type() == at::kLong && t.is_contiguous() && t.numel() == pend_slot.numel(),
                "pend_dst must be contiguous int64 [num_seqs, num_kv_heads]");
    pend_dst_ptr = t.data_ptr<int64_t>();
  }
  const at::cuda::OptionalCUDAGuard guard(device_of(query));
  const size_t ws_bytes = kvc_paged_attention_fused_metrics_workspace_bytes(c.S, c.Hq, (int)max_context_len);
  auto ws = torch::empty({(int64_t)ws_bytes}, query.options().dtype(at::kByte));
  check_status(kvc_paged_attention_fused_step_deferred(
                   out.data_ptr(), kv_metrics.data_ptr<float>(), query.data_ptr(), key.data_ptr(), value.data_ptr(),
                   key_cache.data_ptr(), value_cache.data_ptr(), slot_mapping.data_ptr<int64_t>(),
                   kv_metric_head_bias.data_ptr<float>(), pend_kv.data_ptr(), pend_slot.data_ptr<int64_t>(),
                   pend_dst_ptr, c.S, c.Hq,
                   (int)num_kv_heads, c.D, (float)scale, i32(block_tables, "block_tables"),
                   i32(context_lens, "context_lens"), i32(kv_position, "kv_position"),
                   i32(last_position, "last_position"), i32(kv_metric_buffer_len, "kv_metric_buffer_len"),
                   (int)block_size, (int)max_context_len, c.MB, c.q_stride, key.stride(0), value.stride(0),
                   c.kv_block_stride, c.alibi, c.dtype, ws.data_ptr(), ws_bytes, cur_stream()),
               "kvcompress_paged_attention_fused_step_deferred");
}

}  // namespace

// FRAGMENT so that the library can coexist with another extension that already owns `_C`.
TORCH_LIBRARY_FRAGMENT(_C, ops) {
  ops.def(
      "kvcompress_paged_attention_v1("
      "    Tensor! out, Tensor! kv_metric_out, Tensor query,"
      "    Tensor key_cache, Tensor value_cache,"
      "    int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor context_lens,"
      "    Tensor kv_position, Tensor last_position,"
      "    Tensor kv_metric_buffer_len, int block_size,"
      "    int max_context_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype, float k_scale, float v_scale,"
      "    bool record_kv_metrics) -> ()");
  ops.impl("kvcompress_paged_attention_v1", torch::kCUDA, &kvcompress_paged_attention_v1);
  ops.def(
      "kvcompress_paged_attention_v2("
      "    Tensor! out, Tensor! kv_metric_out,"
      "    Tensor exp_sums, Tensor max_logits,"
      "    Tensor tmp_out, Tensor tmp_kv_metric_out,"
      "    Tensor query, Tensor key_cache, Tensor value_cache,"
      "    int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor context_lens,"
      "    Tensor kv_position, Tensor last_position,"
      "    Tensor kv_metric_buffer_len, int block_size,"
      "    int max_context_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype, float k_scale, float v_scale,"
      "    bool record_kv_metrics) -> ()");
  ops.impl("kvcompress_paged_attention_v2", torch::kCUDA, &kvcompress_paged_attention_v2);
  ops.def(
      "kvcompress_paged_attention_fused_metrics("
      "    Tensor! out, Tensor! kv_metrics, Tensor query,"
      "    Tensor key_cache, Tensor value_cache,"
      "    int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor context_lens,"
      "    Tensor kv_position, Tensor last_position,"
      "    Tensor kv_metric_buffer_len, int block_size,"
      "    int max_context_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype) -> ()");
  ops.impl("kvcompress_paged_attention_fused_metrics", torch::kCUDA, &kvcompress_paged_attention_fused_metrics);
  ops.def(
      "kvcompress_paged_attention_fused_step("
      "    Tensor! out, Tensor! kv_metrics, Tensor query, Tensor key, Tensor value,"
      "    Tensor! key_cache, Tensor! value_cache, Tensor slot_mapping,"
      "    Tensor kv_metric_head_bias, int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor context_lens,"
      "    Tensor kv_position, Tensor last_position,"
      "    Tensor kv_metric_buffer_len, int block_size,"
      "    int max_context_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype) -> ()");
  ops.impl("kvcompress_paged_attention_fused_step", torch::kCUDA, &kvcompress_paged_attention_fused_step);
  ops.def(
      "kvcompress_paged_attention_fused_step_deferred("
      "    Tensor! out, Tensor! kv_metrics, Tensor query, Tensor key, Tensor value,"
      "    Tensor! key_cache, Tensor! value_cache, Tensor slot_mapping,"
      "    Tensor kv_metric_head_bias, Tensor! pend_kv, Tensor! pend_slot, Tensor? pend_dst,"
      "    int num_kv_heads, float scale,"
      "    Tensor block_tables, Tensor context_lens,"
      "    Tensor kv_position, Tensor last_position,"
      "    Tensor kv_metric_buffer_len, int block_size,"
      "    int max_context_len, Tensor? alibi_slopes,"
      "    str kv_cache_dtype) -> ()");
  ops.impl("kvcompress_paged_attention_fused_step_deferred", torch::kCUDA, &kvcompress_paged_attention_fused_step_deferred);
  // the reference wrapper calls torch.ops._C.schedule_cache_evictions although the op is
  // registered under _C_kvc_ops (vllm/_custom_ops.py:982 vs torch_bindings.cpp:372-386):
  // export it under both names so either resolves.
  ops.def(
      "schedule_cache_evictions(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices,"
      "                           Tensor! evicted_kv_count, Tensor! remaining_kv_count,"
      "                           Tensor evicted_kv_offsets, Tensor sorted_indices,"
      "                           Tensor seq_block_offsets, Tensor layer_by_block,"
      "                           Tensor head_by_block, Tensor virtual_block_num_by_block,"
      "                           Tensor evicted_blocks_per_seq, Tensor context_lens,"
      "                           Tensor hanging_token_count, Tensor kv_position,"
      "                           Tensor last_position, Tensor protected_window_size,"
      "                           int block_size, bool evict_evenly_per_layer,"
      "                           Tensor? control_layers, int max_evicted_kv,"
      "                           int null_eviction_index, bool truncate) -> ()");
  ops.impl("schedule_cache_evictions", torch::kCUDA, &schedule_cache_evictions);
}

TORCH_LIBRARY_FRAGMENT(_C_cache_ops, cache_ops) {
  cache_ops.def(
      "kvcompress_reshape_and_cache(Tensor key, Tensor value,"
      "                             Tensor! key_cache, Tensor! value_cache,"
      "                             Tensor kv_metrics, Tensor slot_mapping,"
      "                             Tensor kv_metric_head_bias,"
      "                             str kv_cache_dtype,"
      "                             float k_scale, float v_scale) -> ()");
  cache_ops.impl("kvcompress_reshape_and_cache", torch::kCUDA, &kvcompress_reshape_and_cache);
}

TORCH_LIBRARY_FRAGMENT(_C_kvc_ops, kvc_ops) {
  kvc_ops.def(
      "schedule_cache_evictions(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices,"
      "                           Tensor! evicted_kv_count, Tensor! remaining_kv_count,"
      "                           Tensor evicted_kv_offsets, Tensor sorted_indices,"
      "                           Tensor seq_block_offsets, Tensor layer_by_block,"
      "                           Tensor head_by_block, Tensor virtual_block_num_by_block,"
      "                           Tensor evicted_blocks_per_seq, Tensor context_lens,"
      "                           Tensor hanging_token_count, Tensor kv_position,"
      "                           Tensor last_position, Tensor protected_window_size,"
      "                           int block_size, bool evict_evenly_per_layer,"
      "                           Tensor? control_layers, int max_evicted_kv,"
      "                           int null_eviction_index, bool truncate) -> ()");
  kvc_ops.impl("schedule_cache_evictions", torch::kCUDA, &schedule_cache_evictions);
  kvc_ops.def(
      "truncate_cache_evictions(Tensor! evicted_kv_indices, Tensor! evicted_logical_indices,"
      "                           Tensor! evicted_kv_count, Tensor evicted_kv_offsets,"
      "                           Tensor hanging_token_count, int block_size,"
      "                           int max_evicted_kv, int null_eviction_index) -> ()");
  kvc_ops.impl("truncate_cache_evictions", torch::kCUDA, &truncate_cache_evictions);
  kvc_ops.def(
      "count_block_evictions(Tensor! evicted_block_count, Tensor! evicted_logical_indices,"
      "                        Tensor evicted_kv_offsets, Tensor hanging_token_count,"
      "                        int block_size, int null_value) -> ()");
  kvc_ops.impl("count_block_evictions", torch::kCUDA, &count_block_evictions);
  kvc_ops.def(
      "schedule_t1_cache_moves(Tensor! cache_moves_idx, Tensor! cache_moves_count,"
      "                         Tensor evicted_logical_indices, Tensor evicted_kv_count,"
      "                         Tensor evicted_kv_offsets, Tensor block_tables,"
      "                         Tensor context_lens, int block_size) -> ()");
  kvc_ops.impl("schedule_t1_cache_moves", torch::kCUDA, &schedule_t1_cache_moves);
  kvc_ops.def(
      "execute_cache_moves(Tensor! k_cache, Tensor! v_cache, Tensor! kv_metrics,"
      "                      Tensor! kv_position, Tensor cache_moves_idx,"
      "                      Tensor cache_moves_count, Tensor evicted_kv_offsets,"
      "                      int blocks_per_head, int threads_per_head) -> ()");
  kvc_ops.impl("execute_cache_moves", torch::kCUDA, &execute_cache_moves);
  kvc_ops.def(
      "execute_cache_moves_deferred(Tensor! k_cache, Tensor! v_cache, Tensor! kv_metrics,"
      "                             Tensor! kv_position, Tensor cache_moves_idx,"
      "                             Tensor cache_moves_count, Tensor evicted_kv_offsets,"
      "                             Tensor pend_kv, Tensor pend_slot, Tensor? pend_dst) -> ()");
  kvc_ops.impl("execute_cache_moves_deferred", torch::kCUDA, &execute_cache_moves_deferred);
  kvc_ops.def(
      "flush_deferred(Tensor! k_cache, Tensor! v_cache, Tensor pend_kv, Tensor! pend_slot,"
      "               Tensor? pend_dst) -> ()");
  kvc_ops.impl("flush_deferred", torch::kCUDA, &flush_deferred);
  // additive (not in the reference op list)
  kvc_ops.def(
      "select_evictions(Tensor! evicted_logical_indices, Tensor! evicted_kv_count,"
      "                  Tensor! evicted_block_count, Tensor kv_metrics, Tensor kv_position,"
      "                  Tensor block_tables, Tensor context_lens, Tensor hanging_token_count,"
      "                  Tensor evicted_kv_offsets, Tensor evicted_blocks_per_seq,"
      "                  Tensor last_position, Tensor protected_window_size,"
      "                  int block_size, int null_eviction_index, Tensor? head_bias=None,"
      "                  Tensor? position_bins=None, float bias_weight=0.0, bool use_average=False,"
      "                  int num_sinks=0) -> ()");
  kvc_ops.impl("select_evictions", torch::kCUDA, &select_evictions);
  kvc_ops.def(
      "prepare_compression(Tensor context_lens, Tensor! hanging_token_count, Tensor! evicted_kv_offsets,"
      "                     Tensor! evicted_blocks_per_seq, int block_size, int max_cache_blocks_per_seq) -> ()");
  kvc_ops.impl("prepare_compression", torch::kCUDA, &prepare_compression);
  kvc_ops.def(
      "free_and_append_slots(Tensor! context_lens, Tensor block_tables, Tensor? evicted_block_count,"
      "                       Tensor? hanging_token_count, Tensor seq_len, Tensor! slot_mapping,"
      "                       Tensor! kv_position, int block_size, bool append, Tensor(a!)? status=None) -> ()");
  kvc_ops.impl("free_and_append_slots", torch::kCUDA, &free_and_append_slots);
  kvc_ops.def(
      "prefill_kv_metrics(Tensor(a!)? kv_metric_out, Tensor(b!)? kv_metrics, Tensor? slot_mapping, Tensor query,"
      "                    Tensor key, Tensor seq_start_loc, Tensor kv_metric_buffer_len, int max_prompt_len,"
      "                    float scale, int n_observed, bool use_l2, bool use_average, bool use_maxpool) -> ()");
  kvc_ops.impl("prefill_kv_metrics", torch::kCUDA, &prefill_kv_metrics);
  kvc_ops.def("aggregate_prefill(Tensor! kv_metrics, Tensor prefill_metrics, Tensor slot_mapping) -> ()");
  kvc_ops.impl("aggregate_prefill", torch::kCUDA, &aggregate_prefill);
  kvc_ops.def("alloc_init(Tensor! free_mask, Tensor! free_count, Tensor! status) -> ()");
  kvc_ops.impl("alloc_init", torch::kCUDA, &alloc_init);
  kvc_ops.def(
      "alloc_append_slots(Tensor! free_mask, Tensor! free_count, Tensor! status, Tensor! block_tables,"
      "                    Tensor! context_lens, Tensor? seq_slots, Tensor seq_len, Tensor! slot_mapping,"
      "                    Tensor! kv_position, Tensor(a!)? seq_index_by_block, Tensor(b!)? layer_index_by_block,"
      "                    Tensor(c!)? head_index_by_block, Tensor(d!)? logical_block_num_by_block, int block_size) -> ()");
  kvc_ops.impl("alloc_append_slots", torch::kCUDA, &alloc_append_slots);
  kvc_ops.def(
      "alloc_add_sequences(Tensor! free_mask, Tensor! free_count, Tensor! status, Tensor! block_tables,"
      "                     Tensor! context_lens, Tensor seq_slots, Tensor seq_lens, int max_seq_len,"
      "                     Tensor! kv_position, Tensor(a!)? seq_index_by_block, Tensor(b!)? layer_index_by_block,"
      "                     Tensor(c!)? head_index_by_block, Tensor(d!)? logical_block_num_by_block, int block_size) -> ()");
  kvc_ops.impl("alloc_add_sequences", torch::kCUDA, &alloc_add_sequences);
  kvc_ops.def(
      "alloc_free_compressed(Tensor! free_mask, Tensor! free_count, Tensor block_tables, Tensor! context_lens,"
      "                       Tensor evicted_block_count, Tensor? seq_slots, Tensor(a!)? seq_index_by_block,"
      "                       int block_size) -> ()");
  kvc_ops.impl("alloc_free_compressed", torch::kCUDA, &alloc_free_compressed);
  kvc_ops.def(
      "alloc_remove_sequences(Tensor! free_mask, Tensor! free_count, Tensor block_tables, Tensor! context_lens,"
      "                        Tensor seq_slots, Tensor(a!)? seq_index_by_block, int block_size) -> ()");
  kvc_ops.impl("alloc_remove_sequences", torch::kCUDA, &alloc_remove_sequences);
}

// importable as a Python module named _C, like the reference's REGISTER_EXTENSION
// (csrc/core/registration.h:17-27)
PyMODINIT_FUNC PyInit__C() {
  static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_C", nullptr, 0, nullptr};
  return PyModule_Create(&module);
}
