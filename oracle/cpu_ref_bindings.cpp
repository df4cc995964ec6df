// TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by the product path).
//
// Registers the *reference's own CPU kernels* -- vLLM's CPU backend, csrc/cpu/attention.cpp and
// csrc/cpu/cache.cpp, compiled unmodified from where they lie under /root/reference/csrc with the flags of
// cmake/cpu_extension.cmake:52-69 (see Makefile.ref) -- under the torch library namespace `_vllm_cpu_ref`.
// The fork has no CPU implementation of any KV-Compress op (csrc/cpu/torch_bindings.cpp registers none), so what
// exists of "the reference's CPU path" for this hot path is the vanilla paged-attention decode + cache write;
// bench.py times it next to the full-path oracle port (cpu_baseline) as `cpu_reference_kernels`.
// Schemas restate csrc/cpu/torch_bindings.cpp (paged_attention_v1, reshape_and_cache).
#include <torch/all.h>
#include <torch/library.h>

void paged_attention_v1(torch::Tensor& out, torch::Tensor& query, torch::Tensor& key_cache,
                        torch::Tensor& value_cache, int64_t num_kv_heads, double scale,
                        torch::Tensor& block_tables, torch::Tensor& seq_lens, int64_t block_size,
                        int64_t max_seq_len, const c10::optional<torch::Tensor>& alibi_slopes,
                        const std::string& kv_cache_dtype, double k_scale, double v_scale, const int64_t tp_rank,
                        const int64_t blocksparse_local_blocks, const int64_t blocksparse_vert_stride,
                        const int64_t blocksparse_block_size, const int64_t blocksparse_head_sliding_step);
void reshape_and_cache(torch::Tensor& key, torch::Tensor& value, torch::Tensor& key_cache,
                       torch::Tensor& value_cache, torch::Tensor& slot_mapping, const std::string& kv_cache_dtype,
                       double k_scale, double v_scale);

TORCH_LIBRARY(_vllm_cpu_ref, m) {
  m.def("paged_attention_v1(Tensor! out, Tensor query, Tensor key_cache, Tensor value_cache, int num_kv_heads,"
        " float scale, Tensor block_tables, Tensor seq_lens, int block_size, int max_seq_len, Tensor? alibi_slopes,"
        " str kv_cache_dtype, float k_scale, float v_scale, int tp_rank, int blocksparse_local_blocks,"
        " int blocksparse_vert_stride, int blocksparse_block_size, int blocksparse_head_sliding_step) -> ()");
  m.impl("paged_attention_v1", torch::kCPU, &paged_attention_v1);
  m.def("reshape_and_cache(Tensor key, Tensor value, Tensor! key_cache, Tensor! value_cache, Tensor slot_mapping,"
        " str kv_cache_dtype, float k_scale, float v_scale) -> ()");
  m.impl("reshape_and_cache", torch::kCPU, &reshape_and_cache);
}
