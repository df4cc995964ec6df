// Block-state device ops around a compression cycle (additive: the reference does these with
// ~30 small torch launches and boolean-mask copies of the [L,S,H,MB] tables, with host syncs).
//
//   prepare_compression      hanging_token_count, evicted_kv_offsets, evicted_blocks_per_seq
//       reference: vllm/kvcompress/block.py:330-335 (hanging = ctx % B or B),
//                  vllm/kvcompress/scheduler.py:274-280 (offsets = exclusive cumsum of ceil(ctx/B)*B
//                  in (s,l,h) order), :131-141 (budget = blocks above max_cache_tokens*L*H/B)
//   free_and_append_slots    context_lens -= freed; then +1 token per head, slot mapping, position
//       reference: vllm/kvcompress/block.py:184-210 (free trailing blocks: ctx -= max(nb*B-(B-hanging),0)),
//                  vllm/kvcompress/block_manager.py:269-294 + block.py:275-321 (batch_append_slots:
//                  slot = table[ctx / B] * B + ctx % B, token position = sequence length)
// Pure int32 element work over [S,L,H] (65 536 heads at batch 256): two/one launches instead of ~30.
#include "kvc_common.cuh"

namespace kvc {

// one CTA per sequence: per-head block counts, exclusive scan inside the sequence, sequence total
__global__ void __launch_bounds__(256) prepare_seq_kernel(const int* __restrict__ context_lens,  // [L,S,H]
                                                         int* __restrict__ hanging,             // [S,L,H]
                                                         int* __restrict__ offsets,             // [S,L,H] (local)
                                                         int* __restrict__ seq_slots,           // [S] allocated slots
                                                         int* __restrict__ budgets,             // [S] or null
                                                         int S, int L, int H, int B, int max_cache_blocks) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int s = blockIdx.x, t = threadIdx.x, lane = t % kWarp, w = t / kWarp;
  const int LH = L * H;
  if (t == 0) s_carry = 0;
  __syncthreads();
  for (int h0 = 0; h0 < LH; h0 += 256) {
    const int hd = h0 + t;
    int alloc = 0;
    if (hd < LH) {
      const int l = hd / H, h = hd % H;
      const int ctx = context_lens[(l * S + s) * H + h];
      const int r = ctx % B;
      hanging[s * LH + hd] = (r == 0) ? B : r;
      alloc = (ctx + B - 1) / B * B;
    }
    int incl = alloc;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == kWarp - 1) s_warp[w] = incl;
    __syncthreads();
    int base = s_carry;
    for (int q = 0; q < w; ++q) base += s_warp[q];
    if (hd < LH) offsets[s * LH + hd] = base + incl - alloc;
    __syncthreads();
    if (t == 255) s_carry = base + incl;
    __syncthreads();
  }
  if (t == 0) {
    seq_slots[s] = s_carry;
    if (budgets) budgets[s] = max(0, s_carry / B - max_cache_blocks);
  }
}

// add the exclusive prefix of the per-sequence totals
__global__ void __launch_bounds__(256) prepare_offsets_kernel(int* __restrict__ offsets, const int* __restrict__ seq_slots,
                                                             int S, int LH) {
  __shared__ int s_red[8];
  const int s = blockIdx.x, t = threadIdx.x, lane = t % kWarp, w = t / kWarp;
  int acc = 0;
  for (int i = t; i < s; i += 256) acc += seq_slots[i];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) s_red[w] = acc;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) base += s_red[q];
  for (int hd = t; hd < LH; hd += 256) offsets[s * LH + hd] += base;
}

__global__ void __launch_bounds__(256) free_and_append_kernel(int* __restrict__ context_lens,         // [L,S,H]
                                                             const int* __restrict__ block_tables,   // [L,S,H,MB]
                                                             const int* __restrict__ evicted_blocks, // [S,L,H] or null
                                                             const int* __restrict__ hanging,        // [S,L,H]
                                                             const int* __restrict__ seq_len,        // [S]
                                                             int64_t* __restrict__ slot_mapping,     // [L,S,H]
                                                             int* __restrict__ kv_position,          // [NB,B]
                                                             int S, int L, int H, int MB, int B, int append,
                                                             int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // flat (l, s, h)
  if (i >= L * S * H) return;
  const int l = i / (S * H), s = (i / H) % S, h = i % H;
  int ctx = context_lens[i];
  if (evicted_blocks) {
    const int g = (s * L + l) * H + h;
    const int freed = max(evicted_blocks[g] * B - (B - hanging[g]), 0);  // block.py:184-210
    ctx -= freed;
  }
  if (append && ctx / B >= MB) {
    // block table full: the reference would raise an index error (block.py:303-321); nothing is appended
    if (status) *status = 2;
    slot_mapping[i] = -1;
  } else if (append) {
    const int blk = block_tables[(int64_t)i * MB + ctx / B];
    const int64_t slot = (int64_t)blk * B + ctx % B;
    slot_mapping[i] = slot;
    kv_position[slot] = seq_len[s];  // position of the token being appended
    ctx += 1;
  }
  context_lens[i] = ctx;
}

}  // namespace kvc

using namespace kvc;

extern "C" int kvc_prepare_compression(const int* context_lens, int* hanging_token_count, int* evicted_kv_offsets,
                                       int* evicted_blocks_per_seq, int* seq_slots_workspace, int num_seqs,
                                       int num_layers, int num_kv_heads, int block_size, int max_cache_blocks_per_seq,
                                       kvc_stream_t stream) {
  KVC_REQUIRE(context_lens && hanging_token_count && evicted_kv_offsets && seq_slots_workspace, KVC_ERR_INVALID,
              "prepare_compression: null tensor");
  KVC_REQUIRE(num_seqs >= 0 && num_layers > 0 && num_kv_heads > 0 && block_size > 0, KVC_ERR_INVALID,
              "prepare_compression: bad sizes");
  if (num_seqs == 0) return KVC_OK;
  cudaStream_t st = as_stream(stream);
  prepare_seq_kernel<<<num_seqs, 256, 0, st>>>(context_lens, hanging_token_count, evicted_kv_offsets,
                                               seq_slots_workspace, evicted_blocks_per_seq, num_seqs, num_layers,
                                               num_kv_heads, block_size, max_cache_blocks_per_seq);
  KVC_CHECK_LAUNCH("prepare_seq_kernel");
  prepare_offsets_kernel<<<num_seqs, 256, 0, st>>>(evicted_kv_offsets, seq_slots_workspace, num_seqs,
                                                   num_layers * num_kv_heads);
  KVC_CHECK_LAUNCH("prepare_offsets_kernel");
  return KVC_OK;
}

extern "C" int kvc_free_and_append_slots(int* context_lens, const int* block_tables, const int* evicted_block_count,
                                         const int* hanging_token_count, const int* seq_len, int64_t* slot_mapping,
                                         int* kv_position, int num_seqs, int num_layers, int num_kv_heads,
                                         int max_num_blocks_per_seq, int block_size, int append, int* status,
                                         kvc_stream_t stream) {
  KVC_REQUIRE(context_lens && block_tables, KVC_ERR_INVALID, "free_and_append_slots: null tensor");
  KVC_REQUIRE(!evicted_block_count || hanging_token_count, KVC_ERR_INVALID,
              "free_and_append_slots: hanging_token_count required with evicted_block_count");
  KVC_REQUIRE(!append || (seq_len && slot_mapping && kv_position), KVC_ERR_INVALID,
              "free_and_append_slots: append needs seq_len, slot_mapping and kv_position");
  const int total = num_seqs * num_layers * num_kv_heads;
  if (total <= 0) return KVC_OK;
  free_and_append_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
      context_lens, block_tables, evicted_block_count, hanging_token_count, seq_len, slot_mapping, kv_position,
      num_seqs, num_layers, num_kv_heads, max_num_blocks_per_seq, block_size, append ? 1 : 0, status);
  KVC_CHECK_LAUNCH("free_and_append_kernel");
  return KVC_OK;
}
