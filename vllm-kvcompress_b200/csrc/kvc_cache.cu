// K6 kvcompress_reshape_and_cache: per-head scatter of new K/V into the paged cache
// plus metric initialisation to the head bias.
// reference: csrc/kvcompress_cache_kernels.cu:27-89 (kernel), :111-139 (launcher)
//
// One warp per (token, kv head).  The slot differs per head (slot_mapping [T, H]),
// so every head of a token lands in a different physical block.  K is written as
// 16-byte chunks (x elements), V element-wise at stride B (that is the layout's cost:
// one 2-byte store per sector; decode writes T*H*D elements per layer, a few MB).
//
// Prefill writes whole blocks: when the 16 tokens t0 .. t0+15 of a head map to the 16 slots of one block in order
// (what the block manager hands out for a prompt), the warp of the block's FIRST token writes the block and the
// other fifteen retire: K as 256 coalesced 16-byte chunks, V transposed through shared memory into the block's
// 128 full 32-byte rows -- 4 KB of full-sector stores instead of 2048 two-byte stores into 128 sectors that
// sixteen warps each touch once (0.158 -> see DESIGN.md section 5 for 4 x 6000-token prompts).  Anything else (decode
// tokens, partial blocks, padding, other dtypes / shapes) takes the element-wise path.
#include "kvc_common.cuh"

namespace kvc {

// the 16 tokens tl .. tl+15 of head h fill block blk in order: K as 256 coalesced 16-byte chunks, V transposed
// through the warp's shared-memory tile into the block's 128 full 32-byte rows (fp16 / bf16, D = 128, x = 8)
__device__ __forceinline__ void write_whole_block(const uint16_t* __restrict__ key, const uint16_t* __restrict__ value,
                                                  uint16_t* __restrict__ key_cache, uint16_t* __restrict__ value_cache,
                                                  float* __restrict__ kv_metrics,
                                                  const float* __restrict__ kv_metric_head_bias, uint16_t* tile,
                                                  int tl, int h, int64_t blk, int64_t key_stride,
                                                  int64_t value_stride, int lane) {
  const uint16_t* __restrict__ k16 = key + (int64_t)h * 128;
  const uint16_t* __restrict__ v16 = value + (int64_t)h * 128;
  uint16_t* __restrict__ kcb = key_cache + blk * 2048;
  uint16_t* __restrict__ vcb = value_cache + blk * 2048;
  if (lane < 16) kv_metrics[blk * 16 + lane] = kv_metric_head_bias[h];
  // V: lane = (token, half of the head dimension), 8 x 16 B loads, 64 two-byte shared stores
  const int tok = lane & 15, half = lane >> 4;
  uint4 vv[8];
#pragma unroll
  for (int q = 0; q < 8; ++q)
    vv[q] = *reinterpret_cast<const uint4*>(v16 + (int64_t)(tl + tok) * value_stride + half * 64 + q * 8);
  // K: item i = (chunk j, token): source 16 B at token stride, destination consecutive
  uint4 kk[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int i = it * 32 + lane;
    kk[it] = *reinterpret_cast<const uint4*>(k16 + (int64_t)(tl + (i & 15)) * key_stride + (i >> 4) * 8);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t w4[4] = {vv[q].x, vv[q].y, vv[q].z, vv[q].w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int d = half * 64 + q * 8 + e;
      tile[d * 24 + tok] = (uint16_t)(w4[e >> 1] >> ((e & 1) * 16));
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) *reinterpret_cast<uint4*>(kcb + (it * 32 + lane) * 8) = kk[it];
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int i = it * 32 + lane, d = i >> 1, part = i & 1;
    *reinterpret_cast<uint4*>(vcb + d * 16 + part * 8) = *reinterpret_cast<const uint4*>(tile + d * 24 + part * 8);
  }
  __syncwarp();
}

// one token of one head, element-wise (the reference's path)
template <typename T>
__device__ __forceinline__ void write_one_token(const T* __restrict__ key, const T* __restrict__ value,
                                                T* __restrict__ key_cache, T* __restrict__ value_cache,
                                                float* __restrict__ kv_metrics,
                                                const float* __restrict__ kv_metric_head_bias, int t, int h,
                                                int64_t slot, int D, int B, int x, int64_t key_stride,
                                                int64_t value_stride, int lane) {
  const int64_t blk = slot / B;
  const int o = (int)(slot % B);
  if (lane == 0) kv_metrics[slot] = kv_metric_head_bias[h];  // reference :55-57
  const T* __restrict__ kp = key + (int64_t)t * key_stride + (int64_t)h * D;
  const T* __restrict__ vp = value + (int64_t)t * value_stride + (int64_t)h * D;
  T* __restrict__ kc = key_cache + blk * (int64_t)B * D;
  T* __restrict__ vc = value_cache + blk * (int64_t)B * D;
  constexpr int kVecBytes = 16;
  if (x * (int)sizeof(T) == kVecBytes && (D % x) == 0 &&
      ((reinterpret_cast<uintptr_t>(kp) | reinterpret_cast<uintptr_t>(kc)) & 15) == 0) {
    for (int j = lane; j < D / x; j += kWarp) {  // one 16 B chunk per lane
      const uint4 v = *reinterpret_cast<const uint4*>(kp + j * x);
      *reinterpret_cast<uint4*>(kc + (int64_t)j * B * x + o * x) = v;  // reference :69-72
    }
  } else {
    for (int e = lane; e < D; e += kWarp)
      kc[(int64_t)(e / x) * B * x + o * x + e % x] = kp[e];
  }
  for (int e = lane; e < D; e += kWarp) vc[(int64_t)e * B + o] = vp[e];  // reference :73-75
}

// Prefill batches: one warp per (window of 16 batch tokens, head).  The warp reads the slots of the 48 tokens
// around its window once, finds the blocks that the batch fills completely and in order (bit mask of "slot =
// previous slot + 1"), writes the one such block that STARTS in its window as a whole and the window's remaining
// tokens (partial tail blocks, padding) element-wise.  Sixteen times fewer warps than one per (token, head), and
// no warp that only finds out it has nothing to do.
__global__ void __launch_bounds__(128) reshape_and_cache_window_kernel(
    const uint16_t* __restrict__ key, const uint16_t* __restrict__ value, uint16_t* __restrict__ key_cache,
    uint16_t* __restrict__ value_cache, float* __restrict__ kv_metrics, const int64_t* __restrict__ slot_mapping,
    const float* __restrict__ kv_metric_head_bias, int num_tokens, int H, int64_t key_stride, int64_t value_stride) {
  __shared__ __align__(16) uint16_t sv[4][128 * 24];
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  const int nwin = (num_tokens + 15) / 16;
  if (w >= nwin * H) return;
  const int g = w / H, h = w % H;
  const int t_base = g * 16;
  // S_u = slot of token t_base - 16 + u, u = 0 .. 47 (own window: u = 16 .. 31 = lanes 16 .. 31 of A)
  const int ta = t_base - 16 + lane, tb = t_base + 16 + lane;
  const int64_t A = (ta >= 0 && ta < num_tokens) ? slot_mapping[(int64_t)ta * H + h] : -1;
  const int64_t Bv = (lane < 16 && tb < num_tokens) ? slot_mapping[(int64_t)tb * H + h] : -1;
  const int64_t Aprev = __shfl_up_sync(0xffffffffu, A, 1);
  const int64_t A31 = __shfl_sync(0xffffffffu, A, 31);
  int64_t Bprev = __shfl_up_sync(0xffffffffu, Bv, 1);
  if (lane == 0) Bprev = A31;
  const unsigned cA = __ballot_sync(0xffffffffu, lane >= 1 && A >= 0 && Aprev >= 0 && A == Aprev + 1);
  const unsigned cB = __ballot_sync(0xffffffffu, lane < 16 && Bv >= 0 && Bprev >= 0 && Bv == Bprev + 1);
  const unsigned long long cons = (unsigned long long)cA | ((unsigned long long)cB << 32);
  const bool own = lane >= 16 && A >= 0;
  const int o = own ? (int)(A % 16) : 0;
  const int u0 = lane - o;   // where the token's block would have to start
  const bool covered = own && ((cons >> (u0 + 1)) & 0x7fffull) == 0x7fffull;
  const unsigned lead = __ballot_sync(0xffffffffu, covered && o == 0);
  const unsigned rest = __ballot_sync(0xffffffffu, own && !covered);
  if (lead) {   // at most one whole block starts in a 16-token window
    const int ul = __ffs(lead) - 1;
    const int64_t sl = __shfl_sync(0xffffffffu, A, ul);
    write_whole_block(key, value, key_cache, value_cache, kv_metrics, kv_metric_head_bias, sv[threadIdx.x / kWarp],
                      t_base - 16 + ul, h, sl / 16, key_stride, value_stride, lane);
  }
  // the window's remaining tokens (partial tail blocks), four at a time with all loads in front of the stores: a
  // prompt's last block would otherwise be up to fifteen dependent load -> store rounds in one warp
  const int half = lane >> 4, j = lane & 15;
  for (unsigned m = rest; m;) {
    int u[4];
    int64_t sl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      u[q] = m ? __ffs(m) - 1 : -1;
      m &= m - 1;   // 0 stays 0
      sl[q] = __shfl_sync(0xffffffffu, A, u[q] & 31);
    }
    uint4 kk[2];
    uint2 vv[4];
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {   // K: lanes 0-15 one token, lanes 16-31 the next; lane = 16-byte chunk
      const int uq = half ? u[2 * ps + 1] : u[2 * ps];
      kk[ps] = make_uint4(0, 0, 0, 0);
      if (uq >= 0)
        kk[ps] = *reinterpret_cast<const uint4*>(key + (int64_t)(t_base - 16 + uq) * key_stride + h * 128 + j * 8);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {      // V: lane = 4 consecutive elements of the token
      vv[q] = make_uint2(0, 0);
      if (u[q] >= 0)
        vv[q] = *reinterpret_cast<const uint2*>(value + (int64_t)(t_base - 16 + u[q]) * value_stride + h * 128 + lane * 4);
    }
#pragma unroll
    for (int ps = 0; ps < 2; ++ps) {
      const int uq = half ? u[2 * ps + 1] : u[2 * ps];
      const int64_t sq = half ? sl[2 * ps + 1] : sl[2 * ps];
      if (uq >= 0) *reinterpret_cast<uint4*>(key_cache + (sq / 16) * 2048 + j * 128 + (sq % 16) * 8) = kk[ps];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (u[q] < 0) continue;
      uint16_t* __restrict__ vcb = value_cache + (sl[q] / 16) * 2048 + (sl[q] % 16);
      vcb[(lane * 4 + 0) * 16] = (uint16_t)vv[q].x;
      vcb[(lane * 4 + 1) * 16] = (uint16_t)(vv[q].x >> 16);
      vcb[(lane * 4 + 2) * 16] = (uint16_t)vv[q].y;
      vcb[(lane * 4 + 3) * 16] = (uint16_t)(vv[q].y >> 16);
      if (lane == 0) kv_metrics[sl[q]] = kv_metric_head_bias[h];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(128) reshape_and_cache_kernel(
    const T* __restrict__ key,    // [T, H, D] token stride key_stride
    const T* __restrict__ value,  // [T, H, D] token stride value_stride
    T* __restrict__ key_cache,    // [NB, D/x, B, x]
    T* __restrict__ value_cache,  // [NB, D, B]
    float* __restrict__ kv_metrics,                 // [NB, B]
    const int64_t* __restrict__ slot_mapping,       // [T, H]
    const float* __restrict__ kv_metric_head_bias,  // [H]
    int num_tokens, int H, int D, int B, int x, int64_t key_stride, int64_t value_stride) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (w >= num_tokens * H) return;
  const int t = w / H, h = w % H;
  const int64_t slot = slot_mapping[w];
  if (slot < 0) return;  // padding token, reference :49-52
  const int64_t blk = slot / B;
  const int o = (int)(slot % B);
  if (sizeof(T) == 2 && B == 16 && D == 128 && x == 8 && num_tokens >= 16 &&
      ((reinterpret_cast<uintptr_t>(key) | reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(key_cache) |
        reinterpret_cast<uintptr_t>(value_cache)) & 15) == 0 &&
      ((key_stride | value_stride) & 7) == 0) {
    // is this token part of a block that the batch fills completely and in order?
    const int tl = t - o;   // the token that would open the block
    bool ok = true;
    if (lane < 16) {
      const int ti = tl + lane;
      ok = ti >= 0 && ti < num_tokens && slot_mapping[(int64_t)ti * H + h] == slot - o + lane;
    }
    if (__all_sync(0xffffffffu, ok)) {
      if (o != 0) return;   // written by the warp of the block's first token
      __shared__ __align__(16) uint16_t sv[4][128 * 24];   // per warp: V^T, rows of 16 tokens padded to 48 B
      write_whole_block(reinterpret_cast<const uint16_t*>(key), reinterpret_cast<const uint16_t*>(value),
                        reinterpret_cast<uint16_t*>(key_cache), reinterpret_cast<uint16_t*>(value_cache), kv_metrics,
                        kv_metric_head_bias, sv[threadIdx.x / kWarp], tl, h, blk, key_stride, value_stride, lane);
      return;
    }
  }
  write_one_token<T>(key, value, key_cache, value_cache, kv_metrics, kv_metric_head_bias, t, h, slot, D, B, x,
                     key_stride, value_stride, lane);
}

}  // namespace kvc

using namespace kvc;

extern "C" int kvc_reshape_and_cache(const void* key, const void* value, void* key_cache,
                                     void* value_cache, float* kv_metrics,
                                     const int64_t* slot_mapping, const float* kv_metric_head_bias,
                                     int num_tokens, int num_heads, int head_size, int block_size,
                                     int x, int64_t key_stride, int64_t value_stride, int dtype,
                                     kvc_stream_t stream) {
  KVC_REQUIRE(key && value && key_cache && value_cache && kv_metrics && slot_mapping &&
                  kv_metric_head_bias,
              KVC_ERR_INVALID, "kvcompress_reshape_and_cache: null tensor");
  KVC_REQUIRE(num_tokens >= 0 && num_heads > 0 && head_size > 0 && block_size > 0 && x > 0 &&
                  head_size % x == 0,
              KVC_ERR_INVALID, "kvcompress_reshape_and_cache: bad sizes");
  if (num_tokens == 0) return KVC_OK;
  cudaStream_t st = as_stream(stream);
  // prefill-sized fp16 / bf16 batches in the production layout: one warp per (16-token window, head)
  if ((dtype == KVC_F16 || dtype == KVC_BF16) && num_tokens >= 512 && head_size == 128 && block_size == 16 && x == 8 &&
      ((reinterpret_cast<uintptr_t>(key) | reinterpret_cast<uintptr_t>(value) |
        reinterpret_cast<uintptr_t>(key_cache) | reinterpret_cast<uintptr_t>(value_cache)) & 15) == 0 &&
      ((key_stride | value_stride) & 7) == 0) {
    const int wwarps = ceil_div(num_tokens, 16) * num_heads;
    reshape_and_cache_window_kernel<<<ceil_div(wwarps, 4), 128, 0, st>>>(
        reinterpret_cast<const uint16_t*>(key), reinterpret_cast<const uint16_t*>(value),
        reinterpret_cast<uint16_t*>(key_cache), reinterpret_cast<uint16_t*>(value_cache), kv_metrics, slot_mapping,
        kv_metric_head_bias, num_tokens, num_heads, key_stride, value_stride);
    KVC_CHECK_LAUNCH("reshape_and_cache_window_kernel");
    return KVC_OK;
  }
  const int warps = num_tokens * num_heads;
  const dim3 grid(ceil_div(warps, 4)), block(128);
#define KVC_LAUNCH_K6(T)                                                                        \
  reshape_and_cache_kernel<T><<<grid, block, 0, st>>>(                                          \
      reinterpret_cast<const T*>(key), reinterpret_cast<const T*>(value),                       \
      reinterpret_cast<T*>(key_cache), reinterpret_cast<T*>(value_cache), kv_metrics,           \
      slot_mapping, kv_metric_head_bias, num_tokens, num_heads, head_size, block_size, x,       \
      key_stride, value_stride)
  switch (dtype) {
    case KVC_F16:
    case KVC_BF16: KVC_LAUNCH_K6(uint16_t); break;
    case KVC_F32: KVC_LAUNCH_K6(uint32_t); break;
    case KVC_U8: KVC_LAUNCH_K6(uint8_t); break;
    default: KVC_FAIL(KVC_ERR_UNSUPPORTED, "Unsupported data type: %d", dtype);
  }
#undef KVC_LAUNCH_K6
  KVC_CHECK_LAUNCH("reshape_and_cache_kernel");
  return KVC_OK;
}
