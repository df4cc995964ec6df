// K6 kvcompress_reshape_and_cache: per-head scatter of new K/V into the paged cache
// plus metric initialisation to the head bias.
// reference: csrc/kvcompress_cache_kernels.cu:27-89 (kernel), :111-139 (launcher)
//
// One warp per (token, kv head).  The slot differs per head (slot_mapping [T, H]),
// so every head of a token lands in a different physical block.  K is written as
// 16-byte chunks (x elements), V element-wise at stride B (that is the layout's cost:
// one 2-byte store per sector; decode writes T*H*D elements per layer, a few MB).
#include "kvc_common.cuh"

namespace kvc {

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
  if (lane == 0) kv_metrics[slot] = kv_metric_head_bias[h];  // reference :55-57
  const int64_t blk = slot / B;
  const int o = (int)(slot % B);
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
  const int warps = num_tokens * num_heads;
  const dim3 grid(ceil_div(warps, 4)), block(128);
  cudaStream_t st = as_stream(stream);
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
