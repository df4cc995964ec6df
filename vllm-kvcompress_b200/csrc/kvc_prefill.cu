// Prefill KV-metric initialisation on the sm_100a tensor cores (tcgen05 + TMEM).
//
// replaces _naive_kvc_attention / _naive_kvc_masked_attention
//   (vllm/attention/backends/flash_attn.py:1122-1211, call site :993-1007) and
//   CompressionMetrics.aggregate_prefill (vllm/kvcompress/metrics.py:395-427).
// For every prompt: the last n_obs = min(prompt_len, n_observed) queries attend causally to all
// prompt keys; per query head  metric[k] = sum_q mask(q,k) * softmax(scale * q.k)^2  (use_l2),
// mask = key at least kv_metric_buffer_len tokens behind the query; optional (k+1)/n_obs rescale
// (use_average); max-pool of width 7 along the keys (use_maxpool); output [T, Hq] f32.
// aggregate_prefill then does  kv_metrics[slot_mapping[t,h]] += sum over the head's Q query heads.
//
// This is the one dense contraction of the path that is not a single-query product: per KV head
// the GQA group's Q x n_obs query rows form a 128-row operand, so one UMMA tile
//   S[128 rows x 128 keys] = Qtile[128 x D=128] . Ktile[128 x 128]^T
// is issued as 8 tcgen05.mma (M=128, N=128, K=16, fp16/bf16 in, fp32 accumulators in TMEM) by one
// thread; the four warps read their 32 TMEM lanes back with tcgen05.ld (one thread = one query
// row, so the row max / row sum of the softmax need no shuffles).  Softmax needs the row
// statistics over ALL keys before a probability can be squared, so the work is two sweeps:
//   pass 1 (stats):   per (sequence, kv head, 128-row tile, key split): running (max, sum exp)
//   pass 2 (metrics): merge the splits' statistics, recompute S, p = exp(s - M) / L, p^2, mask,
//                     column sums per query head by a warp reduce-scatter (31 shuffles per 32 keys)
//   finalize:         max-pool 7 + [T, Hq] output (+ fused aggregate_prefill)
// K tiles are re-read from L2 in pass 2 (a prompt's keys of one head: P x 256 B).  Both operands
// are staged in shared memory in the canonical no-swizzle K-major core-matrix layout
// (8 rows x 16 B contiguous; LBO = 2048 B between K-adjacent cores, SBO = 128 B between row-adjacent
// cores -- the descriptor convention verified on the B200 by tools/tcgen05_probe.cu); the K tile
// and the TMEM accumulator are double-buffered so the MMAs of tile t+1 run under the softmax
// epilogue of tile t.
//
// Arithmetic mirrors the reference: the fp16/bf16 einsum result is rounded to the input type before
// the fp32 scale (flash_attn.py:1178: `scale * einsum(...).float()`), softmax / squares / sums fp32.
#include "kvc_common.cuh"

namespace kvc {

constexpr int kTileKeys = 128;   // UMMA N
constexpr int kTileRows = 128;   // UMMA M
constexpr int kHeadDim = 128;    // UMMA K total (8 steps of 16)
constexpr int kTileBytes = kTileRows * kHeadDim * 2;  // 32 KB per operand tile

struct PrefillParams {
  const void* query;       // [T, Hq, D]
  const void* key;         // [T, H, D]
  const int* seq_start;    // [num_seqs + 1] cumulative prompt lengths
  const int* buffer_len;   // [num_seqs] kv_metric_buffer_len
  float2* stats;           // [num_seqs, H * mtiles, max_splits, 128] (max, sum)
  float* raw;              // [Hq, T] column sums before the max-pool
  int64_t q_stride, k_stride;  // token strides in elements
  int num_seqs, T, H, Q, n_observed, np, qt, mtiles, tiles_per_split, max_splits;
  int nchunks;             // 128-row query tiles per query head (n_observed > 128: np = 128, qt = 1)
  float scale;
  int use_l2, use_average;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major (tools/tcgen05_probe.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // sm_100 descriptor version
  return d;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool BF16>
__device__ __forceinline__ float round_to_input_type(float x) {
  if (BF16) return __bfloat162float(__float2bfloat16_rn(x));
  return __half2float(__float2half_rn(x));
}

// Stage a [128 rows x 128 d] tile into the canonical K-major core-matrix layout:
//   byte address(r, d) = (d/8)*2048 + (r/8)*128 + (r%8)*16 + (d%8)*2.
// A quarter-warp writes 8 consecutive rows of one 16 B chunk (128 contiguous bytes: conflict-free);
// the four quarter-warps take four consecutive chunks, so every row is read 64 B at a time.
// row_ptr(r) returns the global address of the row's 128 elements or nullptr (zero fill).
template <typename RowPtr>
__device__ __forceinline__ void stage_tile(unsigned char* smem_tile, RowPtr row_ptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const int combo = it * 4 + warp;
    const int r = (combo & 15) * 8 + (lane & 7);
    const int c = (combo >> 4) * 4 + (lane >> 3);
    const unsigned char* src = row_ptr(r);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (src) v = *reinterpret_cast<const uint4*>(src + c * 16);
    *reinterpret_cast<uint4*>(smem_tile + c * 2048 + r * 16) = v;
  }
}

// PASS 1: row statistics.  PASS 2: masked squared-probability column sums.
template <int PASS, bool BF16>
__global__ void __launch_bounds__(128) prefill_metrics_kernel(PrefillParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sQ = smem;
  unsigned char* sK0 = smem + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTileBytes);       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 3 * kTileBytes + 16);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seq = blockIdx.z;
  const int h = blockIdx.y / p.mtiles, mt = blockIdx.y % p.mtiles;
  const int split = blockIdx.x;
  const int start = p.seq_start[seq];
  const int P = p.seq_start[seq + 1] - start;
  const int tiles_total = (P + kTileKeys - 1) / kTileKeys;
  const int t0 = split * p.tiles_per_split;
  if (P <= 0 || t0 >= tiles_total) return;  // CTA-uniform, before any allocation
  const int t1 = min(tiles_total, t0 + p.tiles_per_split);
  const int n_obs = min(P, p.n_observed);
  const int q_offset = P - n_obs;

  // this thread's query row.  n_observed <= 128: a tile holds qt query heads x np rows; n_observed > 128: a tile
  // is one 128-query chunk of ONE query head (mtiles = Q * nchunks), the column sums of the chunks add up
  const int chunk = mt % p.nchunks, mtq = mt / p.nchunks;
  const int head_local = tid / p.np, qi = chunk * kTileRows + tid % p.np;
  const int qh_in_group = mtq * p.qt + head_local;
  const bool row_valid = (qi < n_obs) && (head_local < p.qt) && (qh_in_group < p.Q);
  const int limit = qi + q_offset;  // last key the query may attend to (causal)

  // ---- one-time setup: barriers, TMEM, Q tile ----
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const unsigned char* qbase = reinterpret_cast<const unsigned char*>(p.query);
    const int np = p.np, qt = p.qt, Q = p.Q;
    stage_tile(sQ, [&](int r) -> const unsigned char* {
      const int hl = r / np, q = chunk * kTileRows + r % np;
      const int qh = mtq * qt + hl;
      if (q >= n_obs || hl >= qt || qh >= Q) return nullptr;
      return qbase + ((int64_t)(start + q_offset + q) * p.q_stride + (int64_t)(h * Q + qh) * kHeadDim) * 2;
    });
  }
  auto stage_keys = [&](int tile, int buf) {
    const unsigned char* kbase = reinterpret_cast<const unsigned char*>(p.key);
    stage_tile(sK0 + buf * kTileBytes, [&](int r) -> const unsigned char* {
      const int kg = tile * kTileKeys + r;
      if (kg >= P) return nullptr;
      return kbase + ((int64_t)(start + kg) * p.k_stride + (int64_t)h * kHeadDim) * 2;
    });
  };
  stage_keys(t0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  // instruction descriptor: D = f32, A = B = f16 (0) / bf16 (1), both K-major, N = 128, M = 128
  constexpr uint32_t kFmt = BF16 ? 1u : 0u;
  constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(kTileKeys >> 3) << 17) |
                             ((uint32_t)(kTileRows >> 4) << 24);
  auto issue_mma = [&](int buf) {  // one thread
    const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK0 + buf * kTileBytes);
#pragma unroll
    for (int ks = 0; ks < kHeadDim / 16; ++ks) {
      const uint64_t ad = umma_desc(a0 + ks * 4096, 2048, 128);
      const uint64_t bd = umma_desc(b0 + ks * 4096, 2048, 128);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + buf * kTileKeys),
          "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&bars[buf]))
                 : "memory");
  };
  if (tid == 0) issue_mma(0);

  // ---- row state ----
  float row_max = -INFINITY, row_sum = 0.f;   // pass 1
  float M = 0.f, invL = 0.f;                  // pass 2
  int metric_limit = -1;
  const int nsplit = (tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
  const int64_t stat_base = (((int64_t)seq * p.H * p.mtiles + blockIdx.y) * p.max_splits) * kTileRows;
  if (PASS == 2) {
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, p.stats[stat_base + (int64_t)s * kTileRows + tid].x);
    float l = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float2 st = p.stats[stat_base + (int64_t)s * kTileRows + tid];
      if (st.y > 0.f) l += st.y * __expf(st.x - m);
    }
    M = m;
    invL = (l > 0.f) ? 1.f / l : 0.f;
    metric_limit = row_valid ? min(limit, limit - p.buffer_len[seq]) : -1;
  }
  const bool head_valid = (warp * 32 / p.np < p.qt) && (mtq * p.qt + warp * 32 / p.np < p.Q);

  uint32_t phases = 0u;  // bit b = parity to wait for on bars[b]
  for (int t = t0; t < t1; ++t) {
    const int buf = (t - t0) & 1;
    if (t + 1 < t1) {
      // stage the next K tile into the other buffer (its previous MMAs completed: we waited on
      // their barrier an iteration ago) and start its MMAs into the other accumulator (drained
      // by the previous epilogue) before this tile's epilogue
      stage_keys(t + 1, buf ^ 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) issue_mma(buf ^ 1);
    }
    mbar_wait(smem_u32(&bars[buf]), (phases >> buf) & 1u);
    phases ^= 1u << buf;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int key0 = t * kTileKeys;
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t r[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + buf * kTileKeys + c4 * 32, r);
      const int kbase = key0 + c4 * 32;
      if (PASS == 1) {
        float cmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float s = round_to_input_type<BF16>(__uint_as_float(r[j])) * p.scale;
          s = (kbase + j <= limit && kbase + j < P) ? s : -INFINITY;
          r[j] = __float_as_uint(s);
          cmax = fmaxf(cmax, s);
        }
        if (cmax > row_max) {
          row_sum *= __expf(row_max - cmax);  // exp(-inf) = 0 on the first live chunk
          row_max = cmax;
        }
        if (row_max > -INFINITY) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += __expf(__uint_as_float(r[j]) - row_max);
          row_sum += acc;
        }
      } else {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float s = round_to_input_type<BF16>(__uint_as_float(r[j])) * p.scale;
          float pr = __expf(s - M) * invL;
          pr = p.use_l2 ? pr * pr : pr;
          v[j] = (kbase + j <= metric_limit) ? pr : 0.f;
        }
        // reduce-scatter over the warp's 32 rows: afterwards lane j holds the sum of column j
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
          const bool upper = (lane & s) != 0;
#pragma unroll
          for (int j = 0; j < s; ++j) {
            const float send = upper ? v[j] : v[j + s];
            const float keep = upper ? v[j + s] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
          }
        }
        const int kg = kbase + lane;
        if (head_valid && kg < P) {
          float val = v[0];
          if (p.use_average) val *= (float)(kg + 1) / (float)n_obs;
          float* dst = p.raw + (int64_t)(h * p.Q + mtq * p.qt + warp * 32 / p.np) * p.T + start + kg;
          if (p.np == 32) *dst = val; else atomicAdd(dst, val);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (PASS == 1) p.stats[stat_base + (int64_t)split * kTileRows + tid] = make_float2(row_max, row_sum);

  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// max-pool(7, pad 3) over a prompt's keys + [T, Hq] output + optional aggregate_prefill.
// one thread per (token, kv head); its Q query heads in a loop.
__global__ void __launch_bounds__(256) prefill_finalize_kernel(
    float* __restrict__ kv_metric_out, float* __restrict__ kv_metrics, const int64_t* __restrict__ slot_mapping,
    const float* __restrict__ raw, const int* __restrict__ seq_start, int num_seqs, int T, int H, int Q,
    int use_maxpool) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * H) return;
  const int t = idx / H, h = idx % H;
  // the prompt holding token t (binary search over the cumulative starts)
  int lo = 0, hi = num_seqs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (seq_start[mid] <= t) lo = mid; else hi = mid;
  }
  const int a = seq_start[lo], b = seq_start[lo + 1];
  float total = 0.f;
  for (int q = 0; q < Q; ++q) {
    const float* row = raw + (int64_t)(h * Q + q) * T;
    float v = row[t];
    if (use_maxpool) {
      const int j0 = max(a, t - 3), j1 = min(b - 1, t + 3);
      for (int j = j0; j <= j1; ++j) v = fmaxf(v, row[j]);
    }
    if (kv_metric_out) kv_metric_out[(int64_t)t * H * Q + h * Q + q] = v;
    total += v;
  }
  if (kv_metrics && slot_mapping) {
    const int64_t slot = slot_mapping[idx];
    if (slot >= 0) kv_metrics[slot] += total;
  }
}

// aggregate_prefill alone (metrics.py:395-427)
__global__ void __launch_bounds__(256) aggregate_prefill_kernel(float* __restrict__ kv_metrics,
                                                                const float* __restrict__ prefill_metrics,
                                                                const int64_t* __restrict__ slot_mapping, int T,
                                                                int H, int Q) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * H) return;
  float total = 0.f;
  for (int q = 0; q < Q; ++q) total += prefill_metrics[(int64_t)idx * Q + q];
  const int64_t slot = slot_mapping[idx];
  if (slot >= 0) kv_metrics[slot] += total;
}

static inline int pad_observed(int n_observed) { return n_observed <= 32 ? 32 : (n_observed <= 64 ? 64 : 128); }
// 128-row tiles per (kv head): up to 128 observed queries several query heads share a tile, beyond that every
// query head takes ceil(n_observed / 128) tiles
static inline int prefill_mtiles(int Q, int n_observed, int* nchunks) {
  const int np = pad_observed(n_observed), qt = kTileRows / np;
  *nchunks = n_observed <= kTileRows ? 1 : (n_observed + kTileRows - 1) / kTileRows;
  return n_observed <= kTileRows ? (Q + qt - 1) / qt : Q * *nchunks;
}
constexpr int kMaxObserved = 4096;  // the reference's max_observed_block_size: one observation chunk (flash_attn.py:1130)

}  // namespace kvc

using namespace kvc;

extern "C" size_t kvc_prefill_kv_metrics_workspace_bytes(int num_tokens, int num_seqs, int num_heads,
                                                         int num_kv_heads, int n_observed, int max_prompt_len) {
  if (num_tokens <= 0 || num_seqs <= 0 || num_heads <= 0 || num_kv_heads <= 0 || n_observed <= 0 ||
      n_observed > kMaxObserved || max_prompt_len <= 0)
    return 0;
  const int Q = num_heads / num_kv_heads;
  int nchunks;
  const int mtiles = prefill_mtiles(Q, n_observed, &nchunks);
  const int tiles = (max_prompt_len + kTileKeys - 1) / kTileKeys;
  const int tps = 4;
  const int max_splits = (tiles + tps - 1) / tps;
  const size_t stats = (size_t)num_seqs * num_kv_heads * mtiles * max_splits * kTileRows * sizeof(float2);
  const size_t raw = (size_t)num_heads * num_tokens * sizeof(float);
  return stats + raw + 512;
}

extern "C" int kvc_prefill_kv_metrics(
    float* kv_metric_out, float* kv_metrics, const int64_t* slot_mapping, const void* query, const void* key,
    const int* seq_start_loc, const int* kv_metric_buffer_len, int num_tokens, int num_seqs, int num_heads,
    int num_kv_heads, int head_size, int64_t q_stride, int64_t k_stride, int max_prompt_len, float scale,
    int n_observed, int use_l2, int use_average, int use_maxpool, int dtype, void* workspace,
    size_t workspace_bytes, kvc_stream_t stream) {
  KVC_REQUIRE(head_size == kHeadDim, KVC_ERR_UNSUPPORTED, "Unsupported head size: %d", head_size);
  KVC_REQUIRE(dtype == KVC_F16 || dtype == KVC_BF16, KVC_ERR_UNSUPPORTED,
              "Unsupported data type for prefill metrics: %d", dtype);
  KVC_REQUIRE(n_observed >= 1 && n_observed <= kMaxObserved, KVC_ERR_UNSUPPORTED,
              "Unsupported prefill_observed_queries: %d (1..%d: one observation chunk of the reference)", n_observed,
              kMaxObserved);
  KVC_REQUIRE(num_kv_heads > 0 && num_heads > 0 && num_heads % num_kv_heads == 0, KVC_ERR_INVALID,
              "prefill_kv_metrics: num_heads %d not a multiple of num_kv_heads %d", num_heads, num_kv_heads);
  KVC_REQUIRE(num_tokens >= 0 && num_seqs >= 0 && max_prompt_len >= 0, KVC_ERR_INVALID,
              "prefill_kv_metrics: negative size");
  if (num_tokens == 0 || num_seqs == 0 || max_prompt_len == 0) return KVC_OK;
  KVC_REQUIRE((kv_metric_out || (kv_metrics && slot_mapping)) && query && key && seq_start_loc &&
                  kv_metric_buffer_len,
              KVC_ERR_INVALID, "prefill_kv_metrics: null tensor");
  KVC_REQUIRE((reinterpret_cast<uintptr_t>(query) & 15) == 0 && (reinterpret_cast<uintptr_t>(key) & 15) == 0 &&
                  q_stride % 8 == 0 && k_stride % 8 == 0,
              KVC_ERR_INVALID, "prefill_kv_metrics: query/key must be 16-byte aligned with strides multiple of 8");
  const size_t need = kvc_prefill_kv_metrics_workspace_bytes(num_tokens, num_seqs, num_heads, num_kv_heads,
                                                             n_observed, max_prompt_len);
  KVC_REQUIRE(workspace && workspace_bytes >= need, KVC_ERR_WORKSPACE,
              "prefill_kv_metrics: workspace of %zu bytes required, got %zu", need, workspace_bytes);
  PrefillParams p;
  p.query = query; p.key = key; p.seq_start = seq_start_loc; p.buffer_len = kv_metric_buffer_len;
  p.q_stride = q_stride; p.k_stride = k_stride;
  p.num_seqs = num_seqs; p.T = num_tokens; p.H = num_kv_heads; p.Q = num_heads / num_kv_heads;
  p.n_observed = n_observed; p.np = pad_observed(n_observed); p.qt = kTileRows / p.np;
  p.mtiles = prefill_mtiles(p.Q, n_observed, &p.nchunks);
  const int tiles = (max_prompt_len + kTileKeys - 1) / kTileKeys;
  p.tiles_per_split = 4;
  p.max_splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.scale = scale; p.use_l2 = use_l2; p.use_average = use_average;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  p.stats = reinterpret_cast<float2*>(w);
  p.raw = reinterpret_cast<float*>(w + (size_t)num_seqs * num_kv_heads * p.mtiles * p.max_splits * kTileRows *
                                           sizeof(float2));
  cudaStream_t st = as_stream(stream);
  const int smem = 3 * kTileBytes + 64;
  const dim3 grid(p.max_splits, num_kv_heads * p.mtiles, num_seqs);
  cudaError_t e;
#define KVC_PREFILL_LAUNCH(PASS, BF)                                                                       \
  do {                                                                                                     \
    e = cudaFuncSetAttribute(prefill_metrics_kernel<PASS, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             smem);                                                                        \
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "prefill_metrics smem attr: %s", cudaGetErrorString(e));  \
    prefill_metrics_kernel<PASS, BF><<<grid, 128, smem, st>>>(p);                                          \
  } while (0)
  if (p.np != 32) {  // several warps add into one query head's column sums
    e = cudaMemsetAsync(p.raw, 0, (size_t)num_heads * num_tokens * sizeof(float), st);
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "prefill_metrics memset: %s", cudaGetErrorString(e));
  }
  if (dtype == KVC_BF16) KVC_PREFILL_LAUNCH(1, true); else KVC_PREFILL_LAUNCH(1, false);
  KVC_CHECK_LAUNCH("prefill_metrics_kernel<1>");
  if (dtype == KVC_BF16) KVC_PREFILL_LAUNCH(2, true); else KVC_PREFILL_LAUNCH(2, false);
  KVC_CHECK_LAUNCH("prefill_metrics_kernel<2>");
#undef KVC_PREFILL_LAUNCH
  const int n = num_tokens * num_kv_heads;
  prefill_finalize_kernel<<<ceil_div(n, 256), 256, 0, st>>>(kv_metric_out, kv_metrics, slot_mapping, p.raw,
                                                             seq_start_loc, num_seqs, num_tokens, num_kv_heads,
                                                             p.Q, use_maxpool);
  KVC_CHECK_LAUNCH("prefill_finalize_kernel");
  return KVC_OK;
}

extern "C" int kvc_aggregate_prefill(float* kv_metrics, const float* prefill_metrics, const int64_t* slot_mapping,
                                     int num_tokens, int num_heads, int num_kv_heads, kvc_stream_t stream) {
  KVC_REQUIRE(num_kv_heads > 0 && num_heads > 0 && num_heads % num_kv_heads == 0 && num_tokens >= 0,
              KVC_ERR_INVALID, "aggregate_prefill: bad sizes");
  if (num_tokens == 0) return KVC_OK;
  KVC_REQUIRE(kv_metrics && prefill_metrics && slot_mapping, KVC_ERR_INVALID, "aggregate_prefill: null tensor");
  const int n = num_tokens * num_kv_heads;
  aggregate_prefill_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
      kv_metrics, prefill_metrics, slot_mapping, num_tokens, num_kv_heads, num_heads / num_kv_heads);
  KVC_CHECK_LAUNCH("aggregate_prefill_kernel");
  return KVC_OK;
}
