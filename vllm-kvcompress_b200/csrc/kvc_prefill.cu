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
// is issued as 8 tcgen05.mma (M=128, N=128, K=16, fp16/bf16 in, fp32 accumulators in TMEM).  Softmax needs the
// row statistics over ALL keys before a probability can be squared, so the work is two sweeps:
//   pass 1 (stats):   per (sequence, kv head, 128-row tile, key split): running (max, sum exp)
//   pass 2 (metrics): merge the splits' statistics, recompute S, p^2 = ex2(2 (s - M) log2 e) / L^2, mask,
//                     column sums per query head by a warp reduce-scatter (31 shuffles per 32 keys)
//   finalize:         max-pool 7 + [T, Hq] output (+ fused aggregate_prefill)
// A CTA is warp-specialised.  The control thread (warp 8, lane 0) brings the operands in with TMA
// (cp.async.bulk.tensor.2d through tensor maps over the caller's [T, Hq*D] query and [T, H*D] key tensors, built
// per call from their strides; a tile = two 128-byte-swizzled boxes of [rows x 64 d], the layout the K-major
// SWIZZLE_128B shared-memory descriptors expect: 8-row atoms of 1024 B, a K step of 16 elements = +32 B), keeps
// two K tiles in flight, and issues the MMAs of tile i+1 into the other TMEM accumulator while the eight epilogue
// warps work on tile i; a K buffer is requested again as soon as tcgen05.commit reports its MMAs complete.
// Epilogue warps w and w + 4 share TMEM lane quarter w & 3 (tcgen05.ld.32x32b.x32: one thread = one query row,
// so the row max / row sum of the softmax need no shuffles) and split a tile's columns.  All hand-offs are
// mbarriers (transaction barriers for the copies, tcgen05.commit for the MMAs, an arrival count of 8 for a
// drained accumulator); the loop has no CTA-wide barrier.  The key splits are sized so that the grid is one wave
// of 2 CTAs per SM; pass 2 and the finalisation are launched as programmatic dependents.
//
// Arithmetic mirrors the reference: the fp16/bf16 einsum result is rounded to the input type before
// the fp32 scale (flash_attn.py:1178: `scale * einsum(...).float()`), softmax / squares / sums fp32.
#include <cuda.h>   // CUtensorMap (types only: the encoder is looked up through the runtime, no libcuda link)

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
#ifdef KVC_PREFILL_TRACE
  unsigned long long* trace;
#endif
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#ifdef KVC_PREFILL_TRACE
// diagnostic builds only (tools/prefill_trace.py): globaltimer stamps of one epilogue thread and the control thread
__device__ __forceinline__ void pf_stamp(unsigned long long* trace, int pass, int slot) {
  if (trace == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const int cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  trace[((size_t)(pass - 1) * 4096 + cta) * 16 + slot] = t;
}
#define PF_STAMP(slot) pf_stamp(p.trace, PASS, slot)
#else
#define PF_STAMP(slot) do {} while (0)
#endif

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

// two fp32 accumulators rounded to the input type and back (the reference's einsum output is fp16 / bf16)
template <bool BF16>
__device__ __forceinline__ void round_pair(uint32_t a, uint32_t b, float& ra, float& rb) {
  if (BF16) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b));
    ra = __low2float(h);
    rb = __high2float(h);
  } else {
    const float2 f = __half22float2(__floats2half2_rn(__uint_as_float(a), __uint_as_float(b)));
    ra = f.x;
    rb = f.y;
  }
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared-memory matrix descriptor of a K-major operand in the 128-byte-swizzle layout TMA writes (rows of 64
// elements = 128 B, 8-row atoms of 1024 B back to back: SBO = 1024 B; LBO unused; layout type 2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= 1ull << 46;  // sm_100 descriptor version
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mbar_arrive_plain(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int kEpiWarps = 8;           // warps w and w + 4 share TMEM lane quarter w & 3 and split a tile's columns
constexpr int kPrefillThreads = (kEpiWarps + 1) * 32;   // + the control warp: TMA + MMA issue
constexpr int kHalfBytes = kTileKeys * 128;   // [128 keys x 64 d] swizzled half tile

// PASS 1: row statistics.  PASS 2: masked squared-probability column sums.
// Warp-specialised: the control thread (warp 4, lane 0) streams the split's K tiles with TMA (two swizzle-128B boxes
// of [128 keys x 64 d] per tile, two tiles in flight) and issues the MMAs of tile i+1 into the other TMEM
// accumulator while the eight epilogue warps run the softmax epilogue of tile i (one thread per query row and
// column half); nothing in the loop is a CTA-wide barrier.
template <int PASS, bool BF16>
__global__ void __launch_bounds__(kPrefillThreads) prefill_metrics_kernel(PrefillParams p,
                                                                          const __grid_constant__ CUtensorMap qmap,
                                                                          const __grid_constant__ CUtensorMap kmap) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the swizzled K tiles need 1024 B alignment in the shared window; the launch reserves the slack
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sQ = smem;
  unsigned char* sK0 = smem + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTileBytes);   // kbar[2] mmabar[2] accfree[2] qbar
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 3 * kTileBytes + 64);
  float2* s_stats = reinterpret_cast<float2*>(smem + 3 * kTileBytes + 128);   // [128] pass 1: second column half

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
  const int n = t1 - t0;
  const int n_obs = min(P, p.n_observed);
  const int q_offset = P - n_obs;
  if (tid == 0) PF_STAMP(0);
  // programmatic dependent launch: the next kernel's CTAs may take this SM's free resources at once
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const uint32_t kbar = smem_u32(&bars[0]), mmabar = smem_u32(&bars[2]), accfree = smem_u32(&bars[4]),
                 qbar = smem_u32(&bars[6]);
  // this row's place in the tile.  n_observed <= 128: a tile holds qt query heads x np rows; n_observed > 128: a
  // tile is one 128-query chunk of ONE query head (mtiles = Q * nchunks), the column sums of the chunks add up
  const int chunk = mt % p.nchunks, mtq = mt / p.nchunks;
  auto issue_tma = [&](int i) {   // control thread: K tile i of the split -> buffer i & 1
    const int buf = i & 1;
    const uint32_t bar = kbar + 8 * buf, dst = smem_u32(sK0 + buf * kTileBytes);
    const int row = start + (t0 + i) * kTileKeys;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTileBytes) : "memory");
#pragma unroll
    for (int half = 0; half < 2; ++half)
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          ::"r"(dst + half * kHalfBytes), "l"(&kmap), "r"(h * kHeadDim + half * 64), "r"(row), "r"(bar)
          : "memory");
  };
  if (warp == kEpiWarps && lane == 0) {
    // the control thread initialises the barriers and has the Q tile and the first two K tiles on their way before
    // the CTA has even allocated its tensor memory
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(kbar + 8 * b));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mmabar + 8 * b));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(accfree + 8 * b), "r"(kEpiWarps));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(qbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the Q tile: per query head of the tile two swizzled boxes of [np queries x 64 d]; rows of queries beyond
    // the observed window or of padding heads hold finite data of other tokens / heads and are given weight 0
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(qbar), "r"(kTileBytes) : "memory");
    for (int hl = 0; hl < p.qt; ++hl) {
      const int qh = min(mtq * p.qt + hl, p.Q - 1);
#pragma unroll
      for (int half = 0; half < 2; ++half)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(smem_u32(sQ) + half * kHalfBytes + hl * p.np * 128), "l"(&qmap),
            "r"((h * p.Q + qh) * kHeadDim + half * 64), "r"(start + q_offset + chunk * kTileRows), "r"(qbar)
            : "memory");
    }
    issue_tma(0);
    if (n > 1) issue_tma(1);
    PF_STAMP(8);
  }
  __syncwarp();
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) PF_STAMP(1);

  if (warp == kEpiWarps) {
    // =========================== control thread: TMA + MMA ===========================
    if (lane == 0) {
      constexpr uint32_t kFmt = BF16 ? 1u : 0u;
      // instruction descriptor: D = f32, A = B = f16 (0) / bf16 (1), both K-major, N = 128, M = 128
      constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(kTileKeys >> 3) << 17) |
                                 ((uint32_t)(kTileRows >> 4) << 24);
      auto issue_mma = [&](int buf) {
        const uint32_t a0 = smem_u32(sQ), b0 = smem_u32(sK0 + buf * kTileBytes);
#pragma unroll
        for (int ks = 0; ks < kHeadDim / 16; ++ks) {
          const uint64_t ad = umma_desc_sw128(a0 + (ks >> 2) * kHalfBytes + (ks & 3) * 32);
          const uint64_t bd = umma_desc_sw128(b0 + (ks >> 2) * kHalfBytes + (ks & 3) * 32);
          const uint32_t acc = ks > 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + buf * kTileKeys),
              "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         mmabar + 8 * buf)
                     : "memory");
      };
      mbar_wait(qbar, 0);   // the Q tile is staged (and visible to the tensor core)
      PF_STAMP(9);
      for (int i = 0; i < n; ++i) {
        const int buf = i & 1;
        mbar_wait(kbar + 8 * buf, (uint32_t)((i >> 1) & 1));                       // K tile i has landed
        if (i == 0) PF_STAMP(10);
        if (i >= 2) mbar_wait(accfree + 8 * buf, (uint32_t)(((i >> 1) - 1) & 1));  // epilogue i-2 drained the accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_mma(buf);
        if (i >= 1 && i + 1 < n) {   // tile i-1's MMAs are complete: its buffer takes tile i+1
          mbar_wait(mmabar + 8 * (buf ^ 1), (uint32_t)(((i - 1) >> 1) & 1));
          issue_tma(i + 1);
        }
      }
      PF_STAMP(11);
    }
  } else {
    // =========================== epilogue warps: one thread per query row ===========================
    const int quarter = warp & 3, colhalf = warp >> 2;
    const int rt = quarter * 32 + lane;   // the thread's row of the tile = its TMEM lane
    const int head_local = rt / p.np, qi = chunk * kTileRows + rt % p.np;
    const int qh_in_group = mtq * p.qt + head_local;
    const bool row_valid = (qi < n_obs) && (head_local < p.qt) && (qh_in_group < p.Q);
    const int limit = qi + q_offset;  // last key the query may attend to (causal)
    if (tid == 0) PF_STAMP(2);

    // exp(scale * r - max) = ex2(r * c + nb) with c = scale * log2(e): the scale is never applied on its own, and
    // pass 1 tracks the extremum of the ROUNDED accumulators r (the maximum of scale * r for scale >= 0, else the
    // minimum), in units of r
    constexpr float kLog2e = 1.4426950408889634f;
    const float c1 = p.scale * kLog2e;
    const bool pos = p.scale >= 0.f;
    float row_ext = pos ? -INFINITY : INFINITY, row_sum = 0.f;   // pass 1
    bool row_live = false;
    float nb = 0.f, cc = c1, wgt = 0.f;                          // pass 2: v = ex2(r * cc + nb) * wgt
    int metric_limit = -1;
    const int nsplit = (tiles_total + p.tiles_per_split - 1) / p.tiles_per_split;
    const int64_t stat_base = (((int64_t)seq * p.H * p.mtiles + blockIdx.y) * p.max_splits) * kTileRows;
    if (PASS == 2) {
      asm volatile("griddepcontrol.wait;" ::: "memory");   // pass 1's statistics (everything above is independent)
      // the splits' statistics, eight independent loads at a time (a plain loop is one L2 round trip per split)
      const float2* __restrict__ srow = p.stats + stat_base + rt;
      float m = -INFINITY;
      for (int s0 = 0; s0 < nsplit; s0 += 8) {
        float2 st[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st[j] = (s0 + j < nsplit) ? __ldcg(srow + (int64_t)(s0 + j) * kTileRows) : make_float2(-INFINITY, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m = fmaxf(m, st[j].x);
      }
      float l = 0.f;
      for (int s0 = 0; s0 < nsplit; s0 += 8) {
        float2 st[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st[j] = (s0 + j < nsplit) ? __ldcg(srow + (int64_t)(s0 + j) * kTileRows) : make_float2(-INFINITY, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (st[j].y > 0.f) l += st[j].y * __expf(st[j].x - m);
      }
      metric_limit = row_valid ? min(limit, limit - p.buffer_len[seq]) : -1;
      // p = exp(s - m) / l;  use_l2: p^2 = ex2(2 (s - m) log2 e) / l^2.  Rows that are padding (or saw no key)
      // get weight 0 and a finite exponent.
      const float invL = (row_valid && l > 0.f && m > -INFINITY) ? 1.f / l : 0.f;
      const float two = p.use_l2 ? 2.f : 1.f;
      cc = (invL > 0.f) ? two * c1 : 0.f;     // weight-0 rows: ex2(0 * r + 0) * 0, finite whatever the row holds
      nb = (invL > 0.f) ? -two * m * kLog2e : 0.f;
      wgt = p.use_l2 ? invL * invL : invL;
    }
    const bool head_valid = (quarter * 32 / p.np < p.qt) && (mtq * p.qt + quarter * 32 / p.np < p.Q);
    if (tid == 0) PF_STAMP(3);

    for (int i = 0; i < n; ++i) {
      const int buf = i & 1;
      mbar_wait(mmabar + 8 * buf, (uint32_t)((i >> 1) & 1));
      if (tid == 0 && i == 0) PF_STAMP(4);
      if (tid == 0 && i == 1) PF_STAMP(5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int key0 = (t0 + i) * kTileKeys;
#pragma unroll 1
      for (int c4 = 2 * colhalf; c4 < 2 * colhalf + 2; ++c4) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(quarter * 32) << 16) + buf * kTileKeys + c4 * 32, r);
        const int kbase = key0 + c4 * 32;
        float rr[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) round_pair<BF16>(r[j], r[j + 1], rr[j], rr[j + 1]);
        if (PASS == 1) {
          // chunks below the causal boundary and the prompt's end of every row of the warp need no masks
          if (!__all_sync(0xffffffffu, kbase + 31 <= limit && kbase + 31 < P)) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (!(kbase + j <= limit && kbase + j < P)) rr[j] = pos ? -INFINITY : INFINITY;
          }
          float cext = rr[0];
          if (pos) {
#pragma unroll
            for (int j = 1; j < 32; ++j) cext = fmaxf(cext, rr[j]);
          } else {
#pragma unroll
            for (int j = 1; j < 32; ++j) cext = fminf(cext, rr[j]);
          }
          const bool chunk_live = pos ? (cext > -INFINITY) : (cext < INFINITY);
          if (chunk_live && (!row_live || (pos ? cext > row_ext : cext < row_ext))) {
            if (row_live) row_sum *= ex2_ftz((row_ext - cext) * c1);
            row_ext = cext;
            row_live = true;
          }
          if (row_live) {
            const float nb1 = -row_ext * c1;
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              acc0 += ex2_ftz(fmaf(rr[j], c1, nb1));       // masked: ex2(-inf) = 0
              acc1 += ex2_ftz(fmaf(rr[j + 1], c1, nb1));
            }
            row_sum += acc0 + acc1;
          }
        } else {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ex2_ftz(fmaf(rr[j], cc, nb)) * wgt;
          if (!__all_sync(0xffffffffu, wgt == 0.f || kbase + 31 <= metric_limit)) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (!(kbase + j <= metric_limit)) v[j] = 0.f;
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
            float* dst = p.raw + (int64_t)(h * p.Q + mtq * p.qt + quarter * 32 / p.np) * p.T + start + kg;
            if (p.np == 32) *dst = val; else atomicAdd(dst, val);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_plain(accfree + 8 * buf);
    }
    if (tid == 0) PF_STAMP(6);
    const float row_max = row_live ? row_ext * p.scale : -INFINITY;   // the row's largest logit (natural units)
    if (PASS == 1 && colhalf == 1) s_stats[rt] = make_float2(row_max, row_sum);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PASS == 1 && colhalf == 0) {   // the row's two column halves -> one (max, sum) of the split
      const float2 o = s_stats[rt];
      const float m = fmaxf(row_max, o.x);
      float l = 0.f;
      if (row_sum > 0.f) l += row_sum * __expf(row_max - m);
      if (o.y > 0.f) l += o.y * __expf(o.x - m);
      p.stats[stat_base + (int64_t)split * kTileRows + rt] = make_float2(m, l);
    }
    if (tid == 0) PF_STAMP(7);
    return;
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// max-pool(7, pad 3) over a prompt's keys + [T, Hq] output + optional aggregate_prefill.
// one thread per (kv head, token), tokens along the warp: the column sums are read 128 B per warp and the slots of
// 16 consecutive tokens of a head are one block's 64 B.  The seven taps are independent loads with clamped indices
// (a clamped tap repeats a value of the window: the maximum is unchanged).
__global__ void __launch_bounds__(256) prefill_finalize_kernel(
    float* __restrict__ kv_metric_out, float* __restrict__ kv_metrics, const int64_t* __restrict__ slot_mapping,
    const float* __restrict__ raw, const int* __restrict__ seq_start, int num_seqs, int T, int H, int Q,
    int use_maxpool) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, h = blockIdx.y;
  if (t >= T) return;
  // independent of the previous kernel: the prompt holding token t (binary search over the cumulative starts) and
  // the token's slot
  int lo = 0, hi = num_seqs;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (seq_start[mid] <= t) lo = mid; else hi = mid;
  }
  const int a = seq_start[lo], b = seq_start[lo + 1];
  const int64_t slot = (kv_metrics && slot_mapping) ? slot_mapping[(int64_t)t * H + h] : -1;
  int tap[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) tap[j] = use_maxpool ? min(max(t + j - 3, a), b - 1) : t;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent of pass 2
  float total = 0.f;
#pragma unroll 4
  for (int q = 0; q < Q; ++q) {
    const float* __restrict__ row = raw + (int64_t)(h * Q + q) * T;
    float x[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) x[j] = __ldcg(row + tap[j]);
    float v = x[3];
#pragma unroll
    for (int j = 0; j < 7; ++j) v = fmaxf(v, x[j]);
    if (kv_metric_out) kv_metric_out[(int64_t)t * H * Q + h * Q + q] = v;
    total += v;
  }
  if (slot >= 0) kv_metrics[slot] += total;
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
constexpr int kMinTilesPerSplit = 2;
// key tiles per CTA: as few as keep the whole grid inside ONE wave of 2 CTAs per SM (the pass then costs one
// pipelined sweep of `tiles_per_split` tiles instead of a full wave plus a mostly empty one)
static inline int prefill_tiles_per_split(int groups, int tiles) {
  const int64_t slots = 2 * (int64_t)sm_count();
  int tps = kMinTilesPerSplit;
  while (tps < tiles && (int64_t)groups * ((tiles + tps - 1) / tps) > slots) ++tps;
  return tps;
}

}  // namespace kvc

using namespace kvc;

#ifdef KVC_PREFILL_TRACE
static unsigned long long* g_pf_trace = nullptr;
extern "C" __attribute__((visibility("default"))) int kvc_debug_set_prefill_trace(unsigned long long* buf) {
  g_pf_trace = buf;
  return 0;
}
#endif

extern "C" size_t kvc_prefill_kv_metrics_workspace_bytes(int num_tokens, int num_seqs, int num_heads,
                                                         int num_kv_heads, int n_observed, int max_prompt_len) {
  if (num_tokens <= 0 || num_seqs <= 0 || num_heads <= 0 || num_kv_heads <= 0 || n_observed <= 0 ||
      n_observed > kMaxObserved || max_prompt_len <= 0)
    return 0;
  const int Q = num_heads / num_kv_heads;
  int nchunks;
  const int mtiles = prefill_mtiles(Q, n_observed, &nchunks);
  const int tiles = (max_prompt_len + kTileKeys - 1) / kTileKeys;
  const int tps = kMinTilesPerSplit;   // upper bound over the launcher's choices
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
  p.tiles_per_split = prefill_tiles_per_split(num_seqs * num_kv_heads * p.mtiles, tiles);
  p.max_splits = (tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.scale = scale; p.use_l2 = use_l2; p.use_average = use_average;
#ifdef KVC_PREFILL_TRACE
  p.trace = g_pf_trace;
#endif
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  p.stats = reinterpret_cast<float2*>(w);
  p.raw = reinterpret_cast<float*>(w + (size_t)num_seqs * num_kv_heads * p.mtiles * p.max_splits * kTileRows *
                                           sizeof(float2));
  cudaStream_t st = as_stream(stream);
  // the key tensor as a 2-D map [T rows, H * D elements] (row stride = k_stride elements); a K tile is two
  // 128-byte-swizzled boxes of [128 rows x 64 elements]
  CUtensorMap qmap, kmap;
  {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
          q != cudaDriverEntryPointSuccess)
        fn = nullptr;
      return reinterpret_cast<EncodeFn>(fn);
    }();
    KVC_REQUIRE(encode != nullptr, KVC_ERR_CUDA, "prefill_kv_metrics: cuTensorMapEncodeTiled not available");
    {
      const cuuint64_t qdim[2] = {(cuuint64_t)num_heads * kHeadDim, (cuuint64_t)num_tokens};
      const cuuint64_t qstride[1] = {(cuuint64_t)q_stride * 2};
      const cuuint32_t qbox[2] = {64, (cuuint32_t)pad_observed(n_observed)};
      const cuuint32_t qes[2] = {1, 1};
      const CUresult rq = encode(&qmap, dtype == KVC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                                 2, const_cast<void*>(query), qdim, qstride, qbox, qes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      KVC_REQUIRE(rq == CUDA_SUCCESS, KVC_ERR_CUDA, "prefill_kv_metrics: query tensor map encode failed (%d)", (int)rq);
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)num_kv_heads * kHeadDim, (cuuint64_t)num_tokens};
    const cuuint64_t gstride[1] = {(cuuint64_t)k_stride * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kTileKeys};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(&kmap, dtype == KVC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                              2, const_cast<void*>(key), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KVC_REQUIRE(r == CUDA_SUCCESS, KVC_ERR_CUDA, "prefill_kv_metrics: tensor map encode failed (%d)", (int)r);
  }
  const int smem = 3 * kTileBytes + 128 + 1024 + 1024;
  const dim3 grid(p.max_splits, num_kv_heads * p.mtiles, num_seqs);
  cudaError_t e;
  // pass 2 and the finalisation are programmatic dependents: their CTAs take over as the previous kernel's CTAs
  // leave, stage their operands, and wait (griddepcontrol.wait) only in front of the previous kernel's results
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
#define KVC_PREFILL_LAUNCH(PASS, BF, DEPENDENT)                                                            \
  do {                                                                                                     \
    e = cudaFuncSetAttribute(prefill_metrics_kernel<PASS, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             smem);                                                                        \
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "prefill_metrics smem attr: %s", cudaGetErrorString(e));  \
    cudaLaunchConfig_t cfg = {};                                                                           \
    cfg.gridDim = grid; cfg.blockDim = dim3(kPrefillThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;            \
    cfg.attrs = pdl_attr; cfg.numAttrs = (DEPENDENT) ? 1 : 0;                                              \
    e = cudaLaunchKernelEx(&cfg, prefill_metrics_kernel<PASS, BF>, p, qmap, kmap);                                     \
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "prefill_metrics launch: %s", cudaGetErrorString(e));     \
  } while (0)
  if (p.np != 32) {  // several warps add into one query head's column sums
    e = cudaMemsetAsync(p.raw, 0, (size_t)num_heads * num_tokens * sizeof(float), st);
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "prefill_metrics memset: %s", cudaGetErrorString(e));
  }
  if (dtype == KVC_BF16) KVC_PREFILL_LAUNCH(1, true, false); else KVC_PREFILL_LAUNCH(1, false, false);
  KVC_CHECK_LAUNCH("prefill_metrics_kernel<1>");
  if (dtype == KVC_BF16) KVC_PREFILL_LAUNCH(2, true, true); else KVC_PREFILL_LAUNCH(2, false, true);
  KVC_CHECK_LAUNCH("prefill_metrics_kernel<2>");
#undef KVC_PREFILL_LAUNCH
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ceil_div(num_tokens, 256), num_kv_heads); cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cfg.attrs = pdl_attr; cfg.numAttrs = 1;
    const float* raw_c = p.raw;
    const int Qn = p.Q;
    e = cudaLaunchKernelEx(&cfg, prefill_finalize_kernel, kv_metric_out, kv_metrics, slot_mapping, raw_c,
                           seq_start_loc, num_seqs, num_tokens, num_kv_heads, Qn, use_maxpool);
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "prefill_finalize launch: %s", cudaGetErrorString(e));
  }
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
