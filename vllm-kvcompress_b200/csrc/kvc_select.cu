// Fused eviction selection: per-(layer, head) metric ranking + per-sequence top-k of
// block-completion keys + per-head ascending eviction lists, straight from the paged
// metrics through the block tables.
//
// replaces CompressionMetrics.schedule_evictions (vllm/kvcompress/metrics.py:441-847): the
// boolean-mask gathers over the whole cache, 4 full-length torch.sort + 2 block-length sorts,
// ~20 gathers, a per-sequence Python loop with host syncs (:709-729), count_block_evictions
// (:774) and the final (head, logical index) sort (:822-834) -- and equally the V1 chain
// sort_seq_metrics (:849-966) -> schedule_cache_evictions+truncate (K1/K2,
// csrc/kvcompress_eviction_kernels.cu:27-188) -> host re-sort (scheduler.py:472-477).
// Output contract is the one schedule_t1_cache_moves consumes: every head segment
// [off, off + ceil(ctx/B)*B) of evicted_logical_indices holds the head's evicted logical
// indices ascending, NULL after them; evicted_kv_count = (k-1)*B + hanging for k freed
// blocks (metrics.py:788-792); evicted_block_count = k.
//
// Semantics (identical to K1+K2 on distinct metrics; ties broken by (metric, logical index),
// cross-head ties by (head, block) order): a slot is a candidate iff logical index < ctx and
// kv_position <= last_position - protected_window (K1 :136-139).  Ranking a head's candidates
// by metric, the head frees a block at ranks r0, r0+B, ... (r0 = hanging tokens); the
// "completion key" of that block is the metric at that rank.  Per sequence the E smallest
// completion keys over all heads are taken (E = evicted_blocks_per_seq); a head with k taken
// keys evicts its (k-1)*B + r0 lowest-metric candidates.
//
// Three launches, no host round trip:
//   1. keys:   ctx <= 512: one warp per head, 8 or 16 consecutive slots per lane (two 16 B loads of
//              metrics and of positions), a 256/512-key bitonic sort in registers (warp shuffles), the
//              completion keys = sorted ranks r0-1+kB go to the workspace.  Longer contexts: one CTA
//              per head, keys staged in shared memory, rank by counting.
//   2. select: one CTA per sequence; 4-pass radix select of the E-th smallest completion key,
//              then per-head lower/upper bounds + an exclusive scan for deterministic ties.
//   3. emit:   same grouping as 1; a slot is evicted iff its key is below the head's last granted
//              completion key (ties at that key in logical order); ordered compaction by prefix sums.
// All of it is int32/fp32 gather traffic: 2 x 8 B read per slot, 4 B per evicted index, 4 B per
// block of completion keys; HBM/L2-bound, no tensor-core shapes.
#include <stdlib.h>

#include <cooperative_groups.h>

#include "kvc_common.cuh"

namespace kvc {

constexpr unsigned kInfKey = 0xffffffffu;

__device__ __forceinline__ unsigned orderable(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone float -> uint
}

struct SelectParams {
  int* evicted_logical_indices;
  int* evicted_kv_count;     // [S,L,H]
  int* evicted_block_count;  // [S,L,H]
  const float* kv_metrics;   // [NB,B]
  const int* kv_position;    // [NB,B]
  const int* block_tables;   // [L,S,H,MB]
  const int* context_lens;   // [L,S,H]
  const int* hanging;        // [S,L,H]
  const int* offsets;        // [S,L,H]
  const int* budgets;        // [S]
  const int* last_position;  // [S]
  const int* protected_window;  // [S]
  int S, L, H, MB, B, null_index;
  unsigned* events;  // [total slots / B] workspace
  // optional terms of the live selection (metrics.py:495-506, 539-544); all off by default
  const float* head_bias;     // [L, H, num_bins] or null (KVHeadBias.bias, metrics.py:45-81)
  const int* position_bins;   // [num_bins] ascending bin starts
  int num_bins;
  float bias_weight;          // kv_metric_bias_weight
  int use_average;            // metric / (last_position - token_position)
  int num_sinks;              // positions < num_sinks are never evicted
};

// eviction key of one slot: kInfKey = not a candidate.  The float arithmetic is the reference's
// elementwise torch ops, one rounding each (no FMA contraction): metric / count, + bias * weight.
__device__ __forceinline__ unsigned slot_key(const SelectParams& p, float m, int pos, int max_pos, int last_pos,
                                             int l, int h) {
  if (pos > max_pos || pos < p.num_sinks) return kInfKey;
  if (p.use_average) m = __fdiv_rn(m, (float)(last_pos - pos));
  if (p.head_bias != nullptr) {
    float b = 0.f;
    if (pos >= 0) {
      int bin = -1;  // (#bins with start <= pos) - 1; -1 wraps to the last bin like the reference's index
      for (int i = 0; i < p.num_bins; ++i) bin += (pos >= p.position_bins[i]);
      if (bin < 0) bin = p.num_bins - 1;
      b = p.head_bias[((int64_t)l * p.H + h) * p.num_bins + bin];
    }
    m = __fadd_rn(m, __fmul_rn(b, p.bias_weight));
  }
  return orderable(m);
}

// default selection (no head bias, no averaging, no sinks): the key is the metric itself
__device__ __forceinline__ unsigned slot_key_plain(float m, int pos, int max_pos) {
  return (pos > max_pos || pos < 0) ? kInfKey : orderable(m);
}

// load the (up to) 8 consecutive logical slots i0..i0+7 of a head as orderable keys
// (kInfKey for non-candidates and for i >= n); i0 is a multiple of 8 so the slots are half a block
template <int B>
__device__ __forceinline__ void load_keys8(const SelectParams& p, const int* __restrict__ bt, int i0, int n,
                                           int max_pos, int last_pos, int l, int h, unsigned (&k)[8]) {
  if (B == 16) {
    if (i0 < n) {
      const int64_t slot = (int64_t)bt[i0 >> 4] * 16 + (i0 & 15);
      const float4 m0 = *reinterpret_cast<const float4*>(p.kv_metrics + slot);
      const float4 m1 = *reinterpret_cast<const float4*>(p.kv_metrics + slot + 4);
      const int4 q0 = *reinterpret_cast<const int4*>(p.kv_position + slot);
      const int4 q1 = *reinterpret_cast<const int4*>(p.kv_position + slot + 4);
      const float m[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      const int q[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r) k[r] = (i0 + r < n) ? slot_key(p, m[r], q[r], max_pos, last_pos, l, h) : kInfKey;
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) k[r] = kInfKey;
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = i0 + r;
      unsigned key = kInfKey;
      if (i < n) {
        const int64_t slot = (int64_t)bt[i / B] * B + i % B;
        key = slot_key(p, p.kv_metrics[slot], p.kv_position[slot], max_pos, last_pos, l, h);
      }
      k[r] = key;
    }
  }
}

// ---- warp-per-head path (ctx <= 512): bitonic sort of the head's CANDIDATE keys in registers ----
// ascending bitonic sort of 32*KPL keys, element e = lane * KPL + r
template <int KPL>
__device__ __forceinline__ void warp_bitonic_sort(unsigned (&k)[KPL], int lane) {
  constexpr int N = 32 * KPL;
#pragma unroll
  for (int kk = 2; kk <= N; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j >= 1; j >>= 1) {
      if (j >= KPL) {
        const int lj = j / KPL;
        const bool up = ((lane * KPL) & kk) == 0;    // ascending segment (kk >= 2*KPL here)
        const bool lower = (lane & lj) == 0;         // this lane holds the lower index of the pair
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
          const unsigned o = __shfl_xor_sync(0xffffffffu, k[r], lj);
          const unsigned mn = min(k[r], o), mx = max(k[r], o);
          k[r] = (up == lower) ? mn : mx;
        }
      } else {
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
          const int pr = r ^ j;
          if (pr > r) {
            const bool up = (((lane * KPL + r) & kk) == 0);
            const unsigned a = k[r], b = k[pr];
            const unsigned mn = min(a, b), mx = max(a, b);
            k[r] = up ? mn : mx;
            k[pr] = up ? mx : mn;
          }
        }
      }
    }
  }
}

// completion keys = sorted ranks r0-1 + kB (k < nev) of a warp-sorted key set
template <int B, int KPL>
__device__ __forceinline__ void write_completion_keys(const unsigned (&k)[KPL], int lane, int r0, int nev,
                                                      unsigned* ev) {
#pragma unroll
  for (int r = 0; r < KPL; ++r) {
    const int e = lane * KPL + r;
    const int d = e - (r0 - 1);
    if (d >= 0 && d % B == 0 && d / B < nev) ev[d / B] = k[r];
  }
}

// Only candidates can carry a completion key, and the protected window / the slack of the last block
// usually leave fewer than half of the loaded slots as candidates: the candidate keys are compacted
// through the warp's shared-memory row first, and the sort runs on the smallest power of two that
// holds them (a 128-key network costs ~45 % of the 256-key one).
template <int B, int KPL>
__global__ void __launch_bounds__(128) select_sort_kernel(SelectParams p) {
  __shared__ unsigned s_keys[4][32 * KPL];
  const int lane = threadIdx.x % kWarp, w = threadIdx.x / kWarp;
  const int g = blockIdx.x * 4 + w;
  if (g >= p.S * p.L * p.H) return;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int nblk = (n + B - 1) / B;
  const int off = p.offsets[g];
  const int r0 = p.hanging[g];
  const int last_pos = p.last_position[s];
  const int max_pos = last_pos - p.protected_window[s];
  unsigned k[KPL];
#pragma unroll
  for (int q = 0; q < KPL / 8; ++q) {
    unsigned t8[8];
    load_keys8<B>(p, p.block_tables + (int64_t)lsh * p.MB, lane * KPL + 8 * q, n, max_pos, last_pos, l, h, t8);
#pragma unroll
    for (int r = 0; r < 8; ++r) k[8 * q + r] = t8[r];
  }
  int local = 0;
#pragma unroll
  for (int r = 0; r < KPL; ++r) local += (k[r] != kInfKey);
  int incl = local;
#pragma unroll
  for (int d = 1; d < kWarp; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int ncand = __shfl_sync(0xffffffffu, incl, 31);
  const int nev = (r0 >= 1 && ncand >= r0) ? (ncand - r0) / B + 1 : 0;
  unsigned* ev = p.events + off / B;
  if (nev > 0) {  // warp-uniform
    if (ncand <= 32 * (KPL / 2)) {
      // compact the candidates, pad with inf, sort half (or a quarter) as many keys
      unsigned* row = s_keys[w];
      int pos = incl - local;
#pragma unroll
      for (int r = 0; r < KPL; ++r)
        if (k[r] != kInfKey) row[pos++] = k[r];
      for (int i = ncand + lane; i < 32 * (KPL / 2); i += kWarp) row[i] = kInfKey;
      __syncwarp();
      if (KPL >= 16 && ncand <= 32 * (KPL / 4)) {
        unsigned kq[KPL / 4 > 0 ? KPL / 4 : 1];
#pragma unroll
        for (int r = 0; r < KPL / 4; ++r) kq[r] = row[lane * (KPL / 4) + r];
        warp_bitonic_sort<(KPL / 4 > 0 ? KPL / 4 : 1)>(kq, lane);
        write_completion_keys<B, (KPL / 4 > 0 ? KPL / 4 : 1)>(kq, lane, r0, nev, ev);
      } else {
        unsigned kh[KPL / 2];
#pragma unroll
        for (int r = 0; r < KPL / 2; ++r) kh[r] = row[lane * (KPL / 2) + r];
        warp_bitonic_sort<KPL / 2>(kh, lane);
        write_completion_keys<B, KPL / 2>(kh, lane, r0, nev, ev);
      }
    } else {
      warp_bitonic_sort<KPL>(k, lane);
      write_completion_keys<B, KPL>(k, lane, r0, nev, ev);
    }
  }
  for (int q = nev + lane; q < nblk; q += kWarp) ev[q] = kInfKey;
}

// evicted logical indices of one head, ascending, NULL after them (one warp)
template <int B, int KPL>
__device__ __forceinline__ void emit_head(const SelectParams& p, int s, int l, int h, int lsh, int n, int off,
                                          int cnt, int kblk, int lane) {
  const int nblk = (n + B - 1) / B;
  int* __restrict__ out = p.evicted_logical_indices + off;
  if (cnt > 0) {
    const int last_pos = p.last_position[s];
    const int max_pos = last_pos - p.protected_window[s];
    unsigned k[KPL];
#pragma unroll
    for (int q = 0; q < KPL / 8; ++q) {
      unsigned t8[8];
      load_keys8<B>(p, p.block_tables + (int64_t)lsh * p.MB, lane * KPL + 8 * q, n, max_pos, last_pos, l, h, t8);
#pragma unroll
      for (int r = 0; r < 8; ++r) k[8 * q + r] = t8[r];
    }
    const unsigned T = p.events[off / B + kblk - 1];  // key of the last evicted rank
    int nless = 0, neq = 0;
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
      nless += (k[r] < T);
      neq += (k[r] == T);
    }
    int tot_less = nless;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) tot_less += __shfl_xor_sync(0xffffffffu, tot_less, d);
    int eq_incl = neq;  // ties at T are granted in logical order
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, eq_incl, d);
      if (lane >= d) eq_incl += v;
    }
    int eq_seen = eq_incl - neq;
    const int allowed = cnt - tot_less;
    bool evict[KPL];
    int nev = 0;
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
      evict[r] = (k[r] < T) || (k[r] == T && eq_seen < allowed);
      eq_seen += (k[r] == T);
      nev += evict[r];
    }
    int incl = nev;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int pos = incl - nev;
#pragma unroll
    for (int r = 0; r < KPL; ++r)
      if (evict[r]) out[pos++] = lane * KPL + r;
  }
  for (int i = cnt + lane; i < nblk * B; i += kWarp) out[i] = p.null_index;
}

template <int B, int KPL>
__global__ void __launch_bounds__(128) select_emit_warp_kernel(SelectParams p) {
  const int lane = threadIdx.x % kWarp;
  const int g = blockIdx.x * 4 + threadIdx.x / kWarp;
  if (g >= p.S * p.L * p.H) return;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int lsh = (l * p.S + s) * p.H + h;
  emit_head<B, KPL>(p, s, l, h, lsh, p.context_lens[lsh], p.offsets[g], p.evicted_kv_count[g],
                    p.evicted_block_count[g], lane);
}

// ---- CTA-per-head path (long contexts): rank by counting over keys staged in shared memory ----
__global__ void __launch_bounds__(128) select_rank_kernel(SelectParams p, int max_slots) {
  extern __shared__ unsigned keys[];  // [max_slots]
  __shared__ int s_ncand;
  const int t = threadIdx.x;
  const int g = blockIdx.x;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int nblk = (n + p.B - 1) / p.B;
  const int off = p.offsets[g];
  const int r0 = p.hanging[g];
  const int last_pos = p.last_position[s];
  const int max_pos = last_pos - p.protected_window[s];
  const int* __restrict__ bt = p.block_tables + (int64_t)lsh * p.MB;
  if (t == 0) s_ncand = 0;
  __syncthreads();
  int local = 0;
  for (int i = t; i < n; i += 128) {
    const int64_t slot = (int64_t)bt[i / p.B] * p.B + i % p.B;
    const unsigned k = slot_key(p, p.kv_metrics[slot], p.kv_position[slot], max_pos, last_pos, l, h);
    local += (k != kInfKey);
    keys[i] = k;
  }
  if (local) atomicAdd(&s_ncand, local);
  __syncthreads();
  const int ncand = s_ncand;
  const int nev = (r0 >= 1 && ncand >= r0) ? (ncand - r0) / p.B + 1 : 0;
  unsigned* ev = p.events + off / p.B;
  for (int i = t; i < n; i += 128) {
    const unsigned ki = keys[i];
    if (ki == kInfKey) continue;
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned kj = keys[j];
      r += (kj < ki) || (kj == ki && j < i);
    }
    const int d = r - (r0 - 1);
    if (r0 >= 1 && d >= 0 && d % p.B == 0) ev[d / p.B] = ki;
  }
  for (int k = nev + t; k < nblk; k += 128) ev[k] = kInfKey;
}

// ---- CTA-per-head path, contexts up to 32768 slots: bitonic sort of the keys in shared memory ----
// Only key VALUES at ranks r0-1+kB are needed (equal keys are interchangeable), so the sort carries
// no payload.  Each head sorts the next power of two >= its own context length.
constexpr int kSortThreads = 512;
__global__ void __launch_bounds__(kSortThreads) select_sort_cta_kernel(SelectParams p) {
  extern __shared__ unsigned keys[];
  __shared__ int s_ncand;
  const int t = threadIdx.x;
  const int g = blockIdx.x;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int nblk = (n + p.B - 1) / p.B;
  const int off = p.offsets[g];
  const int r0 = p.hanging[g];
  const int last_pos = p.last_position[s];
  const int max_pos = last_pos - p.protected_window[s];
  const int* __restrict__ bt = p.block_tables + (int64_t)lsh * p.MB;
  int m = 64;
  while (m < n) m <<= 1;
  if (t == 0) s_ncand = 0;
  __syncthreads();
  int local = 0;
  for (int i = t; i < m; i += kSortThreads) {
    unsigned k = kInfKey;
    if (i < n) {
      const int64_t slot = (int64_t)bt[i / p.B] * p.B + i % p.B;
      k = slot_key(p, p.kv_metrics[slot], p.kv_position[slot], max_pos, last_pos, l, h);
    }
    local += (k != kInfKey);
    keys[i] = k;
  }
  if (local) atomicAdd(&s_ncand, local);
  __syncthreads();
  for (int kk = 2; kk <= m; kk <<= 1) {
    for (int j = kk >> 1; j >= 1; j >>= 1) {
      for (int idx = t; idx < (m >> 1); idx += kSortThreads) {
        const int i = ((idx & ~(j - 1)) << 1) | (idx & (j - 1));  // idx with a 0 inserted at bit log2(j)
        const unsigned a = keys[i], b = keys[i | j];
        const bool up = (i & kk) == 0;
        if ((a > b) == up) {
          keys[i] = b;
          keys[i | j] = a;
        }
      }
      __syncthreads();
    }
  }
  const int ncand = s_ncand;
  const int nev = (r0 >= 1 && ncand >= r0) ? (ncand - r0) / p.B + 1 : 0;
  unsigned* ev = p.events + off / p.B;
  for (int k = t; k < nblk; k += kSortThreads) ev[k] = (k < nev) ? keys[r0 - 1 + k * p.B] : kInfKey;
}

__global__ void __launch_bounds__(128) select_emit_cta_kernel(SelectParams p) {
  __shared__ int s_less, s_w[4], s_w2[4];
  const int t = threadIdx.x, lane = t % kWarp, w = t / kWarp;
  const int g = blockIdx.x;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int nblk = (n + p.B - 1) / p.B;
  const int off = p.offsets[g];
  const int cnt = p.evicted_kv_count[g];
  const int kblk = p.evicted_block_count[g];
  const int last_pos = p.last_position[s];
  const int max_pos = last_pos - p.protected_window[s];
  const int* __restrict__ bt = p.block_tables + (int64_t)lsh * p.MB;
  int* __restrict__ out = p.evicted_logical_indices + off;
  if (cnt > 0) {  // CTA-uniform
    const unsigned T = p.events[off / p.B + kblk - 1];
    auto key_of = [&](int i) -> unsigned {
      const int64_t slot = (int64_t)bt[i / p.B] * p.B + i % p.B;
      return slot_key(p, p.kv_metrics[slot], p.kv_position[slot], max_pos, last_pos, l, h);
    };
    if (t == 0) s_less = 0;
    __syncthreads();
    int local = 0;
    for (int i = t; i < n; i += 128) local += (key_of(i) < T);
    if (local) atomicAdd(&s_less, local);
    __syncthreads();
    const int allowed = cnt - s_less;
    int eq_base = 0, written = 0;
    for (int i0 = 0; i0 < n; i0 += 128) {
      const int i = i0 + t;
      const unsigned k = (i < n) ? key_of(i) : kInfKey;
      const bool less = k < T, eq = (k == T) && (k != kInfKey || T != kInfKey) && i < n;
      const unsigned em = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_w[w] = __popc(em);
      __syncthreads();
      int eq_before = eq_base + __popc(em & lanemask_lt());
      int eq_tot = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < w) eq_before += s_w[q];
        eq_tot += s_w[q];
      }
      const bool ev = (i < n) && (less || (eq && eq_before < allowed));
      const unsigned vm = __ballot_sync(0xffffffffu, ev);
      if (lane == 0) s_w2[w] = __popc(vm);
      __syncthreads();
      int pos = written + __popc(vm & lanemask_lt());
      int tot = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q < w) pos += s_w2[q];
        tot += s_w2[q];
      }
      if (ev) out[pos] = i;
      written += tot;
      eq_base += eq_tot;
      __syncthreads();
    }
  }
  for (int i = cnt + t; i < nblk * p.B; i += 128) out[i] = p.null_index;
}

// =====================================================================================================
// Long contexts (post-prefill compression: a few sequences, thousands of slots per head) WITHOUT a sort.
// Only the number of blocks a head completes matters: with R = #candidate keys of the head <= x, the head has
// f(R) = (R >= r0 ? (R - r0) / B + 1 : 0) completion keys <= x (its k-th block completes at sorted rank
// r0 - 1 + k B).  The per-sequence threshold K = the E-th smallest completion key is therefore the smallest x with
// F(x) = sum over the sequence's heads of f(R_head(x)) >= E, found by a 4 x 8-bit radix descent in which every
// head only HISTOGRAMS its keys: O(n) reads per pass instead of an O(n log^2 n) bitonic sort of 8192 keys per head.
//   radix_pass<0>   CTA per head: keys from metrics / positions through the block table (the one DRAM pass),
//                   written to the workspace (4 B per slot, L2 resident: 18 MB at 3 x 6000-token prompts);
//                   digit histogram -> inclusive scan -> f(base + cum[d]) added into Fsum[pass][seq][d]
//   radix_pass<1-3> the same on the stored keys; the digit chosen by the previous pass is re-derived by every CTA
//                   from Fsum (1 KB) -- no host step, no extra launch
//   radix_count     per head: #keys < K and #keys == K  -> grants (ties at K in head order, as the sort path)
//   radix_grant     CTA per sequence: exclusive scan of the tie counts over the heads
//   radix_emit      CTA per head: the head's cnt-th smallest key by a radix select over its keys in shared memory,
//                   then the ordered emission of select_emit_cta_kernel
// Same keys, same tie rules (completion keys by (key, head, block), slots by (key, logical index)): results are
// bit-identical to the sort path (tests run both).
// =====================================================================================================
struct RadixWs {
  unsigned* keys;   // [N] orderable key per slot of every head segment (offsets[g] + i)
  int* fsum;        // [4][S][256]
};

__device__ __forceinline__ int completions(int R, int r0, int B) { return (r0 >= 1 && R >= r0) ? (R - r0) / B + 1 : 0; }

// smallest digit d with F[d] >= E in fsum row (256 ints), or 256 if none; every thread of a 256-thread CTA calls it
__device__ __forceinline__ int pick_digit(const int* __restrict__ row, int E, int* s_tmp) {
  const int t = threadIdx.x;
  if (t == 0) *s_tmp = 256;
  __syncthreads();
  if (row[t] >= E) atomicMin(s_tmp, t);
  __syncthreads();
  const int d = *s_tmp;
  __syncthreads();
  return d;
}

// state after `PASS` digits: prefix (high bits decided so far) or "all" when even the largest key does not reach E
__device__ __forceinline__ bool radix_state(const RadixWs& w, int S, int s, int E, int npass, unsigned& prefix,
                                            int* s_tmp) {
  prefix = 0;
  for (int q = 0; q < npass; ++q) {
    const int d = pick_digit(w.fsum + ((size_t)q * S + s) * 256, E, s_tmp);
    if (d >= 256) return false;  // fewer than E finite completion keys (can only happen at q == 0)
    prefix |= (unsigned)d << (24 - 8 * q);
  }
  return true;
}

__device__ __forceinline__ void hist_add(int* hist, unsigned digit) { atomicAdd(&hist[digit], 1); }

// shared scratch of the radix kernels (one instance per CTA; the fused kernel reuses it across its phases)
struct RadixSmem {
  int hist[256];
  int scan[256];
  int w[8], w2[8];
  int tmp, base, less, eq, carry, lesssum, need;
  unsigned prefix;
  int hist0[256];   // cooperative kernel: the head's top-byte histogram of its finite keys (pass 0), for the emission
  int finite;       // ... and their number
};

// 256-thread CTA: the bin b with cum[b-1] < need <= cum[b] (inclusive prefix sums of hist) and need - cum[b-1];
// parallel replacement of a serial walk over the bins.  Returns through shared memory.
__device__ __forceinline__ void pick_bin(const int* hist, int need, int* s_w, unsigned* s_bin, int* s_need) {
  const int t = threadIdx.x, lane = t % kWarp, wid = t / kWarp;
  int v = hist[t];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += u;
  }
  if (lane == 31) s_w[wid] = v;
  if (t == 0) { *s_bin = 255u; *s_need = 1; }
  __syncthreads();
  for (int q = 0; q < wid; ++q) v += s_w[q];
  const int before = v - hist[t];
  if (before < need && need <= v) { *s_bin = (unsigned)t; *s_need = need - before; }
  __syncthreads();
}

// One digit pass of head g.  kbuf: the head's keys in shared memory (fused kernel) or nullptr (keys live in the
// workspace).  Every thread of the CTA returns together (no barrier is skipped by a subset).
template <int PASS>
__device__ __forceinline__ void radix_pass_body(const SelectParams& p, const RadixWs& w, RadixSmem& sm, int g,
                                                unsigned* kbuf, int* s_bt = nullptr) {
  int* hist = sm.hist;
  int* s_w = sm.w;
  const int t = threadIdx.x, lane = t % kWarp, wid = t / kWarp;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int E = p.budgets[s];
  if (E <= 0) return;  // CTA-uniform
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int off = p.offsets[g];
  const int r0 = p.hanging[g];
  unsigned prefix = 0;
  if (PASS > 0 && !radix_state(w, p.S, s, E, PASS, prefix, &sm.tmp)) return;
  constexpr int shift = 24 - 8 * PASS;
  const unsigned hi_mask = PASS == 0 ? 0u : ~((1u << (shift + 8)) - 1u);
  hist[t] = 0;
  if (t == 0) sm.base = 0;
  __syncthreads();
  unsigned* keys = kbuf ? kbuf : w.keys + off;
  int base = 0;  // keys below the prefix range
  if (PASS == 0) {
    const int last_pos = p.last_position[s];
    const int max_pos = last_pos - p.protected_window[s];
    const int* __restrict__ bt = p.block_tables + (int64_t)lsh * p.MB;
    // the head's block-table row goes to shared memory first (one coalesced read), so the gather below is ONE
    // round of independent loads per slot instead of table -> data; eight slots per thread in flight
    const int nblk = (n + p.B - 1) / p.B;
    for (int b = t; b < nblk; b += 256) s_bt[b] = bt[b];
    __syncthreads();
    for (int i0 = t; i0 < n; i0 += 8 * 256) {
      float mm[8];
      int qq[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 256;
        const int64_t slot = (i < n) ? (int64_t)s_bt[i / p.B] * p.B + i % p.B : 0;
        mm[u] = p.kv_metrics[slot];
        qq[u] = p.kv_position[slot];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 256;
        if (i < n) {
          const unsigned k = slot_key(p, mm[u], qq[u], max_pos, last_pos, l, h);
          keys[i] = k;
          if (k != kInfKey) hist_add(hist, k >> 24);
        }
      }
    }
  } else {
    auto one = [&](unsigned k) {
      if (k == kInfKey) return;
      if ((k & hi_mask) == prefix) hist_add(hist, (k >> shift) & 255u);
      else if (k < prefix) ++base;
    };
    if (kbuf || (off & 3) == 0) {  // 16-byte key loads (segments start at multiples of the block size)
      const uint4* k4 = reinterpret_cast<const uint4*>(keys);
      const int n4 = n >> 2;
      for (int i = t; i < n4; i += 256) {
        const uint4 v = k4[i];
        one(v.x); one(v.y); one(v.z); one(v.w);
      }
      for (int i = (n4 << 2) + t; i < n; i += 256) one(keys[i]);
    } else {
      for (int i = t; i < n; i += 256) one(keys[i]);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) base += __shfl_xor_sync(0xffffffffu, base, d);
    if (lane == 0 && base) atomicAdd(&sm.base, base);
  }
  __syncthreads();
  // inclusive scan of the 256 bins
  int v = hist[t];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += u;
  }
  if (lane == 31) s_w[wid] = v;
  __syncthreads();
  for (int q = 0; q < wid; ++q) v += s_w[q];
  const int c = completions(sm.base + v, r0, p.B);
  if (c) atomicAdd(w.fsum + ((size_t)PASS * p.S + s) * 256 + t, c);
  __syncthreads();   // the scratch is reused by the next phase of the fused kernel
}

template <int PASS>
__global__ void __launch_bounds__(256) select_radix_pass_kernel(SelectParams p, RadixWs w) {
  extern __shared__ __align__(16) int s_bt_dyn[];   // pass 0: the head's block-table row (MB ints)
  __shared__ RadixSmem sm;
  radix_pass_body<PASS>(p, w, sm, blockIdx.x, nullptr, s_bt_dyn);
}

// per head: blocks completed by keys < K (-> evicted_block_count) and the extra ones completed by keys == K
// (-> evicted_kv_count), K = inf: every finite key
__device__ __forceinline__ void radix_count_body(const SelectParams& p, const RadixWs& w, RadixSmem& sm, int g,
                                                 const unsigned* kbuf) {
  const int t = threadIdx.x, lane = t % kWarp;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int E = p.budgets[s];
  if (E <= 0) {
    if (t == 0) { p.evicted_block_count[g] = 0; p.evicted_kv_count[g] = 0; }
    return;
  }
  unsigned K = 0;
  const bool found = radix_state(w, p.S, s, E, 4, K, &sm.tmp);
  const int n = p.context_lens[(l * p.S + s) * p.H + h];
  const unsigned* keys = kbuf ? kbuf : w.keys + p.offsets[g];
  if (t == 0) sm.less = sm.eq = 0;
  __syncthreads();
  int less = 0, eq = 0;
  auto one = [&](unsigned k) {
    if (k == kInfKey) return;
    if (!found || k < K) ++less;
    else if (k == K) ++eq;
  };
  if (kbuf || (p.offsets[g] & 3) == 0) {
    const uint4* k4 = reinterpret_cast<const uint4*>(keys);
    const int n4 = n >> 2;
    for (int i = t; i < n4; i += 256) {
      const uint4 v = k4[i];
      one(v.x); one(v.y); one(v.z); one(v.w);
    }
    for (int i = (n4 << 2) + t; i < n; i += 256) one(keys[i]);
  } else {
    for (int i = t; i < n; i += 256) one(keys[i]);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    less += __shfl_xor_sync(0xffffffffu, less, d);
    eq += __shfl_xor_sync(0xffffffffu, eq, d);
  }
  if (lane == 0) {
    if (less) atomicAdd(&sm.less, less);
    if (eq) atomicAdd(&sm.eq, eq);
  }
  __syncthreads();
  if (t == 0) {
    const int r0 = p.hanging[g];
    const int c_less = completions(sm.less, r0, p.B);
    p.evicted_block_count[g] = c_less;
    p.evicted_kv_count[g] = completions(sm.less + sm.eq, r0, p.B) - c_less;  // completion keys == K
  }
}

__global__ void __launch_bounds__(256) select_radix_count_kernel(SelectParams p, RadixWs w) {
  __shared__ RadixSmem sm;
  radix_count_body(p, w, sm, blockIdx.x, nullptr);
}

// one CTA per sequence: grants at K go to the heads in order (same rule as select_threshold_kernel)
__device__ __forceinline__ void radix_grant_body(const SelectParams& p, RadixSmem& sm, int s) {
  int* s_scan = sm.scan;
  const int tid = threadIdx.x;
  const int LH = p.L * p.H;
  const int g0 = s * LH;
  const int E = p.budgets[s];
  if (E <= 0) return;  // counts already zeroed
  if (tid == 0) sm.carry = sm.lesssum = 0;
  __syncthreads();
  int local = 0;
  for (int hd = tid; hd < LH; hd += 256) local += p.evicted_block_count[g0 + hd];
  if (local) atomicAdd(&sm.lesssum, local);
  __syncthreads();
  const int need = E - sm.lesssum;  // grants left for completion keys == K (<= 0: K was inf, no ties exist)
  for (int h0 = 0; h0 < LH; h0 += 256) {
    const int hd = h0 + tid;
    const int eq = (hd < LH) ? p.evicted_kv_count[g0 + hd] : 0;
    s_scan[tid] = eq;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
      const int v = (tid >= d) ? s_scan[tid - d] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int eq_before = sm.carry + s_scan[tid] - eq;
    if (hd < LH) {
      const int g = g0 + hd;
      const int k = p.evicted_block_count[g] + max(0, min(eq, need - eq_before));
      p.evicted_block_count[g] = k;
      p.evicted_kv_count[g] = (k > 0) ? (k - 1) * p.B + p.hanging[g] : 0;  // metrics.py:788-792
    }
    __syncthreads();
    if (tid == 255) sm.carry += s_scan[255];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) select_radix_grant_kernel(SelectParams p) {
  __shared__ RadixSmem sm;
  radix_grant_body(p, sm, blockIdx.x);
}

// CTA per head: T = the head's cnt-th smallest key (radix select over the keys staged in shared memory), then
// the ordered emission of select_emit_cta_kernel.  staged: skeys already holds the head's keys (fused kernel).
__device__ __forceinline__ void radix_emit_body(const SelectParams& p, const RadixWs& w, RadixSmem& sm, int g,
                                                unsigned* skeys, bool staged, bool have_hist0 = false) {
  int* hist = sm.hist;
  int* s_w = sm.w;
  int* s_w2 = sm.w2;
  const int t = threadIdx.x, lane = t % kWarp, wid = t / kWarp;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int n = p.context_lens[(l * p.S + s) * p.H + h];
  const int nblk = (n + p.B - 1) / p.B;
  const int off = p.offsets[g];
  const int cnt = p.evicted_kv_count[g];
  int* __restrict__ out = p.evicted_logical_indices + off;
  if (cnt > 0) {  // CTA-uniform
    if (!staged) {
      const unsigned* __restrict__ keys = w.keys + off;
      if ((off & 3) == 0) {
        const uint4* __restrict__ k4 = reinterpret_cast<const uint4*>(keys);
        const int n4 = n >> 2;
        for (int i = t; i < n4; i += 256) reinterpret_cast<uint4*>(skeys)[i] = k4[i];
        for (int i = (n4 << 2) + t; i < n; i += 256) skeys[i] = keys[i];
      } else {
        for (int i = t; i < n; i += 256) skeys[i] = keys[i];
      }
    }
    __syncthreads();
    unsigned prefix = 0, mask = 0;
    int need = cnt;  // rank (1-based) among the keys that match the prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
      if (shift == 24 && have_hist0) {
        // top byte: pass 0's histogram of the finite keys; the non-candidates (key = inf) sit in bin 255
        hist[t] = sm.hist0[t] + ((t == 255) ? n - sm.finite : 0);
      } else {
        hist[t] = 0;
        __syncthreads();
        for (int i = t; i < n; i += 256) {
          const unsigned k = skeys[i];
          if ((k & mask) == prefix) hist_add(hist, (k >> shift) & 255u);
        }
      }
      __syncthreads();
      pick_bin(hist, need, s_w, &sm.prefix, &sm.need);
      prefix |= sm.prefix << shift;
      need = sm.need;
      mask |= 255u << shift;
      __syncthreads();
    }
    const unsigned T = prefix;  // cnt-th smallest key; `need` of the keys == T are evicted, in logical order
    const int allowed = need;
    // ordered emission: every warp owns a contiguous range of the logical indices; one CTA barrier for the
    // ranges' counts, then the warps emit on their own (ballot prefix sums, no barrier per 256 keys)
    const int chunk = ((n + 8 * kWarp - 1) / (8 * kWarp)) * kWarp;
    const int lo = min(n, wid * chunk), hi = min(n, lo + chunk);
    int cl = 0, ce = 0;
    for (int i = lo + lane; i < hi; i += kWarp) {
      const unsigned k = skeys[i];
      cl += k < T;
      ce += k == T;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      cl += __shfl_xor_sync(0xffffffffu, cl, d);
      ce += __shfl_xor_sync(0xffffffffu, ce, d);
    }
    if (lane == 0) { s_w[wid] = cl; s_w2[wid] = ce; }
    __syncthreads();
    int eq_base = 0, written = 0;
    for (int q = 0; q < wid; ++q) {
      written += s_w[q] + max(0, min(s_w2[q], allowed - eq_base));
      eq_base += s_w2[q];
    }
    for (int i0 = lo; i0 < hi; i0 += kWarp) {
      const int i = i0 + lane;
      const unsigned k = (i < hi) ? skeys[i] : kInfKey;
      const bool less = (i < hi) && k < T, eq = (i < hi) && (k == T);
      const unsigned em = __ballot_sync(0xffffffffu, eq);
      const bool ev = less || (eq && eq_base + __popc(em & lanemask_lt()) < allowed);
      const unsigned vm = __ballot_sync(0xffffffffu, ev);
      if (ev) out[written + __popc(vm & lanemask_lt())] = i;
      written += __popc(vm);
      eq_base += __popc(em);
    }
  }
  for (int i = cnt + t; i < nblk * p.B; i += 256) out[i] = p.null_index;
}

__global__ void __launch_bounds__(256) select_radix_emit_kernel(SelectParams p, RadixWs w) {
  extern __shared__ __align__(16) unsigned skeys[];
  __shared__ RadixSmem sm;
  radix_emit_body(p, w, sm, blockIdx.x, skeys, false);
}

// All of the above in ONE cooperative launch when every head's CTA is resident at once (post-prefill compression
// of a few prompts): the head's keys are built once and stay in shared memory for the four digit passes and the
// emission (the workspace copy and its five re-reads disappear), the kernel boundaries become grid barriers, and
// because a CTA keeps its histogram across a barrier the running count of keys below the prefix is carried from
// pass to pass: no per-key comparison with the prefix, and no count sweep at the end (keys < K = carried base +
// bins below the last digit, keys == K = that digit's bin).
struct FusedRadixState {
  unsigned prefix = 0;
  int base = 0;        // keys of the head below the prefix range
  bool alive = true;   // false: the sequence has fewer than E finite completion keys (K = inf)
};

template <int PASS>
__device__ __forceinline__ void radix_pass_fused(const SelectParams& p, const RadixWs& w, RadixSmem& sm, int g,
                                                 unsigned* skeys, int* s_bt, FusedRadixState& st) {
  int* hist = sm.hist;
  const int t = threadIdx.x, lane = t % kWarp, wid = t / kWarp;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int E = p.budgets[s];
  if (E <= 0) return;  // CTA-uniform
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int r0 = p.hanging[g];
  if (PASS > 0) {
    if (!st.alive) return;
    // the digit the previous pass decided (same for every head of the sequence); sm.scan = that pass's inclusive
    // bin sums of this head
    const int d = pick_digit(w.fsum + ((size_t)(PASS - 1) * p.S + s) * 256, E, &sm.tmp);
    if (d >= 256) { st.alive = false; return; }   // can only happen after pass 0
    st.base += (d > 0) ? sm.scan[d - 1] : 0;
    st.prefix |= (unsigned)d << (24 - 8 * (PASS - 1));
    __syncthreads();   // sm.scan is rewritten below
  }
  constexpr int shift = 24 - 8 * PASS;
  const unsigned hi_mask = PASS == 0 ? 0u : ~((1u << (shift + 8)) - 1u);
  hist[t] = 0;
  __syncthreads();
  if (PASS == 0) {
    const int last_pos = p.last_position[s];
    const int max_pos = last_pos - p.protected_window[s];
    const int* __restrict__ bt = p.block_tables + (int64_t)lsh * p.MB;
    const int nblk = (n + p.B - 1) / p.B;
    for (int b = t; b < nblk; b += 256) s_bt[b] = bt[b];
    __syncthreads();
    for (int i0 = t; i0 < n; i0 += 8 * 256) {
      float mm[8];
      int qq[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 256;
        const int64_t slot = (i < n) ? (int64_t)s_bt[i / p.B] * p.B + i % p.B : 0;
        mm[u] = p.kv_metrics[slot];
        qq[u] = p.kv_position[slot];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 256;
        if (i < n) {
          const unsigned k = slot_key(p, mm[u], qq[u], max_pos, last_pos, l, h);
          skeys[i] = k;
          if (k != kInfKey) hist_add(hist, k >> 24);
        }
      }
    }
  } else {
    const unsigned prefix = st.prefix;
    auto one = [&](unsigned k) {
      if (k != kInfKey && (k & hi_mask) == prefix) hist_add(hist, (k >> shift) & 255u);
    };
    const uint4* k4 = reinterpret_cast<const uint4*>(skeys);
    const int n4 = n >> 2;
    for (int i = t; i < n4; i += 256) {
      const uint4 v = k4[i];
      one(v.x); one(v.y); one(v.z); one(v.w);
    }
    for (int i = (n4 << 2) + t; i < n; i += 256) one(skeys[i]);
  }
  __syncthreads();
  // inclusive scan of the 256 bins, kept in sm.scan for the next phase
  int v = hist[t];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += u;
  }
  if (lane == 31) sm.w[wid] = v;
  __syncthreads();
  for (int q = 0; q < wid; ++q) v += sm.w[q];
  sm.scan[t] = v;
  if (PASS == 0) {
    sm.hist0[t] = hist[t];
    if (t == 255) sm.finite = v;
  }
  const int c = completions(st.base + v, r0, p.B);
  if (c) atomicAdd(w.fsum + ((size_t)PASS * p.S + s) * 256 + t, c);
  __syncthreads();
}

// after the last pass: blocks completed by keys < K and the extra ones completed by keys == K, without a sweep
__device__ __forceinline__ void radix_count_fused(const SelectParams& p, const RadixWs& w, RadixSmem& sm, int g,
                                                  FusedRadixState& st) {
  const int t = threadIdx.x;
  const int s = g / (p.L * p.H);
  const int E = p.budgets[s];
  if (E <= 0) {
    if (t == 0) { p.evicted_block_count[g] = 0; p.evicted_kv_count[g] = 0; }
    return;
  }
  int less, eq;
  if (st.alive) {
    const int d = pick_digit(w.fsum + ((size_t)3 * p.S + s) * 256, E, &sm.tmp);   // < 256: pass 0 found a digit
    less = st.base + ((d > 0) ? sm.scan[d - 1] : 0);
    eq = sm.hist[d];
  } else {   // K = inf: every finite key (sm.scan still holds pass 0's sums)
    less = sm.scan[255];
    eq = 0;
  }
  if (t == 0) {
    const int r0 = p.hanging[g];
    const int c_less = completions(less, r0, p.B);
    p.evicted_block_count[g] = c_less;
    p.evicted_kv_count[g] = completions(less + eq, r0, p.B) - c_less;  // completion keys == K
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 6) select_radix_fused_kernel(SelectParams p, RadixWs w, int stop_after) {
  extern __shared__ __align__(16) unsigned skeys[];
  __shared__ RadixSmem sm;
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int g = blockIdx.x;
  int* s_bt = reinterpret_cast<int*>(skeys + p.MB * p.B);   // the head's block-table row, behind its keys
  FusedRadixState st;
  // (profiling aid: KVC_SELECT_STOP_AFTER = n truncates the kernel after phase n, tools/select_long_phases.py)
  radix_pass_fused<0>(p, w, sm, g, skeys, s_bt, st);
  if (stop_after == 1) return;
  grid.sync();
  radix_pass_fused<1>(p, w, sm, g, skeys, s_bt, st);
  grid.sync();
  radix_pass_fused<2>(p, w, sm, g, skeys, s_bt, st);
  grid.sync();
  radix_pass_fused<3>(p, w, sm, g, skeys, s_bt, st);
  if (stop_after == 2) return;
  grid.sync();
  radix_count_fused(p, w, sm, g, st);
  grid.sync();
  if (g < p.S) radix_grant_body(p, sm, g);
  if (stop_after == 3) return;
  grid.sync();
  radix_emit_body(p, w, sm, g, skeys, true, true);
}

// one CTA per sequence
__global__ void __launch_bounds__(256) select_threshold_kernel(SelectParams p) {
  __shared__ int hist[256];
  __shared__ unsigned s_prefix;
  __shared__ int s_need;
  __shared__ int s_scan[256];
  __shared__ int s_carry;
  const int s = blockIdx.x, tid = threadIdx.x;
  const int LH = p.L * p.H;
  const int g0 = s * LH;
  const int E = p.budgets[s];
  const int e0 = p.offsets[g0] / p.B;
  int e1;
  if (s + 1 < p.S) {
    e1 = p.offsets[g0 + LH] / p.B;
  } else {  // last sequence: end of the last head's segment
    const int gl = p.S * LH - 1;
    const int l = ((gl / p.H) % p.L), h = gl % p.H;
    const int n = p.context_lens[(l * p.S + s) * p.H + h];
    e1 = p.offsets[gl] / p.B + (n + p.B - 1) / p.B;
  }
  const int M = e1 - e0;
  const unsigned* __restrict__ ev = p.events + e0;

  unsigned K = kInfKey;
  int need = E;  // rank (1-based) of the wanted key among keys matching the prefix
  if (E > 0 && E <= M) {  // E > M: fewer keys than the budget, K stays inf = take every finite key
    unsigned prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[tid] = 0;
      __syncthreads();
      for (int i = tid; i < M; i += 256) {
        const unsigned k = ev[i];
        if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 256; ++b) {
          if (acc + hist[b] >= need) break;
          acc += hist[b];
        }
        s_prefix = prefix | ((unsigned)b << shift);
        s_need = need - acc;
      }
      __syncthreads();
      prefix = s_prefix;
      need = s_need;
      mask |= 255u << shift;
      __syncthreads();
    }
    K = prefix;  // E-th smallest key (kInfKey if fewer than E finite keys exist)
  }
  // per head: k = #keys < K + share of the keys == K (in head order), or all finite keys if K is inf
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int h0 = 0; h0 < LH; h0 += 256) {
    const int hd = h0 + tid;
    int less = 0, eq = 0, r0 = 0;
    if (hd < LH) {
      const int g = g0 + hd;
      const int l = hd / p.H, h = hd % p.H;
      const int n = p.context_lens[(l * p.S + s) * p.H + h];
      const int nblk = (n + p.B - 1) / p.B;
      const unsigned* __restrict__ seg = p.events + p.offsets[g] / p.B;
      r0 = p.hanging[g];
      if (E > 0) {
        int lo = 0, hi = nblk;  // lower_bound(K)
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg[mid] < K) lo = mid + 1; else hi = mid; }
        less = lo;
        if (K != kInfKey) {
          hi = nblk;            // upper_bound(K)
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg[mid] <= K) lo = mid + 1; else hi = mid; }
          eq = lo - less;
        }
      }
    }
    // exclusive scan of eq over heads (ties at K are granted in head order)
    s_scan[tid] = eq;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
      const int v = (tid >= d) ? s_scan[tid - d] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int eq_before = s_carry + s_scan[tid] - eq;
    if (hd < LH) {
      const int take = max(0, min(eq, need - eq_before));
      const int k = less + ((K != kInfKey) ? take : 0);
      const int g = g0 + hd;
      p.evicted_block_count[g] = k;
      p.evicted_kv_count[g] = (k > 0) ? (k - 1) * p.B + r0 : 0;  // metrics.py:788-792
    }
    __syncthreads();
    if (tid == 255) s_carry += s_scan[255];
    __syncthreads();
  }
}


// =====================================================================================================
// One launch for the whole selection when a call carries many sequences of compressed contexts (decode-time
// continual compression: 256 sequences x 256 heads x ~9 blocks): ONE CTA PER SEQUENCE, a warp per head
// round-robin, the per-sequence top-E step done in the same CTA between two __syncthreads, and -- the point
// of it -- a PRUNED sort.  A head with r0 hanging tokens frees its first block once its r0 cheapest
// candidates go, so u = (r0-th smallest of the head's 32 per-lane minima) >= that block's completion key;
// T1 = the E-th smallest u over the sequence's heads is therefore >= the E-th smallest completion key K of
// the sequence, and no key above T1 can be evicted or decide a granted completion.  Only keys <= T1 (a
// handful per head in the steady state, where E ~ number of heads) are compacted and sorted, on a 32/64/...
// key network instead of the full 128/256.  Results are identical to the three-kernel path by construction
// (same keys, same tie rules, same threshold code on the same completion keys <= K).
//   A  load keys (DRAM), per-lane minimum -> 32-key network -> u, argmin (first index attaining the minimum)
//   B  T1 = E-th smallest u                                  (block radix select in shared memory)
//   C  reload keys (L2), keep <= T1, compact, small sort -> completion keys
//   D  K = E-th smallest completion key; per-head grants     (= select_threshold_kernel)
//   E  emit: a head granted its first block with r0 = 1 evicts exactly its argmin (no reload); other
//      granted heads reload and emit as select_emit_warp_kernel does
// =====================================================================================================
constexpr int kSeqWarps = 16;
constexpr int kSeqThreads = kSeqWarps * kWarp;

// need-th smallest (1-based) of vals[0..M) for the whole CTA; need in [1, M].  hist: 256 ints, sh: 2 words.
__device__ unsigned block_select_kth(const unsigned* vals, int M, int need, int* hist, unsigned* sh) {
  const int tid = threadIdx.x;
  unsigned prefix = 0, mask = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < M; i += blockDim.x) {
      const unsigned k = vals[i];
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < kWarp) {  // warp 0: 8 bins per lane, exclusive scan over lanes, first bin reaching `need`
      int loc[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = hist[tid * 8 + j]; sum += loc[j]; }
      int incl = sum;
#pragma unroll
      for (int d = 1; d < kWarp; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (tid >= d) incl += v;
      }
      int acc = incl - sum, found = -1, before = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (found < 0 && acc + loc[j] >= need) { found = tid * 8 + j; before = acc; }
        acc += loc[j];
      }
      const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
      const int src = bal ? __ffs(bal) - 1 : 0;
      const int b = __shfl_sync(0xffffffffu, found, src), ab = __shfl_sync(0xffffffffu, before, src);
      if (tid == 0) {
        sh[0] = prefix | ((unsigned)(bal ? b : 255) << shift);
        sh[1] = (unsigned)(need - (bal ? ab : 0));
      }
    }
    __syncthreads();
    prefix = sh[0];
    need = (int)sh[1];
    mask |= 255u << shift;
    __syncthreads();
  }
  return prefix;
}

template <int B, int N>
__device__ __forceinline__ void sort_row_write(const unsigned* row, int lane, int r0, int nev, unsigned* ev) {
  unsigned kk[N];
#pragma unroll
  for (int r = 0; r < N; ++r) kk[r] = row[lane * N + r];
  warp_bitonic_sort<N>(kk, lane);
  write_completion_keys<B, N>(kk, lane, r0, nev, ev);
}

// BTS: the sequence's block-table rows are staged in shared memory first (LH * MB words), so that every
// per-head step below is ONE round of independent loads instead of context length -> table -> data.
// KEEP2: phase A leaves every lane's two smallest keys (and the head's smallest "third") in shared memory;
// whenever that third is above T1 -- no lane holds three keys <= T1 -- phase C needs no second read of the
// metrics at all.  PLAIN: no head bias / averaging / sinks (the reference defaults): option tests compiled out.
// Lane <-> slot mapping in phases A / C: lane owns the 4-slot groups lane, lane + 32, ... (16 B of metrics and
// of positions each), which spreads a head's candidates over all lanes.
template <int KPL, bool BTS, bool KEEP2, bool PLAIN>
__global__ void __launch_bounds__(kSeqThreads, 2) select_fused_seq_kernel(SelectParams p, int stop_after) {
  constexpr int B = 16;
  constexpr int NG = KPL / 4;                // 4-slot groups per lane
  constexpr int kPair = (KPL == 8) ? 2 : 1;  // heads in flight per warp (register budget: 64)
  extern __shared__ unsigned dyn[];
  __shared__ int hist[256];
  __shared__ unsigned sh[2];
  __shared__ int s_wsum[kSeqWarps];
  __shared__ int s_carry;
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid % kWarp, w = tid / kWarp;
  const int LH = p.L * p.H;
  unsigned* s_first = dyn;                                   // [LH] upper bound of the first completion key
  int* s_argmin = reinterpret_cast<int*>(dyn + LH);          // [LH] first logical index attaining the minimum
  int* s_k = s_argmin + LH;                                  // [LH] granted blocks
  int* s_n = s_k + LH;                                       // [LH] context length
  int* s_r0 = s_n + LH;                                      // [LH] hanging tokens
  int* s_off = s_r0 + LH;                                    // [LH] segment offset
  unsigned* s_third = dyn + 6 * LH;                          // [LH] smallest third-smallest key of a lane (KEEP2)
  unsigned* s_min = dyn + 7 * LH;                            // [LH] smallest candidate key of the head
  unsigned* row = dyn + 8 * LH + w * (32 * KPL);             // per-warp compaction row
  unsigned* s_small = dyn + 8 * LH + kSeqWarps * 32 * KPL;   // [LH][64] two smallest keys per lane (KEEP2)
  int* s_bt = reinterpret_cast<int*>(s_small + (KEEP2 ? LH * 64 : 0));  // [LH][MB] (BTS)
  const int g0 = s * LH;
  const int E = p.budgets[s];
  const int last_pos = p.last_position[s];
  const int max_pos = last_pos - p.protected_window[s];

  for (int hd = tid; hd < LH; hd += kSeqThreads) {
    s_n[hd] = p.context_lens[((hd / p.H) * p.S + s) * p.H + hd % p.H];
    s_r0[hd] = p.hanging[g0 + hd];
    s_off[hd] = p.offsets[g0 + hd];
  }
  if (BTS) {
    if ((p.MB & 3) == 0) {  // rows are 16 B aligned: a thread copies a row with 16 B loads
      const int nv = p.MB >> 2;
      for (int i = tid; i < LH * nv; i += kSeqThreads) {
        const int hd = i / nv, v = i - hd * nv;
        const int4* __restrict__ src = reinterpret_cast<const int4*>(
            p.block_tables + (int64_t)(((hd / p.H) * p.S + s) * p.H + hd % p.H) * p.MB);
        reinterpret_cast<int4*>(s_bt + hd * p.MB)[v] = src[v];
      }
    } else {
      for (int i = tid; i < LH * p.MB; i += kSeqThreads) {
        const int hd = i / p.MB, b = i - hd * p.MB;
        s_bt[i] = p.block_tables[(int64_t)(((hd / p.H) * p.S + s) * p.H + hd % p.H) * p.MB + b];
      }
    }
  }
  __syncthreads();
  if (stop_after == 1) return;  // (profiling aid: KVC_SELECT_STOP_AFTER truncates the kernel after a phase)

  // keys of the lane's slots of head hd: k[4 j + r] <-> logical index 4 (lane + 32 j) + r.  The loads are
  // unconditional (lanes beyond the context read block 0) and separate from the key arithmetic, so the loads of
  // the next head can be in flight while this one is processed.
  struct Raw { float4 m[NG]; int4 q[NG]; };
  auto issue = [&](int hd, Raw& r) {
    const int n = s_n[hd];
    const int l = hd / p.H, h = hd % p.H;
    const int* __restrict__ bt = p.block_tables + (int64_t)((l * p.S + s) * p.H + h) * p.MB;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      const int i0 = 4 * (lane + 32 * j);
      const bool in = i0 < n;
      const int bi = in ? (i0 >> 4) : 0;
      const int phys = BTS ? s_bt[hd * p.MB + bi] : bt[bi];
      const int64_t slot = in ? (int64_t)phys * 16 + (i0 & 15) : 0;
      r.m[j] = *reinterpret_cast<const float4*>(p.kv_metrics + slot);
      r.q[j] = *reinterpret_cast<const int4*>(p.kv_position + slot);
    }
  };
  auto build = [&](int hd, const Raw& r, unsigned (&k)[KPL]) {
    const int n = s_n[hd];
    const int l = hd / p.H, h = hd % p.H;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      const int i0 = 4 * (lane + 32 * j);
      const float m[4] = {r.m[j].x, r.m[j].y, r.m[j].z, r.m[j].w};
      const int q[4] = {r.q[j].x, r.q[j].y, r.q[j].z, r.q[j].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned key = PLAIN ? slot_key_plain(m[t], q[t], max_pos) : slot_key(p, m[t], q[t], max_pos, last_pos, l, h);
        k[4 * j + t] = (i0 + t < n) ? key : kInfKey;
      }
    }
  };
  auto head_keys = [&](int hd, int n, unsigned (&k)[KPL]) {
    (void)n;
    Raw r;
    issue(hd, r);
    build(hd, r, k);
  };

  // keep keys <= T1, compact, sort on the smallest network that holds them, write the completion keys
  auto prune_sort_write = [&](auto& k, unsigned T1, int r0, int nblk, unsigned* ev) {
    constexpr int NK = sizeof(k) / sizeof(unsigned);
    int local = 0;
#pragma unroll
    for (int r = 0; r < NK; ++r) {
      if (k[r] > T1) k[r] = kInfKey;
      local += (k[r] != kInfKey);
    }
    int incl = local;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const int ncand = __shfl_sync(0xffffffffu, incl, 31);
    const int nev = (r0 >= 1 && ncand >= r0) ? (ncand - r0) / B + 1 : 0;
    if (nev > 0) {  // warp-uniform
      int size = 32;
      while (size < ncand) size <<= 1;
      int pos = incl - local;
#pragma unroll
      for (int r = 0; r < NK; ++r)
        if (k[r] != kInfKey) row[pos++] = k[r];
      for (int i = ncand + lane; i < size; i += kWarp) row[i] = kInfKey;
      __syncwarp();
      if (size == 32) sort_row_write<B, 1>(row, lane, r0, nev, ev);
      else if (size == 64) sort_row_write<B, 2>(row, lane, r0, nev, ev);
      else if (NK >= 4 && size == 128) sort_row_write<B, 4>(row, lane, r0, nev, ev);
      else if (NK >= 8 && size == 256) sort_row_write<B, 8>(row, lane, r0, nev, ev);
      else if (NK >= 16) sort_row_write<B, 16>(row, lane, r0, nev, ev);
      __syncwarp();
    }
    for (int q = nev + lane; q < nblk; q += kWarp) ev[q] = kInfKey;
  };

  unsigned T1 = kInfKey;
  if (E > 0) {
    // ---- A: bounds ----  (one head at a time, the next head's loads already in flight)
    Raw cur;
    if (w < LH) issue(w, cur);
    for (int hd = w; hd < LH; hd += kSeqWarps) {
      Raw nxt;
      const bool more = hd + kSeqWarps < LH;
      if (kPair == 2 && more) issue(hd + kSeqWarps, nxt);
      unsigned kk[KPL];
      build(hd, cur, kk);
      const int r0 = s_r0[hd];
      unsigned m1 = kInfKey, m2 = kInfKey, m3 = kInfKey;
      int idx = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < NG; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const unsigned k = kk[4 * j + r];
          if (k < m1) idx = 4 * (lane + 32 * j) + r;   // strict: the lane's first index wins ties
          const unsigned b1 = max(m1, k);
          m1 = min(m1, k);
          const unsigned b2 = max(m2, b1);
          m2 = min(m2, b1);
          m3 = min(m3, b2);
        }
      unsigned gm = m1;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) gm = min(gm, __shfl_xor_sync(0xffffffffu, gm, d));
      int amin = (m1 == gm) ? idx : 0x7fffffff;       // smallest logical index attaining the minimum
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) amin = min(amin, __shfl_xor_sync(0xffffffffu, amin, d));
      unsigned u = gm;                                // r0 == 1: the minimum itself
      if (r0 != 1) {
        unsigned sm1[1] = {m1};
        warp_bitonic_sort<1>(sm1, lane);
        u = (r0 >= 1 && r0 <= kWarp) ? __shfl_sync(0xffffffffu, sm1[0], (r0 - 1) & 31) : kInfKey;
      }
      if (KEEP2) {
        unsigned t3 = m3;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) t3 = min(t3, __shfl_xor_sync(0xffffffffu, t3, d));
        *reinterpret_cast<uint2*>(s_small + hd * 64 + 2 * lane) = make_uint2(m1, m2);
        if (lane == 0) s_third[hd] = t3;
      }
      if (lane == 0) { s_first[hd] = u; s_argmin[hd] = amin; s_min[hd] = gm; }
      if (kPair == 2) cur = nxt;
      else if (more) issue(hd + kSeqWarps, cur);
    }
    __syncthreads();
    if (stop_after == 2) return;
    // ---- B: T1 ----
    if (E <= LH) T1 = block_select_kth(s_first, LH, E, hist, sh);
  }
  if (stop_after == 3) return;

  // ---- C: pruned sort -> completion keys ----
  for (int hd = w; hd < LH; hd += kSeqWarps) {
    const int nblk = (s_n[hd] + B - 1) / B;
    unsigned* ev = p.events + s_off[hd] / B;
    if (E <= 0) {
      for (int q = lane; q < nblk; q += kWarp) ev[q] = kInfKey;
    } else if (KEEP2 && s_third[hd] > T1) {   // every key <= T1 is among the lanes' two smallest
      const uint2 v = *reinterpret_cast<const uint2*>(s_small + hd * 64 + 2 * lane);
      const int r0 = s_r0[hd];
      const int ncand = __popc(__ballot_sync(0xffffffffu, v.x <= T1 && v.x != kInfKey)) +
                        __popc(__ballot_sync(0xffffffffu, v.y <= T1 && v.y != kInfKey));
      if (r0 == 1 && ncand <= B) {
        // at most one completion key (rank 0 = the head's smallest candidate): no sort needed
        for (int q = lane; q < nblk; q += kWarp) ev[q] = (q == 0 && ncand >= 1) ? s_min[hd] : kInfKey;
      } else {
        unsigned k2[2] = {v.x, v.y};
        prune_sort_write(k2, T1, r0, nblk, ev);
      }
    } else {
      unsigned k[KPL];
      head_keys(hd, s_n[hd], k);
      prune_sort_write(k, T1, s_r0[hd], nblk, ev);
    }
  }
  __syncthreads();
  if (stop_after == 4) return;

  // ---- D: K and the per-head grants (select_threshold_kernel) ----
  const int e0 = s_off[0] / B;
  const int e1 = s_off[LH - 1] / B + (s_n[LH - 1] + B - 1) / B;   // segments of a sequence are contiguous
  const int M = e1 - e0;
  unsigned K = kInfKey;
  int need = E;
  if (E > 0 && E <= M) {
    K = block_select_kth(p.events + e0, M, E, hist, sh);
    need = (int)sh[1];   // how many of the keys equal to K are still wanted
    __syncthreads();
  }
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int h0 = 0; h0 < LH; h0 += kSeqThreads) {
    const int hd = h0 + tid;
    int less = 0, eq = 0, r0 = 0;
    if (hd < LH) {
      const int nblk = (s_n[hd] + B - 1) / B;
      const unsigned* __restrict__ seg = p.events + s_off[hd] / B;
      r0 = s_r0[hd];
      if (E > 0) {
        int lo = 0, hi = nblk;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg[mid] < K) lo = mid + 1; else hi = mid; }
        less = lo;
        if (K != kInfKey) {
          hi = nblk;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (seg[mid] <= K) lo = mid + 1; else hi = mid; }
          eq = lo - less;
        }
      }
    }
    // exclusive scan of eq in head order (ties at K are granted in head order)
    int incl = eq;
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == kWarp - 1) s_wsum[w] = incl;
    __syncthreads();
    int before = s_carry;
    for (int q = 0; q < w; ++q) before += s_wsum[q];
    const int eq_before = before + incl - eq;
    if (hd < LH) {
      const int take = max(0, min(eq, need - eq_before));
      const int k = less + ((K != kInfKey) ? take : 0);
      const int g = g0 + hd;
      s_k[hd] = k;
      p.evicted_block_count[g] = k;
      p.evicted_kv_count[g] = (k > 0) ? (k - 1) * B + r0 : 0;  // metrics.py:788-792
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < kSeqWarps; ++q) tot += s_wsum[q];
      s_carry += tot;
    }
    __syncthreads();
  }

  if (stop_after == 5) return;
  // ---- E: emit ----
  for (int hd = w; hd < LH; hd += kSeqWarps) {
    const int n = s_n[hd];
    const int off = s_off[hd];
    const int k = s_k[hd];
    const int r0 = s_r0[hd];
    const int cnt = (k > 0) ? (k - 1) * B + r0 : 0;
    if (cnt <= 1) {  // nothing, or the first block with one hanging token: the head's cheapest candidate goes
      // the segment starts at a multiple of B = 16 entries: 16 B stores
      int4* __restrict__ out4 = reinterpret_cast<int4*>(p.evicted_logical_indices + off);
      const int nvec = ((n + B - 1) / B) * (B / 4);
      const int nul = p.null_index;
      for (int i = lane; i < nvec; i += kWarp)
        out4[i] = make_int4((i == 0 && cnt == 1) ? s_argmin[hd] : nul, nul, nul, nul);
    } else {
      const int l = hd / p.H, h = hd % p.H;
      emit_head<B, KPL>(p, s, l, h, (l * p.S + s) * p.H + h, n, off, cnt, k, lane);
    }
  }
}

}  // namespace kvc

using namespace kvc;

static int g_select_mode = 0;
extern "C" int kvc_set_select_mode(int mode) {
  KVC_REQUIRE(mode >= 0 && mode <= 3, KVC_ERR_INVALID, "select mode must be 0 .. 3, got %d", mode);
  g_select_mode = mode;
  return KVC_OK;
}
extern "C" int kvc_get_select_mode(void) { return g_select_mode; }

// exact workspace of a call: completion keys (sort paths); for long contexts (more than 512 slots per head: the
// radix path) additionally one orderable key per slot and 4 x 256 threshold sums per sequence
extern "C" size_t kvc_select_evictions_workspace_bytes_ex(int64_t max_evicted_kv, int block_size, int num_seqs,
                                                          int max_num_blocks_per_seq) {
  if (max_evicted_kv <= 0 || block_size <= 0) return 0;
  size_t n = ((size_t)max_evicted_kv / block_size + 1) * sizeof(unsigned) + 512;
  if ((int64_t)max_num_blocks_per_seq * block_size > 512)
    n += ((size_t)max_evicted_kv + 64) * sizeof(unsigned) + 256 + (size_t)(num_seqs > 0 ? num_seqs : 0) * 4 * 256 * sizeof(int);
  return n;
}
// shape-agnostic upper bound (any context length, up to 1024 sequences per call)
extern "C" size_t kvc_select_evictions_workspace_bytes(int64_t max_evicted_kv, int block_size) {
  return kvc_select_evictions_workspace_bytes_ex(max_evicted_kv, block_size, 1024, 1 << 20);
}

extern "C" int kvc_select_evictions(
    int* evicted_logical_indices, int* evicted_kv_count, int* evicted_block_count,
    const float* kv_metrics, const int* kv_position, const int* block_tables, const int* context_lens,
    const int* hanging_token_count, const int* evicted_kv_offsets, const int* evicted_blocks_per_seq,
    const int* last_position, const int* protected_window_size, int num_seqs, int num_layers,
    int num_kv_heads, int max_num_blocks_per_seq, int block_size, int64_t max_evicted_kv,
    int null_eviction_index, void* workspace, size_t workspace_bytes, kvc_stream_t stream) {
  return kvc_select_evictions_ex(evicted_logical_indices, evicted_kv_count, evicted_block_count, kv_metrics,
                                 kv_position, block_tables, context_lens, hanging_token_count,
                                 evicted_kv_offsets, evicted_blocks_per_seq, last_position,
                                 protected_window_size, num_seqs, num_layers, num_kv_heads,
                                 max_num_blocks_per_seq, block_size, max_evicted_kv, null_eviction_index,
                                 nullptr, nullptr, 0, 0.f, 0, 0, workspace, workspace_bytes, stream);
}

extern "C" int kvc_select_evictions_ex(
    int* evicted_logical_indices, int* evicted_kv_count, int* evicted_block_count,
    const float* kv_metrics, const int* kv_position, const int* block_tables, const int* context_lens,
    const int* hanging_token_count, const int* evicted_kv_offsets, const int* evicted_blocks_per_seq,
    const int* last_position, const int* protected_window_size, int num_seqs, int num_layers,
    int num_kv_heads, int max_num_blocks_per_seq, int block_size, int64_t max_evicted_kv,
    int null_eviction_index, const float* head_bias, const int* position_bins, int num_bins,
    float bias_weight, int use_average, int num_sinks, void* workspace, size_t workspace_bytes,
    kvc_stream_t stream) {
  KVC_REQUIRE(block_size == 1 || block_size == 2 || block_size == 4 || block_size == 16,
              KVC_ERR_UNSUPPORTED, "Unsupported block size: %d", block_size);
  KVC_REQUIRE(evicted_logical_indices && evicted_kv_count && evicted_block_count && kv_metrics &&
                  kv_position && block_tables && context_lens && hanging_token_count &&
                  evicted_kv_offsets && evicted_blocks_per_seq && last_position && protected_window_size,
              KVC_ERR_INVALID, "select_evictions: null tensor");
  KVC_REQUIRE(num_seqs >= 0 && num_layers > 0 && num_kv_heads > 0 && max_num_blocks_per_seq > 0,
              KVC_ERR_INVALID, "select_evictions: bad sizes");
  if (num_seqs == 0) return KVC_OK;
  const size_t need = kvc_select_evictions_workspace_bytes_ex(max_evicted_kv, block_size, num_seqs,
                                                              max_num_blocks_per_seq);
  KVC_REQUIRE(workspace && workspace_bytes >= need, KVC_ERR_WORKSPACE,
              "select_evictions: workspace of %zu bytes required, got %zu", need, workspace_bytes);
  SelectParams p;
  p.evicted_logical_indices = evicted_logical_indices;
  p.evicted_kv_count = evicted_kv_count;
  p.evicted_block_count = evicted_block_count;
  p.kv_metrics = kv_metrics; p.kv_position = kv_position; p.block_tables = block_tables;
  p.context_lens = context_lens; p.hanging = hanging_token_count; p.offsets = evicted_kv_offsets;
  p.budgets = evicted_blocks_per_seq; p.last_position = last_position;
  p.protected_window = protected_window_size;
  p.S = num_seqs; p.L = num_layers; p.H = num_kv_heads; p.MB = max_num_blocks_per_seq;
  p.B = block_size; p.null_index = null_eviction_index;
  KVC_REQUIRE(head_bias == nullptr || (position_bins != nullptr && num_bins > 0), KVC_ERR_INVALID,
              "select_evictions: head_bias needs position_bins");
  p.head_bias = head_bias; p.position_bins = position_bins; p.num_bins = num_bins;
  p.bias_weight = bias_weight; p.use_average = use_average; p.num_sinks = num_sinks;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  p.events = reinterpret_cast<unsigned*>(w);
  cudaStream_t st = as_stream(stream);
  const int total = num_seqs * num_layers * num_kv_heads;
  const int max_slots = max_num_blocks_per_seq * block_size;
  const bool warp_path = max_slots <= 512;  // compressed contexts: one warp per head, 8 or 16 keys per lane
  const dim3 wgrid(ceil_div(total, 4));
  // many sequences of compressed contexts: one CTA per sequence does the whole selection with a pruned sort
  {
    const int LH = num_layers * num_kv_heads;
    const int mode = g_select_mode;  // 0 auto, 1 three kernels, 2 fused whenever the shape allows
    const bool can_fuse = warp_path && block_size == 16 && LH <= 8192;
    if (can_fuse && mode != 1 && (mode == 2 || num_seqs >= 96)) {
      const int kpl = max_slots <= 256 ? 8 : 16;
      const size_t base_words = (size_t)8 * LH + (size_t)kSeqWarps * 32 * kpl;
      const size_t bt_words = (size_t)LH * max_num_blocks_per_seq, small_words = (size_t)LH * 64;
      const size_t budget = 100 * 1024 / sizeof(unsigned);   // two CTAs per SM must still fit
      const bool keep2 = base_words + small_words <= budget;
      const bool bts = base_words + (keep2 ? small_words : 0) + bt_words <= budget;
      const size_t smem = (base_words + (keep2 ? small_words : 0) + (bts ? bt_words : 0)) * sizeof(unsigned);
      KVC_REQUIRE((reinterpret_cast<uintptr_t>(evicted_logical_indices) & 15) == 0, KVC_ERR_INVALID,
                  "select_evictions: evicted_logical_indices must be 16-byte aligned");
      cudaError_t e = cudaSuccess;
      const bool plain = head_bias == nullptr && !use_average && num_sinks <= 0;
      const char* stop_env = getenv("KVC_SELECT_STOP_AFTER");
      const int stop_after = stop_env ? atoi(stop_env) : 0;
#define KVC_SEL_FUSED(KPL_, BTS_, KEEP_, PLAIN_)                                                       \
  do {                                                                                                 \
    auto kern = select_fused_seq_kernel<KPL_, BTS_, KEEP_, PLAIN_>;                                    \
    if (smem > 48 * 1024)                                                                              \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e == cudaSuccess) kern<<<num_seqs, kSeqThreads, smem, st>>>(p, stop_after);                    \
  } while (0)
#define KVC_SEL_FUSED3(KPL_, BTS_, KEEP_)                                                              \
  do {                                                                                                 \
    if (plain) KVC_SEL_FUSED(KPL_, BTS_, KEEP_, true); else KVC_SEL_FUSED(KPL_, BTS_, KEEP_, false);   \
  } while (0)
#define KVC_SEL_FUSED2(KPL_)                                                                           \
  do {                                                                                                 \
    if (bts && keep2) KVC_SEL_FUSED3(KPL_, true, true);                                                \
    else if (bts) KVC_SEL_FUSED3(KPL_, true, false);                                                   \
    else if (keep2) KVC_SEL_FUSED3(KPL_, false, true);                                                 \
    else KVC_SEL_FUSED3(KPL_, false, false);                                                           \
  } while (0)
      if (kpl == 8) KVC_SEL_FUSED2(8); else KVC_SEL_FUSED2(16);
#undef KVC_SEL_FUSED3
#undef KVC_SEL_FUSED2
#undef KVC_SEL_FUSED
      if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "select_fused_seq smem attr: %s", cudaGetErrorString(e));
      KVC_CHECK_LAUNCH("select_fused_seq_kernel");
      return KVC_OK;
    }
  }
#define KVC_SEL_WARP(KERNEL)                                                     \
  do {                                                                           \
    if (max_slots <= 256) {                                                      \
      if (block_size == 16) KERNEL<16, 8><<<wgrid, 128, 0, st>>>(p);             \
      else if (block_size == 4) KERNEL<4, 8><<<wgrid, 128, 0, st>>>(p);          \
      else if (block_size == 2) KERNEL<2, 8><<<wgrid, 128, 0, st>>>(p);          \
      else KERNEL<1, 8><<<wgrid, 128, 0, st>>>(p);                               \
    } else {                                                                     \
      if (block_size == 16) KERNEL<16, 16><<<wgrid, 128, 0, st>>>(p);            \
      else if (block_size == 4) KERNEL<4, 16><<<wgrid, 128, 0, st>>>(p);         \
      else if (block_size == 2) KERNEL<2, 16><<<wgrid, 128, 0, st>>>(p);         \
      else KERNEL<1, 16><<<wgrid, 128, 0, st>>>(p);                              \
    }                                                                            \
  } while (0)
  if (!warp_path && g_select_mode != 1 && max_slots * sizeof(unsigned) <= 200 * 1024) {
    // long contexts: radix descent on histograms, no sort (mode 1 keeps the sort path for A/B and tests)
    RadixWs rw;
    uintptr_t kp = (w + ((size_t)max_evicted_kv / block_size + 1) * sizeof(unsigned) + 255) & ~uintptr_t(255);
    rw.keys = reinterpret_cast<unsigned*>(kp);
    rw.fsum = reinterpret_cast<int*>((kp + ((size_t)max_evicted_kv + 64) * sizeof(unsigned) + 255) & ~uintptr_t(255));
    cudaError_t e = cudaMemsetAsync(rw.fsum, 0, (size_t)num_seqs * 4 * 256 * sizeof(int), st);
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "select_evictions memset: %s", cudaGetErrorString(e));
    const size_t smem_bt = (size_t)max_num_blocks_per_seq * sizeof(int);
    const size_t smem_keys = (size_t)max_slots * sizeof(unsigned) + smem_bt;   // fused: keys + block-table row
    {
      // one cooperative launch when all head CTAs fit on the device at once (select mode 3: always the
      // seven-launch path, which larger batches take anyway)
      const bool want_fused = g_select_mode != 3;
      int per_sm = 0;
      if (want_fused) {
        e = cudaFuncSetAttribute(select_radix_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_keys);
        if (e == cudaSuccess)
          e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_radix_fused_kernel, 256, smem_keys);
        if (e != cudaSuccess) { (void)cudaGetLastError(); per_sm = 0; }
      }
      if (per_sm > 0 && (int64_t)per_sm * sm_count() >= total) {
        static const int stop_env = [] { const char* sa = getenv("KVC_SELECT_STOP_AFTER"); return sa ? atoi(sa) : 0; }();
        int stop_after = stop_env;
        void* args[] = {&p, &rw, &stop_after};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(select_radix_fused_kernel), dim3(total), dim3(256),
                                        args, smem_keys, st);
        if (e == cudaSuccess) {
          KVC_CHECK_LAUNCH("select_radix_fused_kernel");
          return KVC_OK;
        }
        // e.g. cudaErrorCooperativeLaunchTooLarge under a partitioned device: the seven launches below need no
        // co-residency
        (void)cudaGetLastError();
      }
    }
    select_radix_pass_kernel<0><<<total, 256, smem_bt, st>>>(p, rw);
    select_radix_pass_kernel<1><<<total, 256, 0, st>>>(p, rw);
    select_radix_pass_kernel<2><<<total, 256, 0, st>>>(p, rw);
    select_radix_pass_kernel<3><<<total, 256, 0, st>>>(p, rw);
    select_radix_count_kernel<<<total, 256, 0, st>>>(p, rw);
    select_radix_grant_kernel<<<num_seqs, 256, 0, st>>>(p);
    const size_t smem = (size_t)max_slots * sizeof(unsigned);
    if (smem > 40 * 1024) {
      e = cudaFuncSetAttribute(select_radix_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "select_radix_emit smem attr: %s", cudaGetErrorString(e));
    }
    select_radix_emit_kernel<<<total, 256, smem, st>>>(p, rw);
    KVC_CHECK_LAUNCH("select_radix kernels");
    return KVC_OK;
  }
  if (warp_path) {
    KVC_SEL_WARP(select_sort_kernel);
    KVC_CHECK_LAUNCH("select_sort_kernel");
  } else if (max_slots <= 32768) {
    int m = 64;
    while (m < max_slots) m <<= 1;
    const size_t smem = (size_t)m * sizeof(unsigned);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(select_sort_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "select_sort_cta smem attr: %s", cudaGetErrorString(e));
    }
    select_sort_cta_kernel<<<total, kSortThreads, smem, st>>>(p);
    KVC_CHECK_LAUNCH("select_sort_cta_kernel");
  } else {
    const size_t smem = (size_t)max_slots * sizeof(unsigned);
    KVC_REQUIRE(smem <= 220 * 1024, KVC_ERR_UNSUPPORTED,
                "select_evictions: %d slots per head exceed the shared-memory staging limit", max_slots);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(select_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "select_rank smem attr: %s", cudaGetErrorString(e));
    }
    select_rank_kernel<<<total, 128, smem, st>>>(p, max_slots);
    KVC_CHECK_LAUNCH("select_rank_kernel");
  }
  select_threshold_kernel<<<num_seqs, 256, 0, st>>>(p);
  KVC_CHECK_LAUNCH("select_threshold_kernel");
  if (warp_path) {
    KVC_SEL_WARP(select_emit_warp_kernel);
  } else {
    select_emit_cta_kernel<<<total, 128, 0, st>>>(p);
  }
#undef KVC_SEL_WARP
  KVC_CHECK_LAUNCH("select_emit_kernel");
  return KVC_OK;
}
