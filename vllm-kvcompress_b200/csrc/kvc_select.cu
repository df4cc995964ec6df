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

template <int B, int KPL>
__global__ void __launch_bounds__(128) select_emit_warp_kernel(SelectParams p) {
  const int lane = threadIdx.x % kWarp;
  const int g = blockIdx.x * 4 + threadIdx.x / kWarp;
  if (g >= p.S * p.L * p.H) return;
  const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
  const int lsh = (l * p.S + s) * p.H + h;
  const int n = p.context_lens[lsh];
  const int nblk = (n + B - 1) / B;
  const int off = p.offsets[g];
  const int cnt = p.evicted_kv_count[g];
  const int kblk = p.evicted_block_count[g];
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

}  // namespace kvc

using namespace kvc;

extern "C" size_t kvc_select_evictions_workspace_bytes(int64_t max_evicted_kv, int block_size) {
  if (max_evicted_kv <= 0 || block_size <= 0) return 0;
  return ((size_t)max_evicted_kv / block_size + 1) * sizeof(unsigned) + 512;  // completion keys
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
  const size_t need = kvc_select_evictions_workspace_bytes(max_evicted_kv, block_size);
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
