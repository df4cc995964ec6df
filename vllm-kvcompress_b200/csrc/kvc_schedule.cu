// Eviction / move scheduling kernels (integer work, bit-exact vs the reference).
//
//   K1 schedule_cache_evictions   reference csrc/kvcompress_eviction_kernels.cu:27-156
//   K2 truncate_cache_evictions   reference :158-188
//   K3 count_block_evictions      reference :190-221
//   K4 schedule_t1_cache_moves    reference :223-289
//
// The reference runs K1 as ONE thread per sequence walking up to 1.5 M sorted
// entries with six dependent loads each, and K2-K4 as one thread per head.
// Here every op is re-expressed as a data-parallel problem:
//   K1 = stable multi-bin ranking of the sorted list by head id (chunk histograms
//        -> per-head exclusive scan over chunks -> first chunk where the number of
//        completed blocks reaches the budget -> stable scatter with an in-chunk cut)
//   K2 = segmented fill, one warp per head segment (+ the whole grid on the tail
//        of the last head, which runs to max_evicted_kv)
//   K3 = one warp per head, ballot over segment rows
//   K4 = closed form: dst_k = k-th smallest evicted index below new_n,
//        src_k = k-th largest surviving index >= new_n; one warp per head, a
//        1024-bit occupancy window per iteration, popc prefix ranks.
// All of it is HBM/L2-latency-bound int32 traffic; no tensor-core shapes here.
#include "kvc_common.cuh"

namespace kvc {

// =============================================================================
// K2  truncate
// =============================================================================
__device__ __forceinline__ int truncated_count(int evicted_kv, int hanging_kv, int B) {
  // reference :171-178
  return (evicted_kv < hanging_kv) ? 0 : evicted_kv - (evicted_kv - hanging_kv) % B;
}

// Pass 1: one thread per head rewrites the count (idempotent, see pass 2).
__global__ void truncate_counts_kernel(int* __restrict__ evicted_kv_count,
                                       const int* __restrict__ hanging_token_count,
                                       int total_heads, int B) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_heads) return;
  evicted_kv_count[g] = truncated_count(evicted_kv_count[g], hanging_token_count[g], B);
}

// Pass 2: one warp per head fills [off+cnt', next_off) with (g, NULL).  Extra warps
// (global warp id >= total_heads) share the last head's tail, which the reference
// extends to max_evicted_kv (:183) and which can be millions of entries long.
__global__ void truncate_fill_kernel(int* __restrict__ evicted_kv_indices,
                                     int* __restrict__ evicted_logical_indices,
                                     const int* __restrict__ evicted_kv_count,  // already truncated
                                     const int* __restrict__ evicted_kv_offsets,
                                     int total_heads, int max_evicted_kv, int null_idx,
                                     int tail_warps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp < total_heads - 1) {
    const int start = evicted_kv_offsets[warp] + evicted_kv_count[warp];
    const int end = evicted_kv_offsets[warp + 1];
    for (int i = start + lane; i < end; i += kWarp) {
      evicted_kv_indices[i] = warp;
      evicted_logical_indices[i] = null_idx;
    }
  } else if (warp < total_heads - 1 + tail_warps) {
    const int g = total_heads - 1;
    const int t = warp - g;  // 0 .. tail_warps-1
    const int start = evicted_kv_offsets[g] + evicted_kv_count[g];
    for (int i = start + t * kWarp + lane; i < max_evicted_kv; i += tail_warps * kWarp) {
      evicted_kv_indices[i] = g;
      evicted_logical_indices[i] = null_idx;
    }
  }
}

static int launch_truncate(int* evicted_kv_indices, int* evicted_logical_indices,
                           int* evicted_kv_count, const int* evicted_kv_offsets,
                           const int* hanging_token_count, int total_heads, int B,
                           int max_evicted_kv, int null_idx, cudaStream_t stream) {
  if (total_heads <= 0) return KVC_OK;
  truncate_counts_kernel<<<ceil_div(total_heads, 256), 256, 0, stream>>>(
      evicted_kv_count, hanging_token_count, total_heads, B);
  KVC_CHECK_LAUNCH("truncate_counts_kernel");
  const int tail_warps = sm_count() * 8;  // one wave of 256-thread CTAs for the open-ended tail
  const int warps = total_heads - 1 + tail_warps;
  truncate_fill_kernel<<<ceil_div(warps, 8), 256, 0, stream>>>(
      evicted_kv_indices, evicted_logical_indices, evicted_kv_count, evicted_kv_offsets,
      total_heads, max_evicted_kv, null_idx, tail_warps);
  KVC_CHECK_LAUNCH("truncate_fill_kernel");
  return KVC_OK;
}

// =============================================================================
// K3  count_block_evictions: one warp per head
// =============================================================================
__global__ void count_block_evictions_kernel(int* __restrict__ evicted_block_count,
                                             int* __restrict__ evicted_logical_indices,
                                             const int* __restrict__ evicted_kv_offsets,
                                             const int* __restrict__ hanging_token_count,
                                             int total_heads, int total_kvs, int B,
                                             int null_value) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (g >= total_heads) return;
  const int start = evicted_kv_offsets[g];
  const int end = (g + 1 >= total_heads) ? total_kvs : evicted_kv_offsets[g + 1];
  // rows r = 0,1,.. start at start + r*B while < end (reference loop :201-207)
  const int rows = (end > start) ? (end - start + B - 1) / B : 0;
  int evicted_blocks = rows;  // all rows non-NULL unless a NULL row is found
  for (int r0 = 0; r0 < rows; r0 += kWarp) {
    const int r = r0 + lane;
    const bool is_null = (r < rows) && (evicted_logical_indices[start + (int64_t)r * B] == null_value);
    const unsigned m = __ballot_sync(0xffffffffu, is_null);
    if (m) {
      evicted_blocks = r0 + __ffs(m) - 1;
      break;
    }
  }
  if (lane == 0) evicted_block_count[g] = evicted_blocks;
  if (evicted_blocks > 0) {  // reference :210-219
    const int last_block_end = start + evicted_blocks * B;
    for (int i = last_block_end - B + hanging_token_count[g] + lane; i < last_block_end; i += kWarp)
      evicted_logical_indices[i] = null_value;
  }
}

// =============================================================================
// K4  schedule_t1_cache_moves: one warp per head
// =============================================================================
template <int B>
__device__ __forceinline__ void schedule_moves_head(
    int* __restrict__ cache_moves_idx,                  // [max_evicted_kv, 2]
    int* __restrict__ cache_moves_count,                // [S,L,H]
    const int* __restrict__ evicted_logical_indices,    // [max_evicted_kv]
    const int* __restrict__ evicted_kv_count,           // [S,L,H]
    const int* __restrict__ evicted_kv_offsets,         // [S,L,H]
    const int* __restrict__ block_tables,               // [L,S,H,MB]
    const int* __restrict__ context_lens,               // [L,S,H]
    int num_seqs, int num_layers, int num_kv_heads, int MB, unsigned (*occ)[kWarp], int wid, int lane, int g) {
  const int s = g / (num_layers * num_kv_heads);
  const int l = (g / num_kv_heads) % num_layers;
  const int h = g % num_kv_heads;
  const int lsh = (l * num_seqs + s) * num_kv_heads + h;
  const int* __restrict__ bt = block_tables + (int64_t)lsh * MB;
  const int n = context_lens[lsh];
  const int c = evicted_kv_count[g];
  const int off = evicted_kv_offsets[g];
  const int* __restrict__ E = evicted_logical_indices + off;
  const int new_n = n - c;

  // m = |E ∩ [0,new_n)| = lower_bound(E[0..c), new_n)   (E ascending, reference host sort
  // vllm/kvcompress/scheduler.py:472-477 / metrics.py:822-834)
  int lo = 0, hi = c;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (E[mid] < new_n) lo = mid + 1; else hi = mid;
  }
  const int m = lo;
  if (lane == 0) cache_moves_count[g] = m;
  if (m == 0) return;

  // Tail region [new_n, n) in reverse: r = n-1-p, r in [0, c).  Evicted tail entries are
  // E[m..c); E[j] has r = n-1-E[j], descending in j.  Survivor with reverse index r gets
  // move rank k = r - #evicted{r' < r}; emit (phys(E[k]), phys(n-1-r)).
  int ev_before = 0;   // evicted tail entries with r' < window start
  int j_hi = c;        // evicted tail entries E[j_hi..c) already consumed (they have the smallest r)
  for (int w0 = 0; w0 < c; w0 += 1024) {
    const int wbits = min(1024, c - w0);
    occ[wid][lane] = 0u;
    __syncwarp();
    // entries of E with r in [w0, w0+wbits)  <=>  E[j] in (n-1-w0-wbits, n-1-w0]
    const int p_lo = n - 1 - w0 - wbits;  // exclusive lower bound on p
    int a = m, b = j_hi;                  // find first j in [m, j_hi) with E[j] > p_lo
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (E[mid] > p_lo) b = mid; else a = mid + 1;
    }
    const int j_lo = a;
    for (int j = j_lo + lane; j < j_hi; j += kWarp) {
      const int r = n - 1 - E[j] - w0;
      atomicOr(&occ[wid][r >> 5], 1u << (r & 31));
    }
    __syncwarp();
    const unsigned word = occ[wid][lane];
    const int nbits = max(0, min(32, wbits - lane * 32));
    const unsigned valid = (nbits == 32) ? 0xffffffffu : ((1u << nbits) - 1u);
    const int cnt = __popc(word & valid);
    int incl = cnt;  // inclusive warp scan of evicted counts
#pragma unroll
    for (int d = 1; d < kWarp; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int ev = ev_before + incl - cnt;  // evicted with smaller r than this lane's word
    unsigned surv = (~word) & valid;
    while (surv) {
      const int bit = __ffs(surv) - 1;
      surv &= surv - 1;
      const int r = w0 + lane * 32 + bit;
      const int k = r - (ev + __popc(word & ((1u << bit) - 1u)));
      if (k < m) {
        const int src = n - 1 - r;
        const int dst = E[k];
        const int src_phys = bt[src / B] * B + src % B;
        const int dst_phys = bt[dst / B] * B + dst % B;
        reinterpret_cast<int2*>(cache_moves_idx)[(int64_t)off + k] = make_int2(dst_phys, src_phys);
      }
    }
    ev_before += __shfl_sync(0xffffffffu, incl, kWarp - 1);
    j_hi = j_lo;
    __syncwarp();
  }
}



// one warp per head (few heads, long eviction lists: post-prefill compression)
template <int B>
__global__ void __launch_bounds__(128) schedule_moves_kernel(
    int* __restrict__ cache_moves_idx, int* __restrict__ cache_moves_count,
    const int* __restrict__ evicted_logical_indices, const int* __restrict__ evicted_kv_count,
    const int* __restrict__ evicted_kv_offsets, const int* __restrict__ block_tables,
    const int* __restrict__ context_lens, int num_seqs, int num_layers, int num_kv_heads, int MB) {
  __shared__ unsigned occ[4][kWarp];  // per-warp 1024-bit window of evicted tail positions
  const int wid = threadIdx.x / kWarp;
  const int lane = threadIdx.x % kWarp;
  const int g = blockIdx.x * (blockDim.x / kWarp) + wid;  // (s, l, h) flat
  if (g >= num_seqs * num_layers * num_kv_heads) return;
  schedule_moves_head<B>(cache_moves_idx, cache_moves_count, evicted_logical_indices, evicted_kv_count,
                         evicted_kv_offsets, block_tables, context_lens, num_seqs, num_layers, num_kv_heads, MB,
                         occ, wid, lane, g);
}

// one CTA per head (few heads, thousands of evictions each: post-prefill compression).  No window bitmaps and no
// binary searches: with the evicted tail entries T_0 > T_1 > ... (E read backwards, T_{-1} = n, and the sentinel
// new_n - 1 after the last one), the survivors between T_{i-1} and T_i have exactly i evicted positions above them,
// so the survivor at position pos is move number k = (n - 1 - pos) - i.  A thread takes the gaps i = tid, tid + 256,
// ...: two coalesced reads of E per gap, and work only where a gap is not empty.
template <int B>
__global__ void __launch_bounds__(256) schedule_moves_cta_kernel(
    int* __restrict__ cache_moves_idx, int* __restrict__ cache_moves_count,
    const int* __restrict__ evicted_logical_indices, const int* __restrict__ evicted_kv_count,
    const int* __restrict__ evicted_kv_offsets, const int* __restrict__ block_tables,
    const int* __restrict__ context_lens, int num_seqs, int num_layers, int num_kv_heads, int MB) {
  __shared__ int s_m;
  const int g = blockIdx.x, t = threadIdx.x;
  const int s = g / (num_layers * num_kv_heads);
  const int l = (g / num_kv_heads) % num_layers;
  const int h = g % num_kv_heads;
  const int lsh = (l * num_seqs + s) * num_kv_heads + h;
  const int* __restrict__ bt = block_tables + (int64_t)lsh * MB;
  const int n = context_lens[lsh];
  const int c = evicted_kv_count[g];
  const int off = evicted_kv_offsets[g];
  const int* __restrict__ E = evicted_logical_indices + off;
  const int new_n = n - c;
  if (t == 0) s_m = 0;
  __syncthreads();
  // m = |E ∩ [0, new_n)|
  int cnt = 0;
  for (int j = t; j < c; j += 256) cnt += E[j] < new_n;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if ((t & 31) == 0 && cnt) atomicAdd(&s_m, cnt);
  __syncthreads();
  const int m = s_m;
  if (t == 0) cache_moves_count[g] = m;
  if (m == 0) return;
  const int ngaps = c - m + 1;   // gap i lies below the i-th largest evicted tail entry
  for (int i = t; i < ngaps; i += 256) {
    const int hi = (i == 0) ? n : E[c - i];                    // exclusive
    const int lo = (i == c - m) ? new_n - 1 : E[c - 1 - i];    // exclusive
    for (int pos = hi - 1; pos > lo; --pos) {
      const int k = (n - 1 - pos) - i;
      const int dst = E[k];
      const int src_phys = bt[pos / B] * B + pos % B;
      const int dst_phys = bt[dst / B] * B + dst % B;
      reinterpret_cast<int2*>(cache_moves_idx)[(int64_t)off + k] = make_int2(dst_phys, src_phys);
    }
  }
}

// many heads with 0 or 1 evicted slot each (decode-time continual compression): a warp looks at 32 heads, every
// lane settles its own head when c <= 1 (closed form: the one evicted index E0 takes the last token unless it
// IS the last token), the heads with longer lists are then done by the whole warp one after the other
template <int B>
__global__ void __launch_bounds__(128) schedule_moves_many_kernel(
    int* __restrict__ cache_moves_idx, int* __restrict__ cache_moves_count,
    const int* __restrict__ evicted_logical_indices, const int* __restrict__ evicted_kv_count,
    const int* __restrict__ evicted_kv_offsets, const int* __restrict__ block_tables,
    const int* __restrict__ context_lens, int num_seqs, int num_layers, int num_kv_heads, int MB) {
  __shared__ unsigned occ[4][kWarp];
  const int wid = threadIdx.x / kWarp;
  const int lane = threadIdx.x % kWarp;
  const int total = num_seqs * num_layers * num_kv_heads;
  const int base = (blockIdx.x * (blockDim.x / kWarp) + wid) * kWarp;
  if (base >= total) return;
  const int g = base + lane;
  int c = 0;
  if (g < total) {
    c = evicted_kv_count[g];
    if (c <= 1) {
      int m = 0;
      if (c == 1) {
        const int s = g / (num_layers * num_kv_heads);
        const int l = (g / num_kv_heads) % num_layers;
        const int h = g % num_kv_heads;
        const int lsh = (l * num_seqs + s) * num_kv_heads + h;
        const int off = evicted_kv_offsets[g];
        const int n = context_lens[lsh];
        const int e0 = evicted_logical_indices[off];
        if (e0 < n - 1) {  // dst = E[0] < new_n = n - 1, src = the last token
          const int* __restrict__ bt = block_tables + (int64_t)lsh * MB;
          const int src = n - 1;
          reinterpret_cast<int2*>(cache_moves_idx)[(int64_t)off] =
              make_int2(bt[e0 / B] * B + e0 % B, bt[src / B] * B + src % B);
          m = 1;
        }
      }
      cache_moves_count[g] = m;
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, c > 1);
  while (todo) {
    const int l0 = __ffs(todo) - 1;
    todo &= todo - 1;
    schedule_moves_head<B>(cache_moves_idx, cache_moves_count, evicted_logical_indices, evicted_kv_count,
                           evicted_kv_offsets, block_tables, context_lens, num_seqs, num_layers, num_kv_heads, MB,
                           occ, wid, lane, base + l0);
    __syncwarp();
  }
}

// =============================================================================
// K1  schedule_cache_evictions as a chunked stable multi-bin rank
// =============================================================================
// Reference semantics (one thread per sequence, :63-155): walk the sequence's
// metric-sorted slot list; every valid slot (inside its head's context, outside
// the protected window) is appended to its head's segment; a head "completes a
// block" at its r0-th, (r0+B)-th, ... appended slot (r0 = initial
// remaining_kv_count = hanging tokens); stop right after the E-th completion.
//
// Parallel form.  Let c_h(i) = #valid slots of head h among the first i entries.
// The slot at sorted position i goes to off[h] + c_h(i) (a stable partition by
// head = one radix pass with L*H bins), and the walk stops at the first i where
//   F(i) = sum_h comp(c_h(i)) >= E,   comp(c) = c >= r0 ? (c - r0)/B + 1 : 0.
// Chunks of kChunk entries: (A) per-chunk per-head histograms, (B) exclusive scan
// over chunks per head + first chunk k* with F >= E, (C) stable in-chunk ranks,
// an ordered event scan inside chunk k* to find the exact cut, scatter.
// With evict_evenly_per_layer each layer is its own group (budget E/L, own cut).

constexpr int kChunk = 1024;         // entries per chunk
constexpr int kK1Threads = 256;      // 8 warps, each owns 128 consecutive entries (4 steps of 32)
constexpr int kNoCut = 0x7fffffff;

struct K1Params {
  int* evicted_kv_indices;
  int* evicted_logical_indices;
  int* evicted_kv_count;
  int* remaining_kv_count;
  const int* evicted_kv_offsets;
  const int* sorted_indices;
  const int* seq_block_offsets;
  const int* layer_by_block;
  const int* head_by_block;
  const int* virtual_block_num_by_block;
  const int* evicted_blocks_per_seq;
  const int* context_lens;
  const int* kv_position;
  const int* last_position;
  const int* protected_window_size;
  const int* control_layers;
  int num_control_layers;
  int S, L, H, total_blocks, B;
  int evenly;        // evict_evenly_per_layer
  // workspace
  int* chunk_base;   // [S+1]
  int* hist;         // [max_chunks, L*H]
  int* prefix;       // [max_chunks, L*H]
  int* kstar;        // [S, NG]
  int* erem;         // [S, NG]
  int max_chunks;
};

__device__ __forceinline__ int k1_num_groups(const K1Params& p) { return p.evenly ? p.L : 1; }

__device__ __forceinline__ int k1_group_budget(const K1Params& p, int s, int layer) {
  int e = p.evicted_blocks_per_seq[s];
  if (!p.evenly) return e;
  for (int i = 0; i < p.num_control_layers; ++i)
    if (p.control_layers[i] == layer) return 0;  // reference :55-61: control layers never evict
  return e / p.L;                                 // reference :76-78
}

__device__ __forceinline__ int k1_completions(int c, int r0, int B) {
  return (r0 >= 1 && c >= r0) ? (c - r0) / B + 1 : 0;
}
__device__ __forceinline__ int k1_remaining(int c, int r0, int B) {
  if (r0 < 1 || c < r0) return r0 - c;
  return B - ((c - r0) % B);
}

__device__ __forceinline__ void k1_seq_range(const K1Params& p, int s, int& start_blk, int& ntok) {
  start_blk = p.seq_block_offsets[s];
  const int end_blk = (s + 1 >= p.S) ? p.total_blocks : p.seq_block_offsets[s + 1];  // :73
  ntok = max(0, end_blk - start_blk) * p.B;
}

// (A0) chunk map: chunk_base[s] = sum_{s'<s} ceil(T_s'/kChunk).  One CTA, serial scan in
// thread 0 is enough for the sequence counts this op sees (<= a few thousand).
__global__ void k1_setup_kernel(K1Params p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int acc = 0;
  for (int s = 0; s < p.S; ++s) {
    int sb, ntok;
    k1_seq_range(p, s, sb, ntok);
    p.chunk_base[s] = acc;
    acc += (ntok + kChunk - 1) / kChunk;
  }
  p.chunk_base[p.S] = acc;
}

// map a flat chunk id to (sequence, chunk-in-sequence); false if beyond the last chunk
__device__ __forceinline__ bool k1_locate(const K1Params& p, int chunk, int& s, int& k) {
  if (chunk >= p.chunk_base[p.S]) return false;
  int lo = 0, hi = p.S;  // last s with chunk_base[s] <= chunk
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.chunk_base[mid] <= chunk) lo = mid; else hi = mid;
  }
  s = lo;
  k = chunk - p.chunk_base[lo];
  return true;
}

struct K1Entry {
  int hd;      // layer * H + head, or -1 when the entry is not evictable
  int v;       // logical (virtual) token index inside the head
};

__device__ __forceinline__ K1Entry k1_decode(const K1Params& p, int s, const int* seq_sorted,
                                             int e, int ntok, int max_pos) {
  K1Entry r;
  r.hd = -1;
  r.v = 0;
  if (e >= ntok) return r;
  const int t = seq_sorted[e];
  const int blk = t / p.B;
  const int layer = p.layer_by_block[blk];
  const int head = p.head_by_block[blk];
  const int v = p.virtual_block_num_by_block[blk] * p.B + t % p.B;
  const int hd = layer * p.H + head;
  // reference :136-139: skip empty slots and the protected window
  if (v >= p.context_lens[s * p.L * p.H + hd] || p.kv_position[t] > max_pos) return r;
  if (k1_group_budget(p, s, layer) == 0) return r;  // :85-87 / control layers
  r.hd = hd;
  r.v = v;
  return r;
}

// (A) per-chunk histograms
__global__ void __launch_bounds__(kK1Threads) k1_hist_kernel(K1Params p) {
  extern __shared__ int sh_hist[];  // [L*H]
  int s, k;
  if (!k1_locate(p, blockIdx.x, s, k)) return;
  const int LH = p.L * p.H;
  for (int i = threadIdx.x; i < LH; i += kK1Threads) sh_hist[i] = 0;
  __syncthreads();
  int sb, ntok;
  k1_seq_range(p, s, sb, ntok);
  const int* seq_sorted = p.sorted_indices + (int64_t)sb * p.B;
  const int max_pos = p.last_position[s] - p.protected_window_size[s];  // :63
#pragma unroll
  for (int i = 0; i < kChunk / kK1Threads; ++i) {
    const int e = k * kChunk + i * kK1Threads + threadIdx.x;
    const K1Entry en = k1_decode(p, s, seq_sorted, e, ntok, max_pos);
    if (en.hd >= 0) atomicAdd(&sh_hist[en.hd], 1);
  }
  __syncthreads();
  int* out = p.hist + (int64_t)blockIdx.x * LH;
  for (int i = threadIdx.x; i < LH; i += kK1Threads) out[i] = sh_hist[i];
}

// (B) per-sequence scan over chunks + cut chunk per group + counts of uncut groups
__global__ void __launch_bounds__(kK1Threads) k1_scan_kernel(K1Params p) {
  extern __shared__ int sh[];  // kstar[NG]
  const int s = blockIdx.x;
  const int LH = p.L * p.H;
  const int NG = k1_num_groups(p);
  const int base = p.chunk_base[s];
  const int nch = p.chunk_base[s + 1] - base;
  int* sh_kstar = sh;
  for (int i = threadIdx.x; i < NG; i += kK1Threads) sh_kstar[i] = kNoCut;
  // step 1: exclusive scan along chunks, one thread per head (coalesced across heads)
  for (int hd = threadIdx.x; hd < LH; hd += kK1Threads) {
    int c = 0;
    constexpr int U = 8;  // batch the loads: hist/prefix may alias for the compiler
    for (int k0 = 0; k0 < nch; k0 += U) {
      int v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        v[u] = (k0 + u < nch) ? p.hist[(int64_t)(base + k0 + u) * LH + hd] : 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k0 + u < nch) p.prefix[(int64_t)(base + k0 + u) * LH + hd] = c;
        c += v[u];
      }
    }
  }
  __syncthreads();
  // step 2: first chunk whose inclusive completion count reaches the group's budget
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int heads_per_group = p.evenly ? p.H : LH;
  for (int k = wid; k < nch; k += kK1Threads / kWarp) {
    const int64_t row = (int64_t)(base + k) * LH;
    for (int grp = 0; grp < NG; ++grp) {
      const int budget = k1_group_budget(p, s, grp);
      if (budget == 0) continue;
      int f = 0;
      for (int j = lane; j < heads_per_group; j += kWarp) {
        const int hd = grp * heads_per_group + j;
        const int r0 = p.remaining_kv_count[s * LH + hd];
        f += k1_completions(p.prefix[row + hd] + p.hist[row + hd], r0, p.B);
      }
#pragma unroll
      for (int d = kWarp / 2; d >= 1; d >>= 1) f += __shfl_xor_sync(0xffffffffu, f, d);
      if (lane == 0 && f >= budget) atomicMin(&sh_kstar[grp], k);
    }
  }
  __syncthreads();
  // step 3: events still missing at the start of the cut chunk; publish kstar / erem
  for (int grp = wid; grp < NG; grp += kK1Threads / kWarp) {
    const int ks = sh_kstar[grp];
    int f = 0;
    if (ks != kNoCut) {
      const int64_t row = (int64_t)(base + ks) * LH;
      for (int j = lane; j < heads_per_group; j += kWarp) {
        const int hd = grp * heads_per_group + j;
        f += k1_completions(p.prefix[row + hd], p.remaining_kv_count[s * LH + hd], p.B);
      }
#pragma unroll
      for (int d = kWarp / 2; d >= 1; d >>= 1) f += __shfl_xor_sync(0xffffffffu, f, d);
    }
    if (lane == 0) {
      p.kstar[s * NG + grp] = ks;
      p.erem[s * NG + grp] = (ks != kNoCut) ? k1_group_budget(p, s, grp) - f : 0;
    }
  }
  // step 4: heads of groups that never reach their budget evict every valid slot
  // (also covers budget == 0 where nothing is valid, reference :80-87: counts are zeroed)
  for (int hd = threadIdx.x; hd < LH; hd += kK1Threads) {
    const int grp = p.evenly ? hd / p.H : 0;
    if (sh_kstar[grp] != kNoCut) continue;  // written by the cut chunk's CTA in (C)
    int c = 0;
    if (nch > 0) {
      const int64_t idx = (int64_t)(base + nch - 1) * LH + hd;
      c = p.prefix[idx] + p.hist[idx];
    }
    const int g = s * LH + hd;
    const int r0 = p.remaining_kv_count[g];
    p.evicted_kv_count[g] = c;
    p.remaining_kv_count[g] = k1_remaining(c, r0, p.B);
  }
}

// (C) stable scatter with the in-chunk cut
__global__ void __launch_bounds__(kK1Threads) k1_scatter_kernel(K1Params p) {
  extern __shared__ int sh[];
  int s, k;
  if (!k1_locate(p, blockIdx.x, s, k)) return;
  const int LH = p.L * p.H;
  const int NG = k1_num_groups(p);
  constexpr int NW = kK1Threads / kWarp;
  constexpr int STEPS = kChunk / kK1Threads;  // 4
  // smem: wcnt[NW][LH] (per-warp counts, then exclusive warp prefix) | cnt2[LH] | misc
  int* wcnt = sh;
  int* cnt2 = sh + NW * LH;
  __shared__ int s_warp_events[NW];
  __shared__ int s_cut_entry;
  __shared__ int s_cut_head;
  for (int i = threadIdx.x; i < (NW + 1) * LH; i += kK1Threads) sh[i] = 0;
  __syncthreads();

  int sb, ntok;
  k1_seq_range(p, s, sb, ntok);
  const int* seq_sorted = p.sorted_indices + (int64_t)sb * p.B;
  const int max_pos = p.last_position[s] - p.protected_window_size[s];
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int64_t row = (int64_t)blockIdx.x * LH;

  K1Entry ent[STEPS];
  int lrank[STEPS];
  bool any_cut_here = false;
#pragma unroll
  for (int i = 0; i < STEPS; ++i) {
    const int e = k * kChunk + wid * (kChunk / NW) + i * kWarp + lane;
    ent[i] = k1_decode(p, s, seq_sorted, e, ntok, max_pos);
    if (ent[i].hd >= 0) {
      const int ks = p.kstar[s * NG + (p.evenly ? ent[i].hd / p.H : 0)];
      if (k > ks) ent[i].hd = -1;          // the reference walk already returned
      else if (k == ks) any_cut_here = true;
    }
    const bool valid = ent[i].hd >= 0;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    lrank[i] = 0;
    if (valid) {
      const unsigned peers = __match_any_sync(vm, ent[i].hd);
      const int leader = __ffs(peers) - 1;
      int basev = 0;
      if (lane == leader) {
        basev = wcnt[wid * LH + ent[i].hd];
        wcnt[wid * LH + ent[i].hd] = basev + __popc(peers);
      }
      basev = __shfl_sync(peers, basev, leader);
      lrank[i] = basev + __popc(peers & lanemask_lt());
    }
    __syncwarp();
  }
  any_cut_here = __syncthreads_or(any_cut_here);
  // exclusive prefix over warps, per head
  for (int hd = threadIdx.x; hd < LH; hd += kK1Threads) {
    int run = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int t = wcnt[w * LH + hd];
      wcnt[w * LH + hd] = run;
      run += t;
    }
  }
  __syncthreads();
  int rank[STEPS];  // global 0-based rank inside the head's segment
#pragma unroll
  for (int i = 0; i < STEPS; ++i)
    rank[i] = (ent[i].hd >= 0) ? p.prefix[row + ent[i].hd] + wcnt[wid * LH + ent[i].hd] + lrank[i] : 0;

  if (any_cut_here) {
    for (int grp = 0; grp < NG; ++grp) {
      if (p.kstar[s * NG + grp] != k) continue;  // uniform across the CTA
      const int need = p.erem[s * NG + grp];
      // ordered event scan: entry order is (warp, step, lane)
      unsigned evmask[STEPS];
      int warp_events = 0;
#pragma unroll
      for (int i = 0; i < STEPS; ++i) {
        bool ev = false;
        if (ent[i].hd >= 0 && (p.evenly ? ent[i].hd / p.H : 0) == grp) {
          const int r0 = p.remaining_kv_count[s * LH + ent[i].hd];
          const int c = rank[i] + 1;
          ev = (r0 >= 1 && c >= r0 && (c - r0) % p.B == 0);
        }
        evmask[i] = __ballot_sync(0xffffffffu, ev);
        warp_events += __popc(evmask[i]);
      }
      if (threadIdx.x == 0) { s_cut_entry = kNoCut; s_cut_head = -1; }
      if (lane == 0) s_warp_events[wid] = warp_events;
      __syncthreads();
      int before = 0;
      for (int w = 0; w < wid; ++w) before += s_warp_events[w];
#pragma unroll
      for (int i = 0; i < STEPS; ++i) {
        const bool ev = (evmask[i] >> lane) & 1u;
        const int incl = before + __popc(evmask[i] & (lanemask_lt() | (1u << lane)));
        if (ev && incl == need) {
          s_cut_entry = wid * (kChunk / NW) + i * kWarp + lane;
          s_cut_head = ent[i].hd;
        }
        before += __popc(evmask[i]);
      }
      __syncthreads();
      const int cut = s_cut_entry;
      const int cut_head = s_cut_head;
#pragma unroll
      for (int i = 0; i < STEPS; ++i) {
        if (ent[i].hd >= 0 && (p.evenly ? ent[i].hd / p.H : 0) == grp) {
          const int epos = wid * (kChunk / NW) + i * kWarp + lane;
          if (epos > cut) ent[i].hd = -1;               // past the reference's return
          else atomicAdd(&cnt2[ent[i].hd], 1);
        }
      }
      __syncthreads();
      const int hpg = p.evenly ? p.H : LH;
      for (int j = threadIdx.x; j < hpg; j += kK1Threads) {
        const int hd = grp * hpg + j;
        const int g = s * LH + hd;
        const int c = p.prefix[row + hd] + cnt2[hd];
        const int r0 = p.remaining_kv_count[g];
        p.evicted_kv_count[g] = c;
        // the completing head returns before its counter is reset (reference :149-153)
        p.remaining_kv_count[g] = (hd == cut_head) ? 0 : k1_remaining(c, r0, p.B);
      }
      __syncthreads();
    }
  }
  // scatter
  const int gbase = s * LH;
#pragma unroll
  for (int i = 0; i < STEPS; ++i) {
    if (ent[i].hd >= 0) {
      const int g = gbase + ent[i].hd;
      const int pos = p.evicted_kv_offsets[g] + rank[i];
      p.evicted_kv_indices[pos] = g;              // :144
      p.evicted_logical_indices[pos] = ent[i].v;  // :145
    }
  }
}

}  // namespace kvc

// =============================================================================
// C ABI
// =============================================================================
using namespace kvc;

static bool supported_sched_block_size(int B) { return B == 1 || B == 2 || B == 4 || B == 16; }

extern "C" int kvc_truncate_cache_evictions(
    int* evicted_kv_indices, int* evicted_logical_indices, int* evicted_kv_count,
    const int* evicted_kv_offsets, const int* hanging_token_count, int num_seqs,
    int num_layers, int num_kv_heads, int block_size, int max_evicted_kv,
    int null_eviction_index, kvc_stream_t stream) {
  KVC_REQUIRE(supported_sched_block_size(block_size), KVC_ERR_UNSUPPORTED,
              "Unsupported block size: %d", block_size);
  if (num_seqs == 0) return KVC_OK;  // empty batch: nothing to do (tensors may be empty / null)
  KVC_REQUIRE(evicted_kv_indices && evicted_logical_indices && evicted_kv_count &&
                  evicted_kv_offsets && hanging_token_count,
              KVC_ERR_INVALID, "truncate_cache_evictions: null tensor");
  KVC_REQUIRE(num_seqs >= 0 && num_layers > 0 && num_kv_heads > 0 && max_evicted_kv >= 0,
              KVC_ERR_INVALID, "truncate_cache_evictions: bad sizes");
  return launch_truncate(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                         evicted_kv_offsets, hanging_token_count,
                         num_seqs * num_layers * num_kv_heads, block_size, max_evicted_kv,
                         null_eviction_index, as_stream(stream));
}

extern "C" int kvc_count_block_evictions(
    int* evicted_block_count, int* evicted_logical_indices, const int* evicted_kv_offsets,
    const int* hanging_token_count, int num_seqs, int num_layers, int num_kv_heads,
    int total_kvs, int block_size, int null_value, kvc_stream_t stream) {
  KVC_REQUIRE(supported_sched_block_size(block_size), KVC_ERR_UNSUPPORTED,
              "Unsupported block size: %d", block_size);
  if (num_seqs == 0) return KVC_OK;
  KVC_REQUIRE(evicted_block_count && evicted_logical_indices && evicted_kv_offsets &&
                  hanging_token_count,
              KVC_ERR_INVALID, "count_block_evictions: null tensor");
  const int total_heads = num_seqs * num_layers * num_kv_heads;
  if (total_heads <= 0) return KVC_OK;
  count_block_evictions_kernel<<<ceil_div(total_heads, 8), 256, 0, as_stream(stream)>>>(
      evicted_block_count, evicted_logical_indices, evicted_kv_offsets, hanging_token_count,
      total_heads, total_kvs, block_size, null_value);
  KVC_CHECK_LAUNCH("count_block_evictions_kernel");
  return KVC_OK;
}

extern "C" int kvc_schedule_t1_cache_moves(
    int* cache_moves_idx, int* cache_moves_count, const int* evicted_logical_indices,
    const int* evicted_kv_count, const int* evicted_kv_offsets, const int* block_tables,
    const int* context_lens, int num_seqs, int num_layers, int num_kv_heads,
    int max_num_blocks_per_seq, int block_size, kvc_stream_t stream) {
  KVC_REQUIRE(supported_sched_block_size(block_size), KVC_ERR_UNSUPPORTED,
              "Unsupported block size: %d", block_size);
  if (num_seqs == 0) return KVC_OK;
  KVC_REQUIRE(cache_moves_idx && cache_moves_count && evicted_logical_indices &&
                  evicted_kv_count && evicted_kv_offsets && block_tables && context_lens,
              KVC_ERR_INVALID, "schedule_t1_cache_moves: null tensor");
  KVC_REQUIRE((reinterpret_cast<uintptr_t>(cache_moves_idx) & 7) == 0, KVC_ERR_INVALID,
              "schedule_t1_cache_moves: cache_moves_idx must be 8-byte aligned");
  const int total = num_seqs * num_layers * num_kv_heads;
  if (total <= 0) return KVC_OK;
  const bool many = total >= 4096;  // decode-time batches: 32 heads per warp
  const dim3 grid(many ? ceil_div(ceil_div(total, 32), 4) : ceil_div(total, 4)), block(128);
  // few heads whose lists can be long (post-prefill compression): one CTA per head
  const bool long_lists = !many && (int64_t)max_num_blocks_per_seq * block_size > 512;
  cudaStream_t st = as_stream(stream);
#define KVC_LAUNCH_K4(BS)                                                              \
  if (long_lists)                                                                      \
    schedule_moves_cta_kernel<BS><<<total, 256, 0, st>>>(                              \
        cache_moves_idx, cache_moves_count, evicted_logical_indices, evicted_kv_count, \
        evicted_kv_offsets, block_tables, context_lens, num_seqs, num_layers,          \
        num_kv_heads, max_num_blocks_per_seq);                                         \
  else if (many)                                                                       \
    schedule_moves_many_kernel<BS><<<grid, block, 0, st>>>(                            \
        cache_moves_idx, cache_moves_count, evicted_logical_indices, evicted_kv_count, \
        evicted_kv_offsets, block_tables, context_lens, num_seqs, num_layers,          \
        num_kv_heads, max_num_blocks_per_seq);                                         \
  else                                                                                 \
    schedule_moves_kernel<BS><<<grid, block, 0, st>>>(                                 \
        cache_moves_idx, cache_moves_count, evicted_logical_indices, evicted_kv_count, \
        evicted_kv_offsets, block_tables, context_lens, num_seqs, num_layers,          \
        num_kv_heads, max_num_blocks_per_seq)
  switch (block_size) {
    case 1: KVC_LAUNCH_K4(1); break;
    case 2: KVC_LAUNCH_K4(2); break;
    case 4: KVC_LAUNCH_K4(4); break;
    default: KVC_LAUNCH_K4(16); break;
  }
#undef KVC_LAUNCH_K4
  KVC_CHECK_LAUNCH("schedule_moves_kernel");
  return KVC_OK;
}

static int k1_max_chunks(int num_seqs, int total_blocks, int block_size) {
  const long long n = (long long)total_blocks * block_size;
  return (int)((n + kChunk - 1) / kChunk) + num_seqs;
}

extern "C" size_t kvc_schedule_cache_evictions_workspace_bytes(
    int num_seqs, int num_layers, int num_kv_heads, int total_blocks, int block_size) {
  if (num_seqs <= 0 || num_layers <= 0 || num_kv_heads <= 0 || total_blocks < 0 || block_size <= 0)
    return 0;
  const size_t LH = (size_t)num_layers * num_kv_heads;
  const size_t mc = (size_t)k1_max_chunks(num_seqs, total_blocks, block_size);
  size_t ints = (size_t)(num_seqs + 1)        // chunk_base
                + 2 * mc * LH                  // hist, prefix
                + 2 * (size_t)num_seqs * num_layers;  // kstar, erem (NG <= L)
  return ints * sizeof(int) + 256;
}

extern "C" int kvc_schedule_cache_evictions(
    int* evicted_kv_indices, int* evicted_logical_indices, int* evicted_kv_count,
    int* remaining_kv_count, const int* evicted_kv_offsets, const int* sorted_indices,
    const int* seq_block_offsets, const int* layer_by_block, const int* head_by_block,
    const int* virtual_block_num_by_block, const int* evicted_blocks_per_seq,
    const int* context_lens, const int* hanging_token_count, const int* kv_position,
    const int* last_position, const int* protected_window_size, int num_seqs, int num_layers,
    int num_kv_heads, int total_blocks, int block_size, int evict_evenly_per_layer,
    const int* control_layers, int num_control_layers, int max_evicted_kv,
    int null_eviction_index, int truncate, void* workspace, size_t workspace_bytes,
    kvc_stream_t stream) {
  KVC_REQUIRE(supported_sched_block_size(block_size), KVC_ERR_UNSUPPORTED,
              "Unsupported block size: %d", block_size);
  if (num_seqs == 0) return KVC_OK;
  KVC_REQUIRE(evicted_kv_indices && evicted_logical_indices && evicted_kv_count &&
                  remaining_kv_count && evicted_kv_offsets && sorted_indices && seq_block_offsets &&
                  layer_by_block && head_by_block && virtual_block_num_by_block &&
                  evicted_blocks_per_seq && context_lens && hanging_token_count && kv_position &&
                  last_position && protected_window_size,
              KVC_ERR_INVALID, "schedule_cache_evictions: null tensor");
  KVC_REQUIRE(num_seqs >= 0 && num_layers > 0 && num_kv_heads > 0 && total_blocks >= 0,
              KVC_ERR_INVALID, "schedule_cache_evictions: bad sizes");
  KVC_REQUIRE(num_control_layers == 0 || control_layers, KVC_ERR_INVALID,
              "schedule_cache_evictions: control_layers is null");
  if (num_seqs == 0) return KVC_OK;
  const int LH = num_layers * num_kv_heads;
  KVC_REQUIRE(LH <= 4096, KVC_ERR_UNSUPPORTED,
              "schedule_cache_evictions: num_layers*num_kv_heads = %d > 4096", LH);
  const size_t need = kvc_schedule_cache_evictions_workspace_bytes(num_seqs, num_layers,
                                                                   num_kv_heads, total_blocks,
                                                                   block_size);
  KVC_REQUIRE(workspace && workspace_bytes >= need, KVC_ERR_WORKSPACE,
              "schedule_cache_evictions: workspace of %zu bytes required, got %zu", need,
              workspace_bytes);
  cudaStream_t st = as_stream(stream);
  K1Params p;
  p.evicted_kv_indices = evicted_kv_indices;
  p.evicted_logical_indices = evicted_logical_indices;
  p.evicted_kv_count = evicted_kv_count;
  p.remaining_kv_count = remaining_kv_count;
  p.evicted_kv_offsets = evicted_kv_offsets;
  p.sorted_indices = sorted_indices;
  p.seq_block_offsets = seq_block_offsets;
  p.layer_by_block = layer_by_block;
  p.head_by_block = head_by_block;
  p.virtual_block_num_by_block = virtual_block_num_by_block;
  p.evicted_blocks_per_seq = evicted_blocks_per_seq;
  p.context_lens = context_lens;
  p.kv_position = kv_position;
  p.last_position = last_position;
  p.protected_window_size = protected_window_size;
  p.control_layers = control_layers;
  p.num_control_layers = evict_evenly_per_layer ? num_control_layers : 0;
  p.S = num_seqs; p.L = num_layers; p.H = num_kv_heads;
  p.total_blocks = total_blocks; p.B = block_size;
  p.evenly = evict_evenly_per_layer ? 1 : 0;
  p.max_chunks = k1_max_chunks(num_seqs, total_blocks, block_size);
  int* w = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  p.chunk_base = w; w += num_seqs + 1;
  p.hist = w;       w += (size_t)p.max_chunks * LH;
  p.prefix = w;     w += (size_t)p.max_chunks * LH;
  p.kstar = w;      w += (size_t)num_seqs * num_layers;
  p.erem = w;

  k1_setup_kernel<<<1, 32, 0, st>>>(p);
  KVC_CHECK_LAUNCH("k1_setup_kernel");
  k1_hist_kernel<<<p.max_chunks, kK1Threads, LH * sizeof(int), st>>>(p);
  KVC_CHECK_LAUNCH("k1_hist_kernel");
  k1_scan_kernel<<<num_seqs, kK1Threads, (size_t)num_layers * sizeof(int), st>>>(p);
  KVC_CHECK_LAUNCH("k1_scan_kernel");
  const size_t scatter_smem = (size_t)(kK1Threads / kWarp + 1) * LH * sizeof(int);
  if (scatter_smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k1_scatter_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)scatter_smem);
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "k1_scatter smem attr: %s", cudaGetErrorString(e));
  }
  k1_scatter_kernel<<<p.max_chunks, kK1Threads, scatter_smem, st>>>(p);
  KVC_CHECK_LAUNCH("k1_scatter_kernel");
  if (truncate) {  // launcher macro :463-472
    return launch_truncate(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                           evicted_kv_offsets, hanging_token_count, num_seqs * LH, block_size,
                           max_evicted_kv, null_eviction_index, st);
  }
  return KVC_OK;
}
