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
//   1. rank:   one warp (ctx <= 256) or one CTA per head; keys staged in shared memory, rank by
//              counting (n^2/threads compares, n ~ 144 in steady state); completion keys and
//              per-slot ranks to workspace.
//   2. select: one CTA per sequence; 4-pass radix select of the E-th smallest completion key,
//              then per-head lower/upper bounds + an exclusive scan for deterministic ties.
//   3. emit:   same grouping as 1; ordered compaction of slots with rank < count.
// All of it is int32/fp32 gather traffic: 8 B read per slot in (1), 4 B written + read for the
// rank, 4 B per evicted index; HBM/L2-bound, no tensor-core shapes.
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
  int* ranks;        // [total slots] workspace
  unsigned* events;  // [total slots / B] workspace
};

// group = GT threads cooperating on one head (32 = a warp, 128 = the whole CTA)
template <int GT>
__device__ __forceinline__ void group_sync() {
  if (GT == kWarp) __syncwarp(); else __syncthreads();
}

template <int GT>
__global__ void __launch_bounds__(128) select_rank_kernel(SelectParams p, int max_slots) {
  extern __shared__ unsigned sk[];  // [groups][max_slots]
  constexpr int GROUPS = 128 / GT;
  const int grp = threadIdx.x / GT, t = threadIdx.x % GT;
  const int g = blockIdx.x * GROUPS + grp;  // (s,l,h) flat
  const int total = p.S * p.L * p.H;
  const bool active = g < total;
  unsigned* keys = sk + grp * max_slots;
  __shared__ int s_ncand[GROUPS];
  if (t == 0) s_ncand[grp] = 0;
  group_sync<GT>();
  int n = 0, off = 0, r0 = 0, nblk = 0;
  if (active) {
    const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
    const int lsh = (l * p.S + s) * p.H + h;
    n = p.context_lens[lsh];
    nblk = (n + p.B - 1) / p.B;
    off = p.offsets[g];
    r0 = p.hanging[g];
    const int max_pos = p.last_position[s] - p.protected_window[s];
    const int* __restrict__ bt = p.block_tables + (int64_t)lsh * p.MB;
    int local = 0;
    for (int i = t; i < n; i += GT) {
      const int64_t slot = (int64_t)bt[i / p.B] * p.B + i % p.B;
      const bool cand = p.kv_position[slot] <= max_pos;  // K1 :136-139 (i < ctx by construction)
      unsigned k = orderable(p.kv_metrics[slot]);
      if (!cand || k == kInfKey) k = kInfKey;
      else ++local;
      keys[i] = k;
    }
    if (local) atomicAdd(&s_ncand[grp], local);
  }
  group_sync<GT>();
  if (!active) return;
  const int ncand = s_ncand[grp];
  const int nev = (r0 >= 1 && ncand >= r0) ? (ncand - r0) / p.B + 1 : 0;
  unsigned* ev = p.events + off / p.B;
  for (int i = t; i < n; i += GT) {
    const unsigned ki = keys[i];
    int rank = -1;
    if (ki != kInfKey) {
      int r = 0;
      for (int j = 0; j < n; ++j) {
        const unsigned kj = keys[j];
        r += (kj < ki) || (kj == ki && j < i);
      }
      rank = r;
      if (r0 >= 1 && r >= r0 - 1 && (r - (r0 - 1)) % p.B == 0) ev[(r - (r0 - 1)) / p.B] = ki;
    }
    p.ranks[off + i] = rank;
  }
  for (int i = n + t; i < nblk * p.B; i += GT) p.ranks[off + i] = -1;
  for (int k = nev + t; k < nblk; k += GT) ev[k] = kInfKey;
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

template <int GT>
__global__ void __launch_bounds__(128) select_emit_kernel(SelectParams p) {
  constexpr int GROUPS = 128 / GT;
  constexpr int GW = GT / kWarp;  // warps per group
  __shared__ int s_wcnt[GROUPS][GW > 0 ? GW : 1];
  const int grp = threadIdx.x / GT, t = threadIdx.x % GT;
  const int lane = threadIdx.x % kWarp, gw = t / kWarp;
  const int g = blockIdx.x * GROUPS + grp;
  const int total = p.S * p.L * p.H;
  const bool active = g < total;
  int n = 0, off = 0, cnt = 0, nblk = 0;
  if (active) {
    const int s = g / (p.L * p.H), l = (g / p.H) % p.L, h = g % p.H;
    n = p.context_lens[(l * p.S + s) * p.H + h];
    nblk = (n + p.B - 1) / p.B;
    off = p.offsets[g];
    cnt = p.evicted_kv_count[g];
  }
  int* __restrict__ out = p.evicted_logical_indices + off;
  int written = 0;
  const int iters = (n + GT - 1) / GT;
  // all threads of a CTA-wide group must take the same trip count (n is group-uniform)
  for (int it = 0; it < iters; ++it) {
    const int i = it * GT + t;
    bool ev = false;
    if (active && i < n) {
      const int r = p.ranks[off + i];
      ev = (r >= 0 && r < cnt);
    }
    const unsigned m = __ballot_sync(0xffffffffu, ev);
    int base = 0;
    if (GW > 1) {
      if (lane == 0) s_wcnt[grp][gw] = __popc(m);
      __syncthreads();
      int tot = 0;
#pragma unroll
      for (int w = 0; w < GW; ++w) {
        if (w < gw) base += s_wcnt[grp][w];
        tot += s_wcnt[grp][w];
      }
      if (ev) out[written + base + __popc(m & lanemask_lt())] = i;
      written += tot;
      __syncthreads();
    } else {
      if (ev) out[written + __popc(m & lanemask_lt())] = i;
      written += __popc(m);
    }
  }
  if (active) {
    for (int i = cnt + t; i < nblk * p.B; i += GT) out[i] = p.null_index;
  }
}

}  // namespace kvc

using namespace kvc;

extern "C" size_t kvc_select_evictions_workspace_bytes(int64_t max_evicted_kv, int block_size) {
  if (max_evicted_kv <= 0 || block_size <= 0) return 0;
  return (size_t)max_evicted_kv * sizeof(int) + ((size_t)max_evicted_kv / block_size + 1) * sizeof(unsigned) + 512;
}

extern "C" int kvc_select_evictions(
    int* evicted_logical_indices, int* evicted_kv_count, int* evicted_block_count,
    const float* kv_metrics, const int* kv_position, const int* block_tables, const int* context_lens,
    const int* hanging_token_count, const int* evicted_kv_offsets, const int* evicted_blocks_per_seq,
    const int* last_position, const int* protected_window_size, int num_seqs, int num_layers,
    int num_kv_heads, int max_num_blocks_per_seq, int block_size, int64_t max_evicted_kv,
    int null_eviction_index, void* workspace, size_t workspace_bytes, kvc_stream_t stream) {
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
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  p.ranks = reinterpret_cast<int*>(w);
  p.events = reinterpret_cast<unsigned*>(w + (size_t)max_evicted_kv * sizeof(int));
  cudaStream_t st = as_stream(stream);
  const int total = num_seqs * num_layers * num_kv_heads;
  const int max_slots = max_num_blocks_per_seq * block_size;
  if (max_slots <= 256) {  // steady-state compressed contexts: one warp per head
    const size_t smem = (size_t)4 * max_slots * sizeof(unsigned);
    select_rank_kernel<32><<<ceil_div(total, 4), 128, smem, st>>>(p, max_slots);
  } else {
    const size_t smem = (size_t)max_slots * sizeof(unsigned);
    KVC_REQUIRE(smem <= 220 * 1024, KVC_ERR_UNSUPPORTED,
                "select_evictions: %d slots per head exceed the shared-memory staging limit", max_slots);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(select_rank_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "select_rank smem attr: %s", cudaGetErrorString(e));
    }
    select_rank_kernel<128><<<total, 128, smem, st>>>(p, max_slots);
  }
  KVC_CHECK_LAUNCH("select_rank_kernel");
  select_threshold_kernel<<<num_seqs, 256, 0, st>>>(p);
  KVC_CHECK_LAUNCH("select_threshold_kernel");
  if (max_slots <= 256) select_emit_kernel<32><<<ceil_div(total, 4), 128, 0, st>>>(p);
  else select_emit_kernel<128><<<total, 128, 0, st>>>(p);
  KVC_CHECK_LAUNCH("select_emit_kernel");
  return KVC_OK;
}
