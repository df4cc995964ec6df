// Device-side block allocator + block-state updates (no host round trip, no full-table passes).
//
// replaces ParallelBlockAllocator (vllm/kvcompress/block_manager.py:75-140) and the block-manager
// paths built on it: _add_sequence (:196-222), _append_to_sequence_batch (:269-294 + block.py:275-321,
// 505-560 metadata), free_compressed_blocks (:466-530 + block.py:184-210) and _remove_sequence(_batch)
// (:224-262).  The allocator state is the reference's own: free_mask [NB] bool (1 = free).  Results are
// bit-identical to the reference's `block_numbers[free_mask][:n]` (the n lowest-numbered free blocks,
// ascending) assigned in row-major order of the boolean masks it builds ([L, batch, H, MB] for decode,
// [L, H, nblk] per prompt), so block tables come out identical on the same call sequence.
//
// The reference compacts the whole free mask per allocation (NB int64 gathers) and materialises
// [L,S,H,MB] boolean masks; here the k-th free block is found through a two-level popcount index:
//   1. one coalesced pass over the mask (NB bytes): free blocks per 256-block chunk + a packed bitmap
//      snapshot of the mask (NB/8 bytes; the later passes select from the snapshot, so clearing mask
//      bytes concurrently is harmless), + per-CTA totals of the heads that need a block
//   2. per 256-chunk super-chunk: exclusive scan of the chunk counts + super totals
//   3. per requesting head: rank (scan of needs) -> super-chunk (binary search in smem) -> chunk
//      (binary search) -> n-th set bit of the chunk's 8 bitmap words; table entry, metadata, slot mapping,
//      token position and context length are written in the same thread.
// Out of blocks: nothing is allocated and *status is set to 1 (the reference raises ValueError on the
// host, block_manager.py:108-110).  A head whose block table is full (context length == MB * B and one more
// token to append: the reference would raise an index error on block_tables[..., MB]): nothing is allocated or
// appended by that call or any later one and *status is set to 2 (sticky until the host resets it).
#include "kvc_common.cuh"

namespace kvc {

constexpr int kChunk = 256;        // blocks per chunk (= bytes of the bool mask)
constexpr int kSuper = 256;        // chunks per super-chunk
constexpr int kNeedCta = 256;      // heads per CTA in the rank scan

struct AllocIndex {
  const uint8_t* free_mask;
  uint32_t* bitmap;    // [n_chunks * 8] packed snapshot of free_mask
  int64_t num_blocks;
  int* chunk_prefix;   // [n_chunks] counts, then within-super exclusive prefix
  int* super_total;    // [n_super]
  int n_chunks, n_super;
};

__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- 1. chunk counts (+ optional need partials, computed by the trailing CTAs) ----
__device__ __forceinline__ uint32_t pack4(uint32_t w) {  // 4 bool bytes (0/1) -> 4 bits
  return (w & 1u) | ((w >> 7) & 2u) | ((w >> 14) & 4u) | ((w >> 21) & 8u);
}

// pieces (512 B of mask = 2 chunks) per warp: 2 keeps the pass short -- it is a chain of dependent shuffles per
// piece, and 96 warps of 16 pieces each took 12 us for a 786 KB mask
constexpr int kCountPieces = 2;
constexpr int kCountChunksPerWarp = 2 * kCountPieces;

__device__ void chunk_count_warp(const AllocIndex& ix, int warp_global, int lane) {
  // piece p: lane reads 16 B at offset (p*32 + lane)*16
  const int64_t base = (int64_t)warp_global * kCountChunksPerWarp * kChunk;
  const int64_t end = (int64_t)ix.n_chunks * kChunk;  // bitmap covers whole chunks; bits >= num_blocks are 0
  if (base >= end) return;
  uint16_t* bm16 = reinterpret_cast<uint16_t*>(ix.bitmap);
#pragma unroll
  for (int p = 0; p < kCountPieces; ++p) {
    const int64_t off = base + ((int64_t)p * 32 + lane) * 16;
    uint32_t bits = 0;
    if (off + 16 <= ix.num_blocks) {
      const uint4 v = *reinterpret_cast<const uint4*>(ix.free_mask + off);
      bits = pack4(v.x) | (pack4(v.y) << 4) | (pack4(v.z) << 8) | (pack4(v.w) << 12);
    } else {
      for (int64_t b = off; b < ix.num_blocks; ++b) bits |= (uint32_t)(ix.free_mask[b] != 0) << (int)(b - off);
    }
    if (off < end) bm16[off >> 4] = (uint16_t)bits;
    int c = __popc(bits);
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    const int64_t chunk = (base >> 8) + 2 * p + (lane >> 4);
    if ((lane & 15) == 0 && chunk < ix.n_chunks) ix.chunk_prefix[chunk] = c;
  }
}

struct AppendParams {
  AllocIndex ix;
  uint8_t* free_mask;
  int* free_count;
  int* status;
  int* block_tables;    // [L, max_seqs, H, MB]
  int* context_lens;    // [L, max_seqs, H]
  const int* seq_slots; // [num_batch] or null (= 0..num_batch-1)
  const int* seq_len;   // [num_batch] position of the appended token
  int64_t* slot_mapping;  // [L, num_batch, H]
  int* kv_position;     // [NB, B]
  int* seq_index_by_block;   // optional metadata arrays [NB]
  int* layer_index_by_block;
  int* head_index_by_block;
  int* logical_block_num_by_block;
  int* need_partials;   // [n_need_ctas]
  int num_batch, max_seqs, L, H, MB, B, n_need_ctas;
};

__device__ __forceinline__ int append_need(const AppendParams& p, int g, int& lsh, int& ctx) {
  // g enumerates (l, batch position, h) row-major: the order of the reference's new_mask (block_manager.py:280-288)
  const int h = g % p.H, sb = (g / p.H) % p.num_batch, l = g / (p.H * p.num_batch);
  const int slot = p.seq_slots ? p.seq_slots[sb] : sb;
  lsh = (l * p.max_seqs + slot) * p.H + h;
  ctx = p.context_lens[lsh];
  return (ctx % p.B) == 0;  // (ctx + 1) % B == 1 (block_manager.py:261-267), true for every token when B == 1
}

__global__ void __launch_bounds__(256) alloc_count_kernel(AppendParams p, int count_ctas) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x < count_ctas) {
    chunk_count_warp(p.ix, blockIdx.x * 8 + warp, lane);
    return;
  }
  // need partials: this CTA covers kNeedCta heads, one per thread
  __shared__ int s_tot;
  if (threadIdx.x == 0) s_tot = 0;
  __syncthreads();
  const int cta = blockIdx.x - count_ctas;
  const int total = p.L * p.num_batch * p.H;
  const int g = cta * kNeedCta + threadIdx.x;
  int lsh, ctx;
  int local = (g < total) ? append_need(p, g, lsh, ctx) : 0;
  if (local && ctx / p.B >= p.MB) *p.status = 2;  // table overflow: alloc_append_kernel bails out
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if (lane == 0 && local) atomicAdd(&s_tot, local);
  __syncthreads();
  if (threadIdx.x == 0) p.need_partials[cta] = s_tot;
}

// chunk counts only (prefill / no scan of needs)
__global__ void __launch_bounds__(256) alloc_count_only_kernel(AllocIndex ix) {
  chunk_count_warp(ix, blockIdx.x * 8 + (threadIdx.x >> 5), threadIdx.x & 31);
}

// ---- 2. per super-chunk: exclusive scan of its 256 chunk counts ----
__global__ void __launch_bounds__(kSuper) alloc_scan_kernel(AllocIndex ix) {
  __shared__ int s_w[kSuper / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int chunk = blockIdx.x * kSuper + t;
  const int v = (chunk < ix.n_chunks) ? ix.chunk_prefix[chunk] : 0;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < kSuper / 32; ++q) {
    if (q < w) woff += s_w[q];
    tot += s_w[q];
  }
  if (chunk < ix.n_chunks) ix.chunk_prefix[chunk] = woff + incl - v;
  if (t == 0) ix.super_total[blockIdx.x] = tot;
}

// ---- 3. rank -> block id ----
// s_super: exclusive prefix of the super totals in shared memory (n_super + 1 entries)
__device__ int find_free_block(const AllocIndex& ix, const int* s_super, int rank) {
  int lo = 0, hi = ix.n_super;  // last super with prefix <= rank
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_super[mid] <= rank) lo = mid; else hi = mid;
  }
  int r = rank - s_super[lo];
  const int c0 = lo * kSuper, c1 = min(ix.n_chunks, c0 + kSuper);
  int a = c0, b = c1;  // last chunk in [c0, c1) with prefix <= r
  while (b - a > 1) {
    const int mid = (a + b) >> 1;
    if (ix.chunk_prefix[mid] <= r) a = mid; else b = mid;
  }
  r -= ix.chunk_prefix[a];
  const uint4* w4 = reinterpret_cast<const uint4*>(ix.bitmap + (int64_t)a * (kChunk / 32));
  const uint4 lo4 = w4[0], hi4 = w4[1];
  const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = __popc(w[j]);
    if (r < c) return (int)((int64_t)a * kChunk + j * 32 + __fns(w[j], 0, r + 1));
    r -= c;
  }
  return -1;  // unreachable when rank < total free
}

// loads the super totals into shared memory as an exclusive prefix; returns total free blocks
__device__ int load_super_prefix(const AllocIndex& ix, int* s_super) {
  // n_super is small (NB / 65536): serial scan by one thread
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < ix.n_super; ++i) {
      s_super[i] = acc;
      acc += ix.super_total[i];
    }
    s_super[ix.n_super] = acc;
  }
  __syncthreads();
  return s_super[ix.n_super];
}

__global__ void __launch_bounds__(kNeedCta) alloc_append_kernel(AppendParams p) {
  extern __shared__ int s_super[];  // [n_super + 1]
  __shared__ int s_w[kNeedCta / 32];
  __shared__ int s_base, s_total_need;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (*p.status == 2) return;  // a block table is full (set by alloc_count_kernel): change nothing
  const int total_free = load_super_prefix(p.ix, s_super);
  if (t == 0) s_base = s_total_need = 0;
  __syncthreads();
  {
    int before = 0, all = 0;  // needs of the CTAs before this one / of all CTAs
    for (int i = t; i < p.n_need_ctas; i += kNeedCta) {
      const int v = p.need_partials[i];
      all += v;
      if (i < (int)blockIdx.x) before += v;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, d);
      all += __shfl_xor_sync(0xffffffffu, all, d);
    }
    if (lane == 0 && all) {
      atomicAdd(&s_total_need, all);
      if (before) atomicAdd(&s_base, before);
    }
  }
  __syncthreads();
  const int total_need = s_total_need;
  if (total_need > total_free) {  // out of blocks: allocate nothing, flag it
    if (blockIdx.x == 0 && t == 0) *p.status = 1;
    return;
  }
  if (blockIdx.x == 0 && t == 0) *p.free_count = total_free - total_need;
  const int total = p.L * p.num_batch * p.H;
  const int g = blockIdx.x * kNeedCta + t;
  int lsh = 0, ctx = 0, need = 0;
  if (g < total) need = append_need(p, g, lsh, ctx);
  int incl = need;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  int woff = 0;
  for (int q = 0; q < w; ++q) woff += s_w[q];
  if (g >= total) return;
  const int h = g % p.H, sb = (g / p.H) % p.num_batch, l = g / (p.H * p.num_batch);
  const int pos = p.seq_len[sb];
  int* bt = p.block_tables + (int64_t)lsh * p.MB;
  int blk;
  if (need) {
    const int rank = s_base + woff + incl - 1;
    blk = find_free_block(p.ix, s_super, rank);
    bt[ctx / p.B] = blk;
    p.free_mask[blk] = 0;
    // metadata of the new block (block.py:505-560, metrics.py:344-362): ids + the whole row of token
    // positions pre-filled with pos, pos+1, ...
    const int slot_idx = p.seq_slots ? p.seq_slots[sb] : sb;
    if (p.seq_index_by_block) p.seq_index_by_block[blk] = slot_idx;
    if (p.layer_index_by_block) p.layer_index_by_block[blk] = l;
    if (p.head_index_by_block) p.head_index_by_block[blk] = h;
    if (p.logical_block_num_by_block) p.logical_block_num_by_block[blk] = ctx / p.B;
    if (p.B == 16) {  // one 64 B row: four 16 B stores
      int4* row = reinterpret_cast<int4*>(p.kv_position + (int64_t)blk * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) row[j] = make_int4(pos + 4 * j, pos + 4 * j + 1, pos + 4 * j + 2, pos + 4 * j + 3);
    } else {
      for (int j = 0; j < p.B; ++j) p.kv_position[(int64_t)blk * p.B + j] = pos + j;
    }
  } else {
    blk = bt[ctx / p.B];
    p.kv_position[(int64_t)blk * p.B + ctx % p.B] = pos;
  }
  p.slot_mapping[g] = (int64_t)blk * p.B + ctx % p.B;  // block.py:303-321 (decode slot mapping)
  p.context_lens[lsh] = ctx + 1;
}

// ---- prefill: add sequences (block_manager.py:196-222) ----
struct AddParams {
  AllocIndex ix;
  uint8_t* free_mask;
  int* free_count;
  int* status;
  int* block_tables;
  int* context_lens;
  const int* seq_slots;   // [n]
  const int* seq_lens;    // [n]
  int* kv_position;
  int* seq_index_by_block;
  int* layer_index_by_block;
  int* head_index_by_block;
  int* logical_block_num_by_block;
  int n, max_seqs, L, H, MB, B;
};

// grid (ceil(L*H*max_nblk / 256), n): thread = one requested block (l, h, b) of sequence i
__global__ void __launch_bounds__(256) alloc_add_kernel(AddParams p, int max_nblk) {
  extern __shared__ int s_super[];
  const int total_free = load_super_prefix(p.ix, s_super);
  const int i = blockIdx.y;
  // blocks requested by the sequences before this one, and in total (n is small: a prefill batch)
  int base = 0, total_need = 0;
  for (int j = 0; j < p.n; ++j) {
    const int nb = (p.seq_lens[j] + p.B - 1) / p.B * p.L * p.H;
    if (j < i) base += nb;
    total_need += nb;
  }
  if (total_need > total_free) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.status = 1;
    return;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *p.free_count = total_free - total_need;
  const int len = p.seq_lens[i];
  const int nblk = (len + p.B - 1) / p.B;
  const int slot = p.seq_slots[i];
  const int e = blockIdx.x * 256 + threadIdx.x;  // (l, h, b) row-major = the reshape at :207-209
  if (e < p.L * p.H) {  // one thread per head also sets the context length (:212)
    const int l = e / p.H, h = e % p.H;
    p.context_lens[(l * p.max_seqs + slot) * p.H + h] = len;
  }
  if (e >= p.L * p.H * nblk) return;
  const int b = e % nblk, h = (e / nblk) % p.H, l = e / (nblk * p.H);
  const int blk = find_free_block(p.ix, s_super, base + e);
  p.block_tables[((int64_t)(l * p.max_seqs + slot) * p.H + h) * p.MB + b] = blk;
  p.free_mask[blk] = 0;
  if (p.seq_index_by_block) p.seq_index_by_block[blk] = slot;
  if (p.layer_index_by_block) p.layer_index_by_block[blk] = l;
  if (p.head_index_by_block) p.head_index_by_block[blk] = h;
  if (p.logical_block_num_by_block) p.logical_block_num_by_block[blk] = b;
  if (p.kv_position)  // token position == logical index right after prefill (block.py:423-428)
    for (int j = 0; j < p.B; ++j) p.kv_position[(int64_t)blk * p.B + j] = b * p.B + j;
  (void)max_nblk;
}

// ---- free the trailing blocks of compressed heads + shrink the context lengths ----
// one thread per (batch position, l, h) of evicted_block_count [num_batch, L, H]; one atomic per CTA
__global__ void __launch_bounds__(256) alloc_free_compressed_kernel(
    uint8_t* free_mask, int* free_count, const int* block_tables, int* context_lens,
    const int* evicted_block_count, const int* seq_slots, int* seq_index_by_block, int num_batch, int max_seqs,
    int L, int H, int MB, int B) {
  __shared__ int s_freed;
  if (threadIdx.x == 0) s_freed = 0;
  __syncthreads();
  const int g = blockIdx.x * 256 + threadIdx.x;
  int freed = 0;
  if (g < num_batch * L * H) {
    const int k = evicted_block_count[g];
    if (k > 0) {
      const int h = g % H, l = (g / H) % L, sb = g / (L * H);
      const int slot = seq_slots ? seq_slots[sb] : sb;
      const int lsh = (l * max_seqs + slot) * H + h;
      const int ctx = context_lens[lsh];
      const int nblk = (ctx + B - 1) / B;   // block.py:337-355, 367-379: the last k allocated blocks
      const int* bt = block_tables + (int64_t)lsh * MB;
      for (int j = max(nblk - k, 0); j < nblk; ++j) {
        const int blk = bt[j];
        free_mask[blk] = 1;
        if (seq_index_by_block) seq_index_by_block[blk] = -1;   // metrics.py:364-370
      }
      freed = min(k, nblk);
      const int hanging = (ctx % B) ? (ctx % B) : B;            // block.py:330-335
      context_lens[lsh] = ctx - max(k * B - (B - hanging), 0);  // block.py:184-210
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) freed += __shfl_xor_sync(0xffffffffu, freed, d);
  if ((threadIdx.x & 31) == 0 && freed) atomicAdd(&s_freed, freed);
  __syncthreads();
  if (threadIdx.x == 0 && s_freed) atomicAdd(free_count, s_freed);
}

// ---- remove sequences: free every allocated block, context length 0 ----
__global__ void __launch_bounds__(256) alloc_remove_kernel(uint8_t* free_mask, int* free_count,
                                                          const int* block_tables, int* context_lens,
                                                          const int* seq_slots, int* seq_index_by_block, int n,
                                                          int max_seqs, int L, int H, int MB, int B) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= n * L * H) return;
  const int h = g % H, l = (g / H) % L, i = g / (L * H);
  const int lsh = (l * max_seqs + seq_slots[i]) * H + h;
  const int ctx = context_lens[lsh];
  const int nblk = (ctx + B - 1) / B;
  const int* bt = block_tables + (int64_t)lsh * MB;
  for (int j = lane; j < nblk; j += 32) {
    const int blk = bt[j];
    free_mask[blk] = 1;
    if (seq_index_by_block) seq_index_by_block[blk] = -1;
  }
  if (lane == 0) {
    if (nblk) atomicAdd(free_count, nblk);
    context_lens[lsh] = 0;
  }
}

__global__ void alloc_init_kernel(uint8_t* free_mask, int* free_count, int* status, int64_t num_blocks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < num_blocks) free_mask[i] = 1;
  if (i == 0) {
    *free_count = (int)num_blocks;
    if (status) *status = 0;
  }
}

static int make_index(AllocIndex& ix, const uint8_t* free_mask, int64_t num_blocks, void* workspace,
                      size_t workspace_bytes, size_t extra_ints, int** extra) {
  ix.free_mask = free_mask;
  ix.num_blocks = num_blocks;
  ix.n_chunks = (int)cdiv64(num_blocks, kChunk);
  ix.n_super = (ix.n_chunks + kSuper - 1) / kSuper;
  const size_t need = ((size_t)ix.n_chunks * 9 + ix.n_super + extra_ints) * sizeof(int) + 256;
  if (!workspace || workspace_bytes < need) return -1;
  uintptr_t w = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  ix.bitmap = reinterpret_cast<uint32_t*>(w);  // 32 B per chunk, 256 B aligned
  ix.chunk_prefix = reinterpret_cast<int*>(ix.bitmap + (size_t)ix.n_chunks * 8);
  ix.super_total = ix.chunk_prefix + ix.n_chunks;
  if (extra) *extra = ix.super_total + ix.n_super;
  return 0;
}

}  // namespace kvc

using namespace kvc;

extern "C" size_t kvc_alloc_workspace_bytes(int64_t num_blocks, int num_batch, int num_layers, int num_kv_heads) {
  if (num_blocks <= 0) return 0;
  const size_t n_chunks = (size_t)cdiv64(num_blocks, kChunk);
  const size_t n_super = (n_chunks + kSuper - 1) / kSuper;
  const size_t heads = (size_t)(num_batch > 0 ? num_batch : 0) * num_layers * num_kv_heads;
  const size_t n_need = (heads + kNeedCta - 1) / kNeedCta;
  return (n_chunks * 9 + n_super + n_need) * sizeof(int) + 512;
}

extern "C" int kvc_alloc_init(uint8_t* free_mask, int* free_count, int* status, int64_t num_blocks,
                              kvc_stream_t stream) {
  KVC_REQUIRE(free_mask && free_count && num_blocks > 0 && num_blocks < (int64_t)1 << 31, KVC_ERR_INVALID,
              "alloc_init: bad arguments");
  alloc_init_kernel<<<(unsigned)cdiv64(num_blocks, 256), 256, 0, as_stream(stream)>>>(free_mask, free_count, status,
                                                                                      num_blocks);
  KVC_CHECK_LAUNCH("alloc_init_kernel");
  return KVC_OK;
}

extern "C" int kvc_alloc_append_slots(
    uint8_t* free_mask, int* free_count, int* status, int64_t num_blocks, int* block_tables, int* context_lens,
    const int* seq_slots, const int* seq_len, int64_t* slot_mapping, int* kv_position, int* seq_index_by_block,
    int* layer_index_by_block, int* head_index_by_block, int* logical_block_num_by_block, int num_batch,
    int max_seqs, int num_layers, int num_kv_heads, int max_num_blocks_per_seq, int block_size, void* workspace,
    size_t workspace_bytes, kvc_stream_t stream) {
  KVC_REQUIRE(free_mask && free_count && status && block_tables && context_lens && seq_len && slot_mapping &&
                  kv_position,
              KVC_ERR_INVALID, "alloc_append_slots: null tensor");
  KVC_REQUIRE(num_batch >= 0 && num_batch <= max_seqs && num_layers > 0 && num_kv_heads > 0 &&
                  max_num_blocks_per_seq > 0 && block_size > 0 && num_blocks > 0,
              KVC_ERR_INVALID, "alloc_append_slots: bad sizes");
  if (num_batch == 0) return KVC_OK;
  AppendParams p;
  const int total = num_layers * num_batch * num_kv_heads;
  p.n_need_ctas = (total + kNeedCta - 1) / kNeedCta;
  KVC_REQUIRE(make_index(p.ix, free_mask, num_blocks, workspace, workspace_bytes, p.n_need_ctas, &p.need_partials) == 0,
              KVC_ERR_WORKSPACE, "alloc_append_slots: workspace of %zu bytes required, got %zu",
              kvc_alloc_workspace_bytes(num_blocks, num_batch, num_layers, num_kv_heads), workspace_bytes);
  KVC_REQUIRE((size_t)(p.ix.n_super + 1) * sizeof(int) <= 40 * 1024, KVC_ERR_UNSUPPORTED,
              "alloc_append_slots: %lld blocks exceed the index capacity", (long long)num_blocks);
  p.free_mask = free_mask; p.free_count = free_count; p.status = status;
  p.block_tables = block_tables; p.context_lens = context_lens; p.seq_slots = seq_slots; p.seq_len = seq_len;
  p.slot_mapping = slot_mapping; p.kv_position = kv_position; p.seq_index_by_block = seq_index_by_block;
  p.layer_index_by_block = layer_index_by_block; p.head_index_by_block = head_index_by_block;
  p.logical_block_num_by_block = logical_block_num_by_block;
  p.num_batch = num_batch; p.max_seqs = max_seqs; p.L = num_layers; p.H = num_kv_heads;
  p.MB = max_num_blocks_per_seq; p.B = block_size;
  cudaStream_t st = as_stream(stream);
  const int count_ctas = (int)cdiv64(num_blocks, 8 * kCountChunksPerWarp * kChunk);
  alloc_count_kernel<<<count_ctas + p.n_need_ctas, 256, 0, st>>>(p, count_ctas);
  KVC_CHECK_LAUNCH("alloc_count_kernel");
  alloc_scan_kernel<<<p.ix.n_super, kSuper, 0, st>>>(p.ix);
  KVC_CHECK_LAUNCH("alloc_scan_kernel");
  alloc_append_kernel<<<p.n_need_ctas, kNeedCta, (p.ix.n_super + 1) * sizeof(int), st>>>(p);
  KVC_CHECK_LAUNCH("alloc_append_kernel");
  return KVC_OK;
}

extern "C" int kvc_alloc_add_sequences(
    uint8_t* free_mask, int* free_count, int* status, int64_t num_blocks, int* block_tables, int* context_lens,
    const int* seq_slots, const int* seq_lens, int* kv_position, int* seq_index_by_block,
    int* layer_index_by_block, int* head_index_by_block, int* logical_block_num_by_block, int num_new,
    int max_seq_len, int max_seqs, int num_layers, int num_kv_heads, int max_num_blocks_per_seq, int block_size,
    void* workspace, size_t workspace_bytes, kvc_stream_t stream) {
  KVC_REQUIRE(free_mask && free_count && status && block_tables && context_lens && seq_slots && seq_lens,
              KVC_ERR_INVALID, "alloc_add_sequences: null tensor");
  KVC_REQUIRE(num_new >= 0 && num_new <= max_seqs && num_layers > 0 && num_kv_heads > 0 && block_size > 0 &&
                  max_seq_len >= 0 && num_blocks > 0,
              KVC_ERR_INVALID, "alloc_add_sequences: bad sizes");
  const int max_nblk = (max_seq_len + block_size - 1) / block_size;
  KVC_REQUIRE(max_nblk <= max_num_blocks_per_seq, KVC_ERR_INVALID,
              "alloc_add_sequences: %d tokens need %d blocks per head, table holds %d", max_seq_len, max_nblk,
              max_num_blocks_per_seq);
  if (num_new == 0) return KVC_OK;
  AddParams p;
  KVC_REQUIRE(make_index(p.ix, free_mask, num_blocks, workspace, workspace_bytes, 0, nullptr) == 0, KVC_ERR_WORKSPACE,
              "alloc_add_sequences: workspace of %zu bytes required, got %zu",
              kvc_alloc_workspace_bytes(num_blocks, 0, num_layers, num_kv_heads), workspace_bytes);
  KVC_REQUIRE((size_t)(p.ix.n_super + 1) * sizeof(int) <= 40 * 1024, KVC_ERR_UNSUPPORTED,
              "alloc_add_sequences: %lld blocks exceed the index capacity", (long long)num_blocks);
  p.free_mask = free_mask; p.free_count = free_count; p.status = status;
  p.block_tables = block_tables; p.context_lens = context_lens; p.seq_slots = seq_slots; p.seq_lens = seq_lens;
  p.kv_position = kv_position; p.seq_index_by_block = seq_index_by_block;
  p.layer_index_by_block = layer_index_by_block; p.head_index_by_block = head_index_by_block;
  p.logical_block_num_by_block = logical_block_num_by_block;
  p.n = num_new; p.max_seqs = max_seqs; p.L = num_layers; p.H = num_kv_heads;
  p.MB = max_num_blocks_per_seq; p.B = block_size;
  cudaStream_t st = as_stream(stream);
  alloc_count_only_kernel<<<(unsigned)cdiv64(num_blocks, 8 * kCountChunksPerWarp * kChunk), 256, 0, st>>>(p.ix);
  KVC_CHECK_LAUNCH("alloc_count_only_kernel");
  alloc_scan_kernel<<<p.ix.n_super, kSuper, 0, st>>>(p.ix);
  KVC_CHECK_LAUNCH("alloc_scan_kernel");
  const int per_seq = num_layers * num_kv_heads * (max_nblk > 0 ? max_nblk : 1);
  const dim3 grid((per_seq + 255) / 256, num_new);
  alloc_add_kernel<<<grid, 256, (p.ix.n_super + 1) * sizeof(int), st>>>(p, max_nblk);
  KVC_CHECK_LAUNCH("alloc_add_kernel");
  return KVC_OK;
}

// the same with one WARP per head (few heads that free hundreds of blocks each: the compression right after a
// prefill; the thread-per-head form above would walk 367 blocks serially in each of 1024 threads)
__global__ void __launch_bounds__(256) alloc_free_compressed_warp_kernel(
    uint8_t* free_mask, int* free_count, const int* block_tables, int* context_lens,
    const int* evicted_block_count, const int* seq_slots, int* seq_index_by_block, int num_batch, int max_seqs,
    int L, int H, int MB, int B) {
  const int g = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (g >= num_batch * L * H) return;
  const int k = evicted_block_count[g];
  if (k <= 0) return;
  const int h = g % H, l = (g / H) % L, sb = g / (L * H);
  const int slot = seq_slots ? seq_slots[sb] : sb;
  const int lsh = (l * max_seqs + slot) * H + h;
  const int ctx = context_lens[lsh];
  const int nblk = (ctx + B - 1) / B;   // block.py:337-355, 367-379: the last k allocated blocks
  const int* bt = block_tables + (int64_t)lsh * MB;
  for (int j = max(nblk - k, 0) + lane; j < nblk; j += 32) {
    const int blk = bt[j];
    free_mask[blk] = 1;
    if (seq_index_by_block) seq_index_by_block[blk] = -1;   // metrics.py:364-370
  }
  __syncwarp();
  if (lane == 0) {
    const int hanging = (ctx % B) ? (ctx % B) : B;            // block.py:330-335
    context_lens[lsh] = ctx - max(k * B - (B - hanging), 0);  // block.py:184-210
    atomicAdd(free_count, min(k, nblk));
  }
}

extern "C" int kvc_alloc_free_compressed(uint8_t* free_mask, int* free_count, const int* block_tables,
                                         int* context_lens, const int* evicted_block_count, const int* seq_slots,
                                         int* seq_index_by_block, int num_batch, int max_seqs, int num_layers,
                                         int num_kv_heads, int max_num_blocks_per_seq, int block_size,
                                         kvc_stream_t stream) {
  KVC_REQUIRE(free_mask && free_count && block_tables && context_lens && evicted_block_count, KVC_ERR_INVALID,
              "alloc_free_compressed: null tensor");
  KVC_REQUIRE(num_batch >= 0 && num_batch <= max_seqs && num_layers > 0 && num_kv_heads > 0 && block_size > 0,
              KVC_ERR_INVALID, "alloc_free_compressed: bad sizes");
  if (num_batch == 0) return KVC_OK;
  const int heads = num_batch * num_layers * num_kv_heads;
  if (heads < 16384) {   // post-prefill calls: long per-head lists
    alloc_free_compressed_warp_kernel<<<(heads + 7) / 8, 256, 0, as_stream(stream)>>>(
        free_mask, free_count, block_tables, context_lens, evicted_block_count, seq_slots, seq_index_by_block,
        num_batch, max_seqs, num_layers, num_kv_heads, max_num_blocks_per_seq, block_size);
    KVC_CHECK_LAUNCH("alloc_free_compressed_warp_kernel");
    return KVC_OK;
  }
  alloc_free_compressed_kernel<<<(heads + 255) / 256, 256, 0, as_stream(stream)>>>(
      free_mask, free_count, block_tables, context_lens, evicted_block_count, seq_slots, seq_index_by_block,
      num_batch, max_seqs, num_layers, num_kv_heads, max_num_blocks_per_seq, block_size);
  KVC_CHECK_LAUNCH("alloc_free_compressed_kernel");
  return KVC_OK;
}

extern "C" int kvc_alloc_remove_sequences(uint8_t* free_mask, int* free_count, const int* block_tables,
                                          int* context_lens, const int* seq_slots, int* seq_index_by_block,
                                          int num_removed, int max_seqs, int num_layers, int num_kv_heads,
                                          int max_num_blocks_per_seq, int block_size, kvc_stream_t stream) {
  KVC_REQUIRE(free_mask && free_count && block_tables && context_lens && seq_slots, KVC_ERR_INVALID,
              "alloc_remove_sequences: null tensor");
  KVC_REQUIRE(num_removed >= 0 && num_removed <= max_seqs && num_layers > 0 && num_kv_heads > 0 && block_size > 0,
              KVC_ERR_INVALID, "alloc_remove_sequences: bad sizes");
  if (num_removed == 0) return KVC_OK;
  const int heads = num_removed * num_layers * num_kv_heads;
  alloc_remove_kernel<<<(heads + 7) / 8, 256, 0, as_stream(stream)>>>(
      free_mask, free_count, block_tables, context_lens, seq_slots, seq_index_by_block, num_removed, max_seqs,
      num_layers, num_kv_heads, max_num_blocks_per_seq, block_size);
  KVC_CHECK_LAUNCH("alloc_remove_kernel");
  return KVC_OK;
}
