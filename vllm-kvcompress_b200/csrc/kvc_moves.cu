// K5 execute_cache_moves: the paged K/V gather -> compact.
// reference: csrc/kvcompress_eviction_kernels.cu:359-435 (kernel), :858-929 (launcher)
//
// Layout (one physical block = B tokens of ONE kv head, B=16, D=128, 2-byte cache):
//   K block [16 j][16 t][8 elem]  -> a token's K = 16 chunks of 16 B at 256 B stride
//   V block [128 d][16 t]         -> a token's V = 128 x 2 B at 32 B stride, i.e. one
//                                    element in each of the block's 128 sectors
// so touching ONE token touches the whole 4 KB V block in DRAM.  The reference moves a
// token with 256+ scalar 2-byte loads and stores per thread; every V store is a 2-byte
// partial write of its own 32 B sector (128 sector writes per moved token).
//
// Here the unit of work is a *run*: the consecutive moves of a head that land in the
// same destination block (schedule_t1_cache_moves emits dst ascending, so they are
// adjacent; <= 16 per run).  One warp owns a run, no shared memory, no barriers:
//   - V: lane <-> rows d = lane + 32 r.  The lane gathers its 4 rows of every source
//     token of the run into registers (4 x 16 tokens x 2 B), which is exactly the
//     destination rows it owns, and writes them back as whole 32 B rows (two 16 B
//     stores, consecutive lanes -> consecutive rows -> 1 KB per instruction pair).
//     Partially covered rows are written with the widest aligned pieces the run's
//     token mask allows (8/4/2 B); only owned bytes are ever written, so there is no
//     read-modify-write of the destination and no hazard between runs.
//   - K: half-warp <-> the 16 chunks of one token, 16 B load + 16 B store per lane.
//   - metric (f32) and position (i32): lanes 0-15 / 16-31, one element each.
// Algorithmic bytes per move: 2*(2*128*2 + 4 + 4) + 8 = 1048.
// Physical floor: >= 4 KB (whole V block) + 16 half-sectors of K per isolated source token.
#include <stdlib.h>

#include "kvc_common.cuh"

namespace kvc {

// -----------------------------------------------------------------------------
// production layout: B=16, D=128, 2-byte elements, x=8
// -----------------------------------------------------------------------------
constexpr int kMovesPerSegment = 32;
constexpr int kMoveWarps = 4;

__device__ __forceinline__ void store_v_row_partial(uint16_t* __restrict__ row, const uint32_t (&w)[8],
                                                    unsigned mask) {
  // row: 16 tokens x 2 B of one d; write only tokens in `mask`, widest aligned pieces first
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // quads of tokens = 8 B
    const unsigned mq = (mask >> (4 * q)) & 0xfu;
    if (mq == 0xfu) {
      *reinterpret_cast<uint2*>(row + 4 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
    } else {
#pragma unroll
      for (int p = 0; p < 2; ++p) {  // pairs = 4 B
        const unsigned mp = (mq >> (2 * p)) & 0x3u;
        const uint32_t word = w[2 * q + p];
        if (mp == 0x3u) {
          *reinterpret_cast<uint32_t*>(row + 4 * q + 2 * p) = word;
        } else if (mp == 0x1u) {
          row[4 * q + 2 * p] = (uint16_t)(word & 0xffffu);
        } else if (mp == 0x2u) {
          row[4 * q + 2 * p + 1] = (uint16_t)(word >> 16);
        }
      }
    }
  }
}

// PEND (deferred append, see kvc_paged_attention_fused_step_deferred): the K/V of the head's newest
// token may live in pend_kv [L, S, H, 2 * D] instead of the cache (its slot in pend_slot [L, S, H], -1 =
// none).  A move whose source is that slot reads the compact copy; a pending token that no move
// consumes is stored to its own slot, so the cache is canonical when the kernel completes.
template <bool PEND>
__device__ __forceinline__ void moves_run_head(
    uint16_t* __restrict__ k_cache,        // [NB, 16, 16, 8]
    uint16_t* __restrict__ v_cache,        // [NB, 128, 16]
    float* __restrict__ kv_metrics,        // [NB, 16]
    int* __restrict__ kv_position,         // [NB, 16]
    const int2* __restrict__ moves,        // [max_evicted_kv] (dst, src)
    const int* __restrict__ move_count,    // [S*L*H]
    const int* __restrict__ move_offsets,  // [S*L*H]
    const uint16_t* __restrict__ pend_kv, const int64_t* __restrict__ pend_slot, int num_seqs,
    int num_layers, int num_kv_heads, int g, int lane) {
  constexpr int B = 16, D = 128;
  constexpr int kBlockElems = B * D;  // 2048 two-byte elements = 4 KB
  const int cnt = move_count[g];
  const int off = move_offsets[g];
  const int nseg = (cnt + kMovesPerSegment - 1) / kMovesPerSegment;
  const int2* __restrict__ mv = moves + off;
  int pslot = -1;                           // slot of the head's pending token (PEND)
  const uint16_t* __restrict__ pk = nullptr;
  bool consumed = false;
  if (PEND) {
    const int sq = g / (num_layers * num_kv_heads);
    const int ly = (g / num_kv_heads) % num_layers;
    const int64_t pidx = ((int64_t)ly * num_seqs + sq) * num_kv_heads + (g % num_kv_heads);
    pslot = (int)pend_slot[pidx];
    pk = pend_kv + pidx * (2 * D);
  }

  for (int seg = blockIdx.y; seg < nseg; seg += gridDim.y) {
    const int base = seg * kMovesPerSegment;
    // window of 48 moves: [base, base+32) in `a`, lookahead [base+32, base+48) in `la`
    const bool a_ok = base + lane < cnt;
    const int2 a = a_ok ? mv[base + lane] : make_int2(-1, -1);
    const bool la_ok = (lane < 16) && (base + 32 + lane < cnt);
    const int2 la = la_ok ? mv[base + 32 + lane] : make_int2(-1, -1);
    const int a_blk = a_ok ? (a.x >> 4) : -1;
    const int la_blk = la_ok ? (la.x >> 4) : -2;
    int prev_blk = __shfl_up_sync(0xffffffffu, a_blk, 1);
    if (lane == 0) prev_blk = (base > 0) ? (mv[base - 1].x >> 4) : -3;
    unsigned starts = __ballot_sync(0xffffffffu, a_ok && a_blk != prev_blk);

    while (starts) {  // warp-uniform loop over the runs that START in this segment
      const int l0 = __ffs(starts) - 1;
      starts &= starts - 1;
      const int run_blk = __shfl_sync(0xffffffffu, a_blk, l0);
      // contiguous members: lanes [l0, end1) of `a`, then a prefix of `la` if the run reaches lane 31
      const unsigned same_a = __ballot_sync(0xffffffffu, a_ok && a_blk == run_blk);
      const unsigned from_l0 = same_a >> l0;                       // bit 0 = lane l0
      const int len1 = (from_l0 == 0xffffffffu >> l0) ? (32 - l0) : __ffs(~from_l0) - 1;
      const bool in1 = lane >= l0 && lane < l0 + len1;
      int len2 = 0;
      if (l0 + len1 == 32) {
        const unsigned same_la = __ballot_sync(0xffffffffu, la_ok && la_blk == run_blk);
        len2 = (same_la == 0xffffffffu) ? 32 : __ffs(~same_la) - 1;
      }
      const bool in2 = lane < len2;

      // per destination offset o: source slot (warp-uniform) or -1
      int src[16];
      unsigned mask = 0;
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const unsigned c1 = __ballot_sync(0xffffffffu, in1 && (a.x & 15) == o);
        const unsigned c2 = __ballot_sync(0xffffffffu, in2 && (la.x & 15) == o);
        const int s1 = __shfl_sync(0xffffffffu, a.y, c1 ? __ffs(c1) - 1 : 0);
        const int s2 = __shfl_sync(0xffffffffu, la.y, c2 ? __ffs(c2) - 1 : 0);
        src[o] = c1 ? s1 : (c2 ? s2 : -1);
        mask |= (c1 | c2) ? (1u << o) : 0u;
      }

      // ---- metric + position: lanes 0-15 metric, 16-31 position; loaded here, stored after the gather so that
      // the load's latency is not paid before the V loads are even issued (in-order issue) ----
      int mp_src = src[0];
      uint32_t mp_val = 0u;
      {
        const int o = lane & 15;
#pragma unroll
        for (int i = 1; i < 16; ++i) mp_src = (o == i) ? src[i] : mp_src;
        if (mp_src >= 0)
          mp_val = (lane < 16) ? __float_as_uint(kv_metrics[mp_src]) : (uint32_t)kv_position[mp_src];
      }

      // ---- K: half-warp per token, lane j copies one 16 B chunk (reference :413-424).  Loads first: they are in
      // flight together with the V loads below; the stores follow the gather ----
      const int half = lane >> 4, j = lane & 15;
      uint4 kv[8];
      bool kp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int s = half ? src[2 * i + 1] : src[2 * i];
        kp[i] = s >= 0;
        if (kp[i]) {
          if (PEND && s == pslot)
            kv[i] = *reinterpret_cast<const uint4*>(pk + j * 8);
          else
            kv[i] = ld_stream16(k_cache + (int64_t)(s >> 4) * kBlockElems + j * (B * 8) + (s & 15) * 8);
        }
      }

      // ---- V gather into registers: lane owns rows d = lane + 32 r ----
      uint32_t w[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[r][i] = 0u;
      if (!PEND) {
        // all 64 element loads of the run are issued before the first one is used: the warp pays one memory
        // round trip per run, not one per source token (in-order issue stalls at the first use of a load)
        uint32_t x[16][4];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const int s = src[o];
          const bool ok = s >= 0;  // warp-uniform
          const uint16_t* __restrict__ sp =
              v_cache + (int64_t)((ok ? s : 0) >> 4) * kBlockElems + ((ok ? s : 0) & 15) + lane * B;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            x[o][r] = 0u;
            if (ok) x[o][r] = sp[r * 32 * B];  // reference :426-431, element d = lane + 32 r
          }
        }
#pragma unroll
        for (int o = 0; o < 16; ++o)
#pragma unroll
          for (int r = 0; r < 4; ++r) w[r][o >> 1] |= (o & 1) ? (x[o][r] << 16) : x[o][r];
      }
#pragma unroll
      for (int o = 0; o < 16; ++o) {
        const int s = src[o];
        if (PEND && s >= 0) {  // warp-uniform
          if (PEND && s == pslot) {
            consumed = true;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const uint32_t x = pk[D + lane + 32 * r];
              w[r][o >> 1] |= (o & 1) ? (x << 16) : x;
            }
          } else {
            const uint16_t* __restrict__ sp =
                v_cache + (int64_t)(s >> 4) * kBlockElems + (s & 15) + lane * B;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const uint32_t x = sp[r * 32 * B];  // reference :426-431, element d = lane + 32 r
              w[r][o >> 1] |= (o & 1) ? (x << 16) : x;
            }
          }
        }
      }

#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (kp[i])
          st_stream16(k_cache + (int64_t)run_blk * kBlockElems + j * (B * 8) + (2 * i + half) * 8, kv[i]);
      }
      if (mp_src >= 0) {
        const int64_t dslot = (int64_t)run_blk * B + (lane & 15);
        if (lane < 16) kv_metrics[dslot] = __uint_as_float(mp_val);   // reference :410
        else           kv_position[dslot] = (int)mp_val;              // reference :411
      }

      // ---- V write-back: whole rows when the run covers all 16 tokens ----
      uint16_t* __restrict__ dp = v_cache + (int64_t)run_blk * kBlockElems + lane * B;
      if (mask == 0xffffu) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uint4* row = reinterpret_cast<uint4*>(dp + r * 32 * B);
          st_stream16(row, make_uint4(w[r][0], w[r][1], w[r][2], w[r][3]));
          st_stream16(row + 1, make_uint4(w[r][4], w[r][5], w[r][6], w[r][7]));
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) store_v_row_partial(dp + r * 32 * B, w[r], mask);
      }
    }
  }
  if (PEND && pslot >= 0 && blockIdx.y == 0) {
    // With gridDim.y > 1 other warps own some of the head's segments: look through the whole list.
    if (gridDim.y > 1) {
      bool hit = false;
      for (int i = lane; i < cnt; i += kWarp) hit |= (mv[i].y == pslot);
      consumed = __any_sync(0xffffffffu, hit);
    }
    if (!consumed) {  // not moved: the token stays where batch_append_slots put it
      uint16_t* kd = k_cache + (int64_t)(pslot >> 4) * kBlockElems + (pslot & 15) * 8;
      uint16_t* vd = v_cache + (int64_t)(pslot >> 4) * kBlockElems + (pslot & 15);
      if (lane < 16) *reinterpret_cast<uint4*>(kd + lane * (B * 8)) = *reinterpret_cast<const uint4*>(pk + lane * 8);
#pragma unroll
      for (int r = 0; r < 4; ++r) vd[(lane + 32 * r) * B] = pk[D + lane + 32 * r];
    }
  }
}

// one WARP per head (post-prefill compaction: few heads, thousands of moves each); blockIdx.y strides over the
// head's 32-move segments
template <bool PEND>
__global__ void __launch_bounds__(kMoveWarps * kWarp) execute_moves_b16_d128_kernel(
    uint16_t* __restrict__ k_cache, uint16_t* __restrict__ v_cache, float* __restrict__ kv_metrics,
    int* __restrict__ kv_position, const int2* __restrict__ moves, const int* __restrict__ move_count,
    const int* __restrict__ move_offsets, int total_heads, const uint16_t* __restrict__ pend_kv,
    const int64_t* __restrict__ pend_slot, int num_seqs, int num_layers, int num_kv_heads) {
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int g = blockIdx.x * kMoveWarps + wid;
  if (g >= total_heads) return;
  moves_run_head<PEND>(k_cache, v_cache, kv_metrics, kv_position, moves, move_count, move_offsets, pend_kv,
                       pend_slot, num_seqs, num_layers, num_kv_heads, g, lane);
}

// -----------------------------------------------------------------------------
// Post-prefill compaction (few heads, thousands of moves each): whole blocks STAGED BY TMA, lane = token.
//
// The register-gather kernel above pulls every V element of a source token with its own 2-byte global load
// (2048 sector requests per run) behind a warp-uniform decode of ~1700 instructions per run, and is bound by
// instruction latency at 16 warps per SM (ncu: 22 % warps active, 62 % of DRAM peak at CR 4).  Here the bytes
// move through the async proxy and a warp executes ~300 instructions per run:
//   * unit of work = a 32-move segment of a head's list; its runs (moves into one destination block) are
//     processed two at a time: lane (h, o) owns token o of run h.  A moved token's source is the slot K4 gave
//     it; an UNMOVED token of a partly overwritten block takes its old column from the destination block
//     itself, so every lane copies one column from some staged block and the block leaves the SM complete;
//   * source V blocks (and the old destination block of a partial run) are bulk-copied whole into a private ring
//     (cp.async.bulk, 4 KB each, one mbarrier per slot; any one token needs all 128 sectors of its block
//     anyway).  K4 emits dst ascending <-> src descending, so tokens that share a source block are neighbours,
//     also across the two runs: the block is fetched once.  kPrSlots copies are issued by kPrSlots lanes in
//     parallel per round;
//   * the lane copies its column, 128 x (2-byte LDS -> 2-byte STS), into a 4 KB staging image of the destination
//     V block; its sixteen 16-byte K chunks go global -> staging with cp.async (no registers);
//   * the finished K and V images leave with ONE bulk store each (cp.async.bulk shared -> global, 4 KB):
//     full-sector writes only, no partial-sector writes for L2 to fill.
// Nobody writes a block concurrently with different bytes: a block has one owner head, a destination block one
// run, sources are never destinations; the boundary block that holds both destinations and sources is rewritten
// with identical bytes in its source columns.  Results are bit-identical to the register kernel.
// -----------------------------------------------------------------------------
constexpr int kPrWarps = 5;
constexpr int kPrSlots = 6;
constexpr int kPrBlockBytes = 4096;
constexpr int kPrWarpBytes = (kPrSlots + 4) * kPrBlockBytes + 512;  // ring | V images | K images | tables + barriers

__device__ __forceinline__ uint32_t mv_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mv_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MVWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MVDONE_%=;\n\t"
      "bra MVWAIT_%=;\n\t"
      "MVDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(kPrWarps * kWarp, 1) execute_moves_staged_kernel(
    uint16_t* __restrict__ k_cache, uint16_t* __restrict__ v_cache, float* __restrict__ kv_metrics,
    int* __restrict__ kv_position, const int2* __restrict__ moves, const int* __restrict__ move_count,
    const int* __restrict__ move_offsets, int total_heads) {
  constexpr int B = 16;
  extern __shared__ __align__(128) unsigned char mv_smem[];
  __shared__ int s_max_cnt;
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  // longest move list of the call (fewer than 4096 heads on this path): bounds the unit enumeration below
  if (threadIdx.x == 0) s_max_cnt = 0;
  __syncthreads();
  {
    int m = 0;
    for (int i = threadIdx.x; i < total_heads; i += kPrWarps * kWarp) m = max(m, move_count[i]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0 && m > 0) atomicMax(&s_max_cnt, m);
  }
  __syncthreads();
  const int max_segments = (s_max_cnt + kMovesPerSegment - 1) / kMovesPerSegment;

  unsigned char* base_w = mv_smem + (size_t)wid * kPrWarpBytes;
  unsigned char* ring = base_w;
  unsigned char* vst = base_w + kPrSlots * kPrBlockBytes;        // [2][4096] destination V images
  unsigned char* kst = vst + 2 * kPrBlockBytes;                  // [2][4096] destination K images
  int* tsrc = reinterpret_cast<int*>(kst + 2 * kPrBlockBytes);   // [32] source slot of (half, offset)
  int* tblk = tsrc + 32;                                         // [2]  destination block of each half (+2 pad)
  int* blist = tblk + 4;                                         // [40] distinct blocks to stage, in order
  const uint32_t bars = mv_smem_u32(blist + 40);                 // [kPrSlots] mbarriers
  const uint32_t ring_u = mv_smem_u32(ring);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kPrSlots; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * i), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t phases = 0;   // bit i: parity the next wait on slot i expects
  const unsigned lt = lanemask_lt();
  const unsigned le = lt | (1u << lane);
  const int h = lane >> 4, o = lane & 15;
  unsigned char* const k_bytes = reinterpret_cast<unsigned char*>(k_cache);
  unsigned char* const v_bytes = reinterpret_cast<unsigned char*>(v_cache);

  // persistent grid: units (segment, head) are dealt round-robin to the warps of the grid, segment-major so that
  // concurrently running warps work on different heads (different regions of the cache)
  const long long total_units = (long long)max_segments * total_heads;
  const int W = gridDim.x * kPrWarps;
  for (long long u = (long long)wid * gridDim.x + blockIdx.x; u < total_units; u += W) {
    const int seg = (int)(u / total_heads);
    const int g = (int)(u - (long long)seg * total_heads);
    const int cnt = move_count[g];
    const int base = seg * kMovesPerSegment;
    if (base >= cnt) continue;
    const int2* __restrict__ mv = moves + move_offsets[g];
    // window of 48 moves: [base, base+32) in `a`, lookahead [base+32, base+48) in `la`
    const bool a_ok = base + lane < cnt;
    const int2 a = a_ok ? mv[base + lane] : make_int2(-1, -1);
    const bool la_ok = (lane < 16) && (base + 32 + lane < cnt);
    const int2 la = la_ok ? mv[base + 32 + lane] : make_int2(-1, -1);
    const int a_blk = a_ok ? (a.x >> 4) : -1;
    const int la_blk = la_ok ? (la.x >> 4) : -2;
    int prev_blk = __shfl_up_sync(0xffffffffu, a_blk, 1);
    if (lane == 0) prev_blk = (base > 0) ? (mv[base - 1].x >> 4) : -3;
    const unsigned starts = __ballot_sync(0xffffffffu, a_ok && a_blk != prev_blk);
    if (starts == 0) continue;  // every move here belongs to a run that started in the previous segment
    // runs that START in this segment, numbered 0 .. nruns-1; the last one may continue into the lookahead
    const int nruns = __popc(starts);
    const int f0 = __ffs(starts) - 1;
    const unsigned ok_a = __ballot_sync(0xffffffffu, a_ok);
    int len2_last = 0;
    if (ok_a == 0xffffffffu) {
      const int last_blk = __shfl_sync(0xffffffffu, a_blk, 31);
      const unsigned same_la = __ballot_sync(0xffffffffu, la_ok && la_blk == last_blk);
      len2_last = (same_la == 0xffffffffu) ? 32 : __ffs(~same_la) - 1;  // la_ok only on lanes < 16 -> <= 16
    }
    const int rank_a = (a_ok && lane >= f0) ? __popc(starts & le) - 1 : -1000;
    const int rank_la = (lane < len2_last) ? nruns - 1 : -1000;

    for (int p = 0; 2 * p < nruns; ++p) {
      // the previous pair's bulk stores have read the staging images
      if (lane < 4) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      tsrc[lane] = -1;
      if (lane < 2) tblk[lane] = -1;
      __syncwarp();
      {
        const int ha = rank_a - 2 * p, hl = rank_la - 2 * p;
        if (ha == 0 || ha == 1) { tsrc[ha * 16 + (a.x & 15)] = a.y; tblk[ha] = a.x >> 4; }
        if (hl == 0 || hl == 1) { tsrc[hl * 16 + (la.x & 15)] = la.y; tblk[hl] = la.x >> 4; }
      }
      __syncwarp();
      const int mysrc = tsrc[lane];
      const int myrun = tblk[h];          // destination block of my half, -1 = no run
      const bool moved = mysrc >= 0;
      const unsigned mvmask = __ballot_sync(0xffffffffu, moved);
      const bool partial = myrun >= 0 && ((mvmask >> (16 * h)) & 0xffffu) != 0xffffu;
      const bool active = moved || partial;
      const int sblk = moved ? (mysrc >> 4) : myrun;   // block my column comes from
      const int so = moved ? (mysrc & 15) : o;
      const unsigned pmask = __ballot_sync(0xffffffffu, partial);
      const int part_a = (pmask & 0xffffu) ? 1 : 0, part_b = (pmask >> 16) ? 1 : 0;
      const int nold = part_a + part_b;
      // distinct source blocks in lane order (tokens of one source block are adjacent among the moved lanes)
      const unsigned below = mvmask & lt;
      const int pl = below ? 31 - __clz(below) : 0;
      const int psb = __shfl_sync(0xffffffffu, sblk, pl);
      const bool isnew = moved && (below == 0 || psb != sblk);
      const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
      const int total_blocks = nold + __popc(newmask);
      const int ord = moved ? nold + __popc(newmask & le) - 1 : (h == 0 ? 0 : part_a);
      if (isnew || (partial && !moved)) blist[ord] = sblk;   // unmoved lanes of a half all write the same value
      __syncwarp();

      // ---- K: sixteen 16-byte chunks of my token, global -> destination image (async, no registers) ----
      if (active) {
        const unsigned char* ks = k_bytes + (int64_t)sblk * kPrBlockBytes + so * 16;
        const uint32_t kd = mv_smem_u32(kst + h * kPrBlockBytes + o * 16);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kd + j * 256), "l"(ks + j * 256) : "memory");
      }
      // ---- metric + position (reference :410-411) ----
      float mt = 0.f;
      int ps = 0;
      if (moved) {
        mt = kv_metrics[mysrc];
        ps = kv_position[mysrc];
      }
      // ---- V: rounds of kPrSlots staged blocks ----
      const int my_round = ord / kPrSlots, my_slot = ord - my_round * kPrSlots;
      const int nrounds = (total_blocks + kPrSlots - 1) / kPrSlots;
      for (int q = 0; q < nrounds; ++q) {
        const int n = q * kPrSlots + lane;
        if (lane < kPrSlots && n < total_blocks) {
          const uint32_t bar = bars + 8 * lane;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the slot
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPrBlockBytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(ring_u + lane * kPrBlockBytes), "l"(v_bytes + (int64_t)blist[n] * kPrBlockBytes),
                         "r"(kPrBlockBytes), "r"(bar) : "memory");
        }
        if (active && my_round == q) {
          mv_mbar_wait(bars + 8 * my_slot, (phases >> my_slot) & 1u);
          const uint16_t* sp = reinterpret_cast<const uint16_t*>(ring + my_slot * kPrBlockBytes) + so;
          uint16_t* dp = reinterpret_cast<uint16_t*>(vst + h * kPrBlockBytes) + o;
          // half 1 runs one row ahead of half 0: the two halves then hit disjoint shared-memory banks
#pragma unroll 16
          for (int r = 0; r < 128; ++r) {
            const int row = (r + h) & 127;
            dp[row * B] = sp[row * B];   // reference :426-431
          }
        }
        const int used = min(kPrSlots, total_blocks - q * kPrSlots);
        phases ^= (1u << used) - 1u;
        __syncwarp();  // every lane is done with this round's slots
      }
      if (moved) {
        const int64_t dslot = (int64_t)myrun * B + o;
        kv_metrics[dslot] = mt;
        kv_position[dslot] = ps;
      }
      // ---- the finished images leave as whole blocks ----
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> bulk store reads
      __syncwarp();
      if (lane < 4) {
        const int hh = lane & 1, is_k = lane >> 1;
        const int dblk = tblk[hh];
        if (dblk >= 0) {
          unsigned char* gdst = (is_k ? k_bytes : v_bytes) + (int64_t)dblk * kPrBlockBytes;
          const uint32_t ssrc = mv_smem_u32((is_k ? kst : vst) + hh * kPrBlockBytes);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(gdst), "r"(ssrc), "r"(kPrBlockBytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (lane < 4) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA retires
}

// decode-time compaction (many heads, most with <= 1 move, which execute_moves_light_kernel owns): a warp
// looks at 32 heads at once and runs only those with more than one move.  gridDim.y == 1.
template <bool PEND>
__global__ void __launch_bounds__(kMoveWarps * kWarp) execute_moves_b16_d128_chunk_kernel(
    uint16_t* __restrict__ k_cache, uint16_t* __restrict__ v_cache, float* __restrict__ kv_metrics,
    int* __restrict__ kv_position, const int2* __restrict__ moves, const int* __restrict__ move_count,
    const int* __restrict__ move_offsets, int total_heads, const uint16_t* __restrict__ pend_kv,
    const int64_t* __restrict__ pend_slot, int num_seqs, int num_layers, int num_kv_heads) {
  // warp w looks at heads w, w + W, w + 2W, ... (W = warps in the grid): heads with many moves cluster by
  // sequence, a strided assignment spreads them over the warps
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int W = gridDim.x * kMoveWarps;
  const int w = blockIdx.x * kMoveWarps + wid;
  for (int g0 = w; g0 < total_heads; g0 += W * kWarp) {
    const int g = g0 + lane * W;
    const int c = (g < total_heads) ? move_count[g] : 0;
    unsigned todo = __ballot_sync(0xffffffffu, c > 1);
    while (todo) {
      const int l0 = __ffs(todo) - 1;
      todo &= todo - 1;
      moves_run_head<PEND>(k_cache, v_cache, kv_metrics, kv_position, moves, move_count, move_offsets, pend_kv,
                           pend_slot, num_seqs, num_layers, num_kv_heads, g0 + l0 * W, lane);
    }
  }
}

// -----------------------------------------------------------------------------
// Heads with at most ONE move: the steady state of continual compression (every head evicts one token
// per step and the newest token takes its slot).  A warp looks at 32 heads at once (coalesced counts,
// offsets, moves, parked slots):
//   * deferred move (pend_dst given, the head's one move takes its parked token): the lane itself moves
//     metric and position and records the destination; the K/V is stored by the next fused attention step;
//   * heads with K/V work (a move from the cache or resolved in place, a parked token to put back) are then
//     handled by the whole warp one after the other: ~32 registers, no 16-token gather;
//   * cnt > 1 heads belong to execute_moves_b16_d128_chunk_kernel.
// FULLROW: the destination V rows are read, patched and written back as whole 32 B sectors instead
// of 128 two-byte partial-sector stores (same DRAM traffic -- the partial stores make L2 fetch the
// sectors -- but full-sector transactions).
// -----------------------------------------------------------------------------
constexpr int kSingleWarps = 8;

template <bool PEND, bool FULLROW>
__global__ void __launch_bounds__(kSingleWarps * kWarp) execute_moves_light_kernel(
    uint16_t* __restrict__ k_cache, uint16_t* __restrict__ v_cache, float* __restrict__ kv_metrics,
    int* __restrict__ kv_position, const int2* __restrict__ moves, const int* __restrict__ move_count,
    const int* __restrict__ move_offsets, int total_heads, const uint16_t* __restrict__ pend_kv,
    const int64_t* __restrict__ pend_slot, int64_t* __restrict__ pend_dst, int num_seqs, int num_layers,
    int num_kv_heads) {
  constexpr int B = 16, D = 128;
  constexpr int kBlockElems = B * D;
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int base = (blockIdx.x * kSingleWarps + wid) * kWarp;
  if (base >= total_heads) return;
  // ---- the lane's own head ----
  const int g = base + lane;
  int my_cnt = 2, my_dst = -1, my_src = -1, my_pslot = -1;
  int64_t my_pidx = 0;
  if (g < total_heads) {
    my_cnt = move_count[g];
    if (my_cnt <= 1) {
      if (PEND) {
        const int sq = g / (num_layers * num_kv_heads);
        const int ly = (g / num_kv_heads) % num_layers;
        my_pidx = ((int64_t)ly * num_seqs + sq) * num_kv_heads + (g % num_kv_heads);
        my_pslot = (int)pend_slot[my_pidx];
      }
      if (my_cnt == 1) {
        const int2 m = moves[move_offsets[g]];
        my_dst = m.x;
        my_src = m.y;
      }
    }
  }
  bool work = false;  // K/V to copy with the whole warp
  if (my_cnt <= 1) {
    if (PEND && pend_dst != nullptr && my_cnt == 1 && my_src == my_pslot) {
      // deferred move: metric and position go now (the next selection reads them), the K/V stays parked and
      // the next fused attention step stores it while it has the destination block staged
      kv_metrics[my_dst] = kv_metrics[my_src];
      kv_position[my_dst] = kv_position[my_src];
      pend_dst[my_pidx] = my_dst;
    } else {
      work = (my_cnt == 1) || (PEND && my_pslot >= 0);
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, work);
  while (todo) {
    const int l0 = __ffs(todo) - 1;
    todo &= todo - 1;
    const int cnt = __shfl_sync(0xffffffffu, my_cnt, l0);
    const int pslot = __shfl_sync(0xffffffffu, my_pslot, l0);
    const int64_t pidx = __shfl_sync(0xffffffffu, my_pidx, l0);
    int dst = __shfl_sync(0xffffffffu, my_dst, l0), src = __shfl_sync(0xffffffffu, my_src, l0);
    const uint16_t* __restrict__ pk = PEND ? pend_kv + pidx * (2 * D) : nullptr;
    if (cnt == 0) dst = src = pslot;  // nothing moved: the parked token goes to its own slot
    const bool from_pend = PEND && src == pslot;
    // a parked token that the head's only move does not consume (its source is an older token)
    const bool flush_too = PEND && pslot >= 0 && !from_pend;
    // ---- loads ----
    uint4 kx = make_uint4(0, 0, 0, 0);
    uint16_t vx[4];
    if (from_pend) {
      if (lane < 16) kx = *reinterpret_cast<const uint4*>(pk + lane * 8);
#pragma unroll
      for (int r = 0; r < 4; ++r) vx[r] = pk[D + lane + 32 * r];
    } else {
      if (lane < 16) kx = ld_stream16(k_cache + (int64_t)(src >> 4) * kBlockElems + lane * (B * 8) + (src & 15) * 8);
      const uint16_t* __restrict__ sp = v_cache + (int64_t)(src >> 4) * kBlockElems + (src & 15) + lane * B;
#pragma unroll
      for (int r = 0; r < 4; ++r) vx[r] = sp[r * 32 * B];
    }
    float mt = 0.f;
    int ps = 0;
    if (cnt == 1) {  // metric and position of a parked token are already at its slot
      if (lane == 16) mt = kv_metrics[src];
      if (lane == 17) ps = kv_position[src];
    }
    uint16_t* __restrict__ vd = v_cache + (int64_t)(dst >> 4) * kBlockElems + lane * B;
    const int o = dst & 15;
    if (FULLROW) {
      uint4 ra[4], rb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        ra[r] = *reinterpret_cast<const uint4*>(vd + r * 32 * B);
        rb[r] = *reinterpret_cast<const uint4*>(vd + r * 32 * B + 8);
      }
      const int wi = o >> 1;
      const uint32_t keep = (o & 1) ? 0xffffu : 0xffff0000u;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t x = (o & 1) ? ((uint32_t)vx[r] << 16) : (uint32_t)vx[r];
        uint32_t w[8] = {ra[r].x, ra[r].y, ra[r].z, ra[r].w, rb[r].x, rb[r].y, rb[r].z, rb[r].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = (i == wi) ? ((w[i] & keep) | x) : w[i];
        *reinterpret_cast<uint4*>(vd + r * 32 * B) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(vd + r * 32 * B + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) vd[r * 32 * B + o] = vx[r];
    }
    if (lane < 16) st_stream16(k_cache + (int64_t)(dst >> 4) * kBlockElems + lane * (B * 8) + o * 8, kx);
    if (cnt == 1) {
      if (lane == 16) kv_metrics[dst] = mt;
      if (lane == 17) kv_position[dst] = ps;
    }
    if (flush_too) {
      uint16_t* kd = k_cache + (int64_t)(pslot >> 4) * kBlockElems + (pslot & 15) * 8;
      uint16_t* pd = v_cache + (int64_t)(pslot >> 4) * kBlockElems + (pslot & 15);
      if (lane < 16) *reinterpret_cast<uint4*>(kd + lane * (B * 8)) = *reinterpret_cast<const uint4*>(pk + lane * 8);
#pragma unroll
      for (int r = 0; r < 4; ++r) pd[(lane + 32 * r) * B] = pk[D + lane + 32 * r];
    }
    __syncwarp();
  }
}

// -----------------------------------------------------------------------------
// generic shapes (the reference's test sizes: D in {1,2,4,128}, x in {1,2,8},
// B in {1,2,4,16}, 1/2/4-byte elements): one warp per move, element-wise copy.
// -----------------------------------------------------------------------------
template <typename T>
__global__ void execute_moves_generic_kernel(T* __restrict__ k_cache, T* __restrict__ v_cache,
                                             float* __restrict__ kv_metrics,
                                             int* __restrict__ kv_position,
                                             const int2* __restrict__ moves,
                                             const int* __restrict__ move_count,
                                             const int* __restrict__ move_offsets,
                                             int total_heads, int D, int x, int B) {
  const int g = blockIdx.x;
  if (g >= total_heads) return;
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int nw = blockDim.x / kWarp;
  const int cnt = move_count[g];
  const int off = move_offsets[g];
  const int64_t block_stride = (int64_t)B * D;
  for (int i = blockIdx.y * nw + wid; i < cnt; i += gridDim.y * nw) {
    const int2 m = moves[off + i];
    const int64_t sb = (int64_t)(m.y / B) * block_stride, db = (int64_t)(m.x / B) * block_stride;
    const int so = m.y % B, dof = m.x % B;
    if (lane == 0) kv_metrics[m.x] = kv_metrics[m.y];
    if (lane == 1 || kWarp == 1) kv_position[m.x] = kv_position[m.y];
    for (int e = lane; e < D; e += kWarp) {
      const int j = e / x, kk = e % x;
      k_cache[db + (int64_t)j * B * x + dof * x + kk] = k_cache[sb + (int64_t)j * B * x + so * x + kk];
      v_cache[db + (int64_t)e * B + dof] = v_cache[sb + (int64_t)e * B + so];
    }
  }
}

}  // namespace kvc

using namespace kvc;

// 0 = register-gather kernels for every shape (default: fastest on the B200),
// 1 = TMA-staged kernel for post-prefill compaction (fewer than 4096 heads)
static int g_moves_mode = -1;
static int moves_mode() {
  if (g_moves_mode < 0) {
    const char* v = getenv("KVC_MOVES_MODE");
    g_moves_mode = v ? atoi(v) : 0;
    if (g_moves_mode < 0 || g_moves_mode > 1) g_moves_mode = 0;
  }
  return g_moves_mode;
}
extern "C" int kvc_set_moves_mode(int mode) {
  KVC_REQUIRE(mode == 0 || mode == 1, KVC_ERR_INVALID, "moves mode must be 0 or 1, got %d", mode);
  g_moves_mode = mode;
  return KVC_OK;
}
extern "C" int kvc_get_moves_mode(void) { return moves_mode(); }

static int execute_moves_impl(
    void* k_cache, void* v_cache, float* kv_metrics, int* kv_position,
    const int* cache_moves_idx, const int* cache_moves_count, const int* evicted_kv_offsets,
    int num_seqs, int num_layers, int num_kv_heads, int64_t num_blocks, int head_size,
    int vec_size, int block_size, int dtype, int blocks_per_head, int threads_per_head,
    const void* pend_kv, const int64_t* pend_slot, int64_t* pend_dst, kvc_stream_t stream) {
  (void)blocks_per_head;
  (void)threads_per_head;  // launch hints in the reference; results never depend on them
  (void)num_blocks;
  // reference dispatch: head size :840-856, vec size :826-838, block size :808-824, dtype :917-927
  KVC_REQUIRE(head_size == 1 || head_size == 2 || head_size == 4 || head_size == 128,
              KVC_ERR_UNSUPPORTED, "Unsupported head size: %d", head_size);
  KVC_REQUIRE(vec_size == 1 || vec_size == 2 || vec_size == 8, KVC_ERR_UNSUPPORTED,
              "Unsupported vec size: %d", vec_size);
  KVC_REQUIRE(block_size == 1 || block_size == 2 || block_size == 4 || block_size == 16,
              KVC_ERR_UNSUPPORTED, "Unsupported block size: %d", block_size);
  KVC_REQUIRE(dtype == KVC_F16 || dtype == KVC_BF16 || dtype == KVC_F32 || dtype == KVC_U8,
              KVC_ERR_UNSUPPORTED, "Unsupported data type: %d", dtype);
  if (num_seqs == 0) return KVC_OK;
  KVC_REQUIRE(k_cache && v_cache && kv_metrics && kv_position && cache_moves_idx &&
                  cache_moves_count && evicted_kv_offsets,
              KVC_ERR_INVALID, "execute_cache_moves: null tensor");
  KVC_REQUIRE((reinterpret_cast<uintptr_t>(cache_moves_idx) & 7) == 0, KVC_ERR_INVALID,
              "execute_cache_moves: cache_moves_idx must be 8-byte aligned");
  KVC_REQUIRE(head_size % vec_size == 0 || head_size < vec_size, KVC_ERR_INVALID,
              "execute_cache_moves: head size %d not divisible by vec size %d", head_size, vec_size);
  const int total_heads = num_seqs * num_layers * num_kv_heads;
  if (total_heads <= 0) return KVC_OK;
  cudaStream_t st = as_stream(stream);
  const int2* mv = reinterpret_cast<const int2*>(cache_moves_idx);
  // enough warps (one per (head, segment stride)) to fill the SMs several times over
  int gy = (sm_count() * 64 + total_heads - 1) / total_heads;
  gy = gy < 1 ? 1 : (gy > 256 ? 256 : gy);
  const bool two_byte = (dtype == KVC_F16 || dtype == KVC_BF16);
  if (two_byte && head_size == 128 && vec_size == 8 && block_size == 16 &&
      (reinterpret_cast<uintptr_t>(k_cache) & 15) == 0 && (reinterpret_cast<uintptr_t>(v_cache) & 15) == 0) {
    // many heads (decode-time compression of a whole batch): a warp per 32 heads; heads with <= 1 move go to
    // the light kernel (deferred moves live there), the others to the run kernel
    const bool split = total_heads >= 4096 || pend_dst != nullptr;
    uint16_t* kc = reinterpret_cast<uint16_t*>(k_cache);
    uint16_t* vc = reinterpret_cast<uint16_t*>(v_cache);
    const uint16_t* pkv = reinterpret_cast<const uint16_t*>(pend_kv);
    const bool pend = pend_slot != nullptr;
    if (split) {
      static const int fullrow = [] { const char* e = getenv("KVC_MOVES_FULLROW"); return e ? atoi(e) : 0; }();
      const int chunks = (total_heads + kWarp - 1) / kWarp;
      const dim3 sgrid((chunks + kSingleWarps - 1) / kSingleWarps);
      const int sthreads = kSingleWarps * kWarp;
#define KVC_LIGHT(PEND_, FULL_)                                                                              \
  execute_moves_light_kernel<PEND_, FULL_><<<sgrid, sthreads, 0, st>>>(                                       \
      kc, vc, kv_metrics, kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, pkv, pend_slot, \
      pend_dst, num_seqs, num_layers, num_kv_heads)
      if (pend) { if (fullrow) KVC_LIGHT(true, true); else KVC_LIGHT(true, false); }
      else { if (fullrow) KVC_LIGHT(false, true); else KVC_LIGHT(false, false); }
#undef KVC_LIGHT
      KVC_CHECK_LAUNCH("execute_moves_light_kernel");
      const dim3 cgrid(min((chunks + kMoveWarps - 1) / kMoveWarps, sm_count() * 8));
      if (pend)
        execute_moves_b16_d128_chunk_kernel<true><<<cgrid, kMoveWarps * kWarp, 0, st>>>(
            kc, vc, kv_metrics, kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, pkv, pend_slot,
            num_seqs, num_layers, num_kv_heads);
      else
        execute_moves_b16_d128_chunk_kernel<false><<<cgrid, kMoveWarps * kWarp, 0, st>>>(
            kc, vc, kv_metrics, kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, nullptr, nullptr,
            0, 0, 0);
    } else if (!pend && moves_mode() == 1) {
      // post-prefill compaction: blocks staged by TMA bulk copies, one persistent CTA of kPrWarps warps per SM
      const size_t smem = (size_t)kPrWarps * kPrWarpBytes;
      static bool attr_set = false;
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(execute_moves_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "execute_moves_staged smem attr: %s", cudaGetErrorString(e));
        attr_set = true;
      }
      execute_moves_staged_kernel<<<sm_count(), kPrWarps * kWarp, smem, st>>>(
          kc, vc, kv_metrics, kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads);
    } else {
      const dim3 grid((total_heads + kMoveWarps - 1) / kMoveWarps, gy);
      if (pend)
        execute_moves_b16_d128_kernel<true><<<grid, kMoveWarps * kWarp, 0, st>>>(
            kc, vc, kv_metrics, kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, pkv, pend_slot,
            num_seqs, num_layers, num_kv_heads);
      else
        execute_moves_b16_d128_kernel<false><<<grid, kMoveWarps * kWarp, 0, st>>>(
            kc, vc, kv_metrics, kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, nullptr, nullptr,
            0, 0, 0);
    }
    KVC_CHECK_LAUNCH("execute_moves_b16_d128_kernel");
    return KVC_OK;
  }
  KVC_REQUIRE(pend_slot == nullptr, KVC_ERR_UNSUPPORTED,
              "execute_cache_moves_deferred: only the production layout (2-byte cache, head size 128, "
              "block size 16, x = 8) supports deferred appends");
  const dim3 grid(total_heads, gy), block(128);
  if (two_byte) {
    execute_moves_generic_kernel<uint16_t><<<grid, block, 0, st>>>(
        reinterpret_cast<uint16_t*>(k_cache), reinterpret_cast<uint16_t*>(v_cache), kv_metrics,
        kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, head_size, vec_size,
        block_size);
  } else if (dtype == KVC_F32) {
    execute_moves_generic_kernel<uint32_t><<<grid, block, 0, st>>>(
        reinterpret_cast<uint32_t*>(k_cache), reinterpret_cast<uint32_t*>(v_cache), kv_metrics,
        kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, head_size, vec_size,
        block_size);
  } else {
    execute_moves_generic_kernel<uint8_t><<<grid, block, 0, st>>>(
        reinterpret_cast<uint8_t*>(k_cache), reinterpret_cast<uint8_t*>(v_cache), kv_metrics,
        kv_position, mv, cache_moves_count, evicted_kv_offsets, total_heads, head_size, vec_size,
        block_size);
  }
  KVC_CHECK_LAUNCH("execute_moves_generic_kernel");
  return KVC_OK;
}

extern "C" int kvc_execute_cache_moves(
    void* k_cache, void* v_cache, float* kv_metrics, int* kv_position,
    const int* cache_moves_idx, const int* cache_moves_count, const int* evicted_kv_offsets,
    int num_seqs, int num_layers, int num_kv_heads, int64_t num_blocks, int head_size,
    int vec_size, int block_size, int dtype, int blocks_per_head, int threads_per_head,
    kvc_stream_t stream) {
  return execute_moves_impl(k_cache, v_cache, kv_metrics, kv_position, cache_moves_idx, cache_moves_count,
                            evicted_kv_offsets, num_seqs, num_layers, num_kv_heads, num_blocks, head_size,
                            vec_size, block_size, dtype, blocks_per_head, threads_per_head, nullptr, nullptr,
                            nullptr, stream);
}

extern "C" int kvc_execute_cache_moves_deferred(
    void* k_cache, void* v_cache, float* kv_metrics, int* kv_position,
    const int* cache_moves_idx, const int* cache_moves_count, const int* evicted_kv_offsets,
    const void* pend_kv, const int64_t* pend_slot, int64_t* pend_dst, int num_seqs, int num_layers,
    int num_kv_heads, int64_t num_blocks, int head_size, int vec_size, int block_size, int dtype,
    kvc_stream_t stream) {
  KVC_REQUIRE(num_seqs == 0 || (pend_kv && pend_slot), KVC_ERR_INVALID,
              "execute_cache_moves_deferred: pend_kv / pend_slot are required");
  KVC_REQUIRE((reinterpret_cast<uintptr_t>(pend_kv) & 15) == 0, KVC_ERR_INVALID,
              "execute_cache_moves_deferred: pend_kv must be 16-byte aligned");
  return execute_moves_impl(k_cache, v_cache, kv_metrics, kv_position, cache_moves_idx, cache_moves_count,
                            evicted_kv_offsets, num_seqs, num_layers, num_kv_heads, num_blocks, head_size,
                            vec_size, block_size, dtype, 1, 16, pend_kv, pend_slot, pend_dst, stream);
}

namespace kvc {
// every parked token to the slot it belongs to: its granted destination if a deferred move is outstanding,
// else the slot batch_append_slots gave it; both markers are cleared
__global__ void __launch_bounds__(256) flush_deferred_kernel(uint16_t* __restrict__ k_cache,
                                                             uint16_t* __restrict__ v_cache,
                                                             const uint16_t* __restrict__ pend_kv,
                                                             int64_t* __restrict__ pend_slot,
                                                             int64_t* __restrict__ pend_dst, int total_heads) {
  constexpr int B = 16, D = 128, kBlockElems = B * D;
  const int lane = threadIdx.x % kWarp;
  const int g = blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
  if (g >= total_heads) return;
  const int64_t pd = pend_dst != nullptr ? pend_dst[g] : -1;
  const int64_t ps = pend_slot[g];
  const int64_t slot = pd >= 0 ? pd : ps;
  __syncwarp();
  if (lane == 0) {
    pend_slot[g] = -1;
    if (pend_dst != nullptr) pend_dst[g] = -1;
  }
  if (slot < 0) return;
  const uint16_t* __restrict__ pk = pend_kv + (int64_t)g * (2 * D);
  uint16_t* kd = k_cache + (slot >> 4) * kBlockElems + (slot & 15) * 8;
  uint16_t* vd = v_cache + (slot >> 4) * kBlockElems + (slot & 15);
  if (lane < 16) *reinterpret_cast<uint4*>(kd + lane * (B * 8)) = *reinterpret_cast<const uint4*>(pk + lane * 8);
#pragma unroll
  for (int r = 0; r < 4; ++r) vd[(lane + 32 * r) * B] = pk[D + lane + 32 * r];
}
}  // namespace kvc

extern "C" int kvc_flush_deferred(void* k_cache, void* v_cache, const void* pend_kv, int64_t* pend_slot,
                                  int64_t* pend_dst, int64_t num_heads, int head_size, int vec_size,
                                  int block_size, int dtype, kvc_stream_t stream) {
  KVC_REQUIRE((dtype == KVC_F16 || dtype == KVC_BF16) && head_size == 128 && vec_size == 8 && block_size == 16,
              KVC_ERR_UNSUPPORTED, "flush_deferred: only the production layout (2-byte cache, head size 128, "
              "block size 16, x = 8) supports deferred appends");
  if (num_heads <= 0) return KVC_OK;
  KVC_REQUIRE(k_cache && v_cache && pend_kv && pend_slot, KVC_ERR_INVALID, "flush_deferred: null tensor");
  kvc::flush_deferred_kernel<<<(unsigned)((num_heads + 7) / 8), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<uint16_t*>(k_cache), reinterpret_cast<uint16_t*>(v_cache),
      reinterpret_cast<const uint16_t*>(pend_kv), pend_slot, pend_dst, (int)num_heads);
  KVC_CHECK_LAUNCH("flush_deferred_kernel");
  return KVC_OK;
}
