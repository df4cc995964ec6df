// K7/K8: KV-Compress paged attention (decode, one query token per sequence) over
// per-(sequence, kv head) block tables and context lengths, with per-KV softmax
// probabilities recorded into the metrics buffer.
// reference: csrc/attention/kvcompress_attention_kernels.cu:95-451 (kernel body),
//            :453-523 (v1/v2 wrappers), :525-652 (v2 reduce), launchers :683-978.
//
// What is different from the reference kernel (a vLLM v0.3.3 PagedAttention derivative):
//  * one CTA per (sequence, KV head) serves ALL Q query heads of the GQA group, so a
//    head's K/V blocks are read from HBM once instead of Q times (reference grid is
//    (num_query_heads, num_seqs), :742);
//  * QK^T and PV run on the tensor cores with the KV tokens on the MMA M dimension
//    (S^T[16 tok x 8 q] = K[16 x 128] Q^T, O^T[128 d x 8 q] = V^T[128 x 16 tok] P^T), so
//    the GQA group's 4-8 query rows sit on N=8 and nothing is padded to M=16+;
//    fragments are loaded straight from the paged layout with 16 B / 8 B vector loads
//    (K block [16 j][16 t][8] -> one 16 B chunk per lane; V block [128 d][16 t] -> four
//    tokens of one row per lane) by permuting the MMA k index consistently on A and B;
//  * online softmax per warp (a warp owns whole blocks), merged across the 4 warps at
//    the end; no [max_context_len] fp32 logits array in shared memory, so shared memory
//    does not grow with the *uncompressed* sequence length (reference :733-739);
//  * recorded probabilities: raw logits are parked in kv_metric_out during the sweep
//    and rewritten in place as __expf(logit - max) * 1/(sum + 1e-6) (reference :288-313)
//    once max/sum are final.
// Arithmetic: fp16/bf16 operands, fp32 accumulate, probabilities cast to the cache dtype
// for PV as in the reference (:356-357).  Parity budget: atol 1e-3.
//
// v1 = one CTA sweeps the whole context.  v2 = 512-token partitions + reduce kernel,
// same buffers as the reference.  fused = v1 + sum_q p^2 accumulated into kv_metrics.
#include <stdlib.h>

#include <type_traits>

#include "kvc_common.cuh"

namespace kvc {

constexpr int kAttnWarps = 4;
constexpr int kAttnThreads = kAttnWarps * kWarp;
constexpr int kD = 128;
constexpr int kB = 16;
constexpr int kPartition = 512;  // reference PARTITION_SIZE, :866
constexpr int kOStride = 132;    // padded d-stride of the cross-warp merge buffer (bank-conflict free)

// kModeFusedWriteMove = kModeFusedWrite + deferred moves applied from pend_dst (its own instantiation: the
// extra code must not cost the plain fused step anything)
enum AttnMode { kModeV1 = 0, kModeV2 = 1, kModeFused = 2, kModeFusedWrite = 3, kModeFusedWriteMove = 4 };

struct AttnParams {
  void* out;
  float* metric;  // kv_metric_out [NB,B,Q] (v1/v2) or kv_metrics [NB,B] (fused)
  float* exp_sums;
  float* max_logits;
  void* tmp_out;
  const void* q;
  const void* k_cache;
  const void* v_cache;
  const int* block_tables;
  const int* context_lens;
  const int* kv_position;
  const int* last_position;
  const int* kv_metric_buffer_len;
  const float* alibi_slopes;
  int num_heads, num_kv_heads, Q;
  float scale;
  int max_num_blocks_per_seq;
  int64_t q_stride, kv_block_stride;
  int record;
  int max_partitions;
  int num_seqs;
  // fused cache write (kModeFusedWrite): the step's new token per (seq, kv head)
  const void* new_k;            // [S, H, D] token stride new_k_stride (elements)
  const void* new_v;
  int64_t new_k_stride, new_v_stride;
  const int64_t* slot_mapping;  // [S, H]; < 0 = no new token for this head
  const float* head_bias;       // [H]
  // deferred append (kModeFusedWrite, optional): a new token that opens a fresh block is parked in
  // pend_kv [S, H, 2 * D] (K | V, compact) with its slot in pend_slot [S, H] instead of being stored to the
  // cache; kvc_execute_cache_moves_deferred of the next step puts it where it finally belongs
  uint16_t* pend_kv;
  int64_t* pend_slot;
  // deferred move (optional, with pend_kv): pend_dst [S, H] >= 0 = slot where the token parked by the PREVIOUS
  // step now belongs (set by kvc_execute_cache_moves_deferred); applied while that block is staged
  int64_t* pend_dst;
#ifdef KVC_ATTN_TRACE
  unsigned long long* trace;  // [items][16] timestamps (tools/attn_trace.py; diagnostic builds only)
#endif
};

#ifdef KVC_ATTN_TRACE
__device__ __forceinline__ unsigned long long kvc_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define KVC_TR(i, v) do { if (p.trace != nullptr && lane == 0) p.trace[(size_t)item * 16 + (i)] = (unsigned long long)(v); } while (0)
#else
#define KVC_TR(i, v) do { } while (0)
#endif

template <bool BF16>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (BF16) {
    asm(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

template <bool BF16>
__device__ __forceinline__ uint16_t to_bits16(float x) {
  if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(x));
  else return __half_as_ushort(__float2half_rn(x));
}
template <bool BF16>
__device__ __forceinline__ float from_bits16(uint16_t x) {
  if constexpr (BF16) return __bfloat162float(__ushort_as_bfloat16(x));
  else return __half2float(__ushort_as_half(x));
}

// -----------------------------------------------------------------------------
// main kernel.  grid (num_kv_heads, num_seqs, partitions), 128 threads.
// NT = number of 8-wide query-head tiles (Q <= 8*NT).
// -----------------------------------------------------------------------------
template <bool BF16, int NT, int MODE>
__global__ void __launch_bounds__(kAttnThreads) paged_attention_kernel(AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int h = blockIdx.x, s = blockIdx.y, part = blockIdx.z;
  const int H = p.num_kv_heads, Q = p.Q;
  const int tid = threadIdx.x, wid = tid / kWarp, lane = tid % kWarp;
  const int g = lane >> 2, c = lane & 3;
  constexpr int NQ = NT * 8;

  // shared memory carve-up
  float* so = reinterpret_cast<float*>(smem_raw);                     // [4][NQ][kOStride]
  float* sm_m = so + kAttnWarps * NQ * kOStride;                       // [4][NQ]
  float* sm_l = sm_m + kAttnWarps * NQ;                                // [4][NQ]
  float* sm_M = sm_l + kAttnWarps * NQ;                                // [NQ] final max
  float* sm_inv = sm_M + NQ;                                           // [NQ] 1/(sum+1e-6)
  uint16_t* ps = reinterpret_cast<uint16_t*>(sm_inv + NQ);             // [4][NT][8][16]
  float* sm_logits = reinterpret_cast<float*>(ps + kAttnWarps * NT * 8 * 16);  // fused: [ntok][Q]

  const int seq_head = s * H + h;
  const int ctx = p.context_lens[seq_head];
  const int nblk = (ctx + kB - 1) / kB;
  int blk_begin = 0, blk_end = nblk;
  if (MODE == kModeV2) {
    blk_begin = part * (kPartition / kB);
    if (blk_begin >= nblk) return;  // reference :139-142
    blk_end = min(nblk, blk_begin + kPartition / kB);
  }
  const int tok_begin = blk_begin * kB;
  const int tok_end = min(ctx, blk_end * kB);
  const int* __restrict__ bt = p.block_tables + (int64_t)seq_head * p.max_num_blocks_per_seq;
  const int max_pos = p.last_position[s] - p.kv_metric_buffer_len[s];  // reference :124
  const uint16_t* __restrict__ kc = reinterpret_cast<const uint16_t*>(p.k_cache);
  const uint16_t* __restrict__ vc = reinterpret_cast<const uint16_t*>(p.v_cache);

  if (MODE != kModeV2 && ctx <= 0) {  // empty head: reference leaves exp_sum = 0 -> out = 0
    for (int i = tid; i < Q * kD; i += kAttnThreads) {
      const int q = i / kD, d = i % kD;
      reinterpret_cast<uint16_t*>(p.out)[((int64_t)s * p.num_heads + h * Q + q) * kD + d] = 0;
    }
    return;
  }

  // ---- Q fragments (B operand of S^T = K Q^T), k index permuted to match the K loads ----
  uint4 qf[NT][4];
  float slope[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int ql = nt * 8 + g;
    const uint16_t* qp = reinterpret_cast<const uint16_t*>(p.q) + (int64_t)s * p.q_stride +
                         (int64_t)(h * Q + ql) * kD;
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4)
      qf[nt][s4] = (ql < Q) ? *reinterpret_cast<const uint4*>(qp + (4 * s4 + c) * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int qc = nt * 8 + 2 * c + i;
      slope[nt][i] = (p.alibi_slopes != nullptr && qc < Q) ? p.alibi_slopes[h * Q + qc] : 0.f;
    }
  }

  float m_run[NT][2], l_run[NT][2];
  float oacc[NT][8][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    m_run[nt][0] = m_run[nt][1] = -INFINITY;
    l_run[nt][0] = l_run[nt][1] = 0.f;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) oacc[nt][dt][i] = 0.f;
  }
  uint16_t* psw = ps + wid * NT * 8 * 16;

  for (int b = blk_begin + wid; b < blk_end; b += kAttnWarps) {
    const int64_t phys = bt[b];
    const uint16_t* __restrict__ kb_ = kc + phys * p.kv_block_stride;
    const uint16_t* __restrict__ vb_ = vc + phys * p.kv_block_stride;
    // K: chunk j = 4*s4 + c of tokens g and g+8 (16 B each)
    uint4 ka[4], kb[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      ka[s4] = ld_stream16(kb_ + (4 * s4 + c) * (kB * 8) + g * 8);
      kb[s4] = ld_stream16(kb_ + (4 * s4 + c) * (kB * 8) + (g + 8) * 8);
    }
    // V: rows d = 16*dt + g and + 8, tokens 4c..4c+3 (8 B each)
    uint2 va[8], vb[8];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      va[dt] = ld_stream8(vb_ + (16 * dt + g) * kB + c * 4);
      vb[dt] = ld_stream8(vb_ + (16 * dt + g + 8) * kB + c * 4);
    }
    const int tok0 = b * kB + g, tok1 = tok0 + 8;
    const int64_t slot0 = phys * kB + g, slot1 = slot0 + 8;
    bool rec0 = false, rec1 = false;
    if (MODE != kModeFused && p.record) {
      rec0 = tok0 < ctx && p.kv_position[slot0] <= max_pos;  // reference :309
      rec1 = tok1 < ctx && p.kv_position[slot1] <= max_pos;
    }

    // ---- S^T = K Q^T ----
    float sacc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        mma16816<BF16>(sacc[nt], ka[s4].x, kb[s4].x, ka[s4].y, kb[s4].y, qf[nt][s4].x, qf[nt][s4].y);
        mma16816<BF16>(sacc[nt], ka[s4].z, kb[s4].z, ka[s4].w, kb[s4].w, qf[nt][s4].z, qf[nt][s4].w);
      }
    }

    // ---- logits, online softmax ----
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float pexp[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int qc = nt * 8 + 2 * c + i;
        float s0 = sacc[nt][i] * p.scale, s1 = sacc[nt][2 + i] * p.scale;
        if (slope[nt][i] != 0.f) {  // reference :253
          s0 += slope[nt][i] * (float)(tok0 - ctx + 1);
          s1 += slope[nt][i] * (float)(tok1 - ctx + 1);
        }
        if (tok0 >= ctx) s0 = -INFINITY;
        if (tok1 >= ctx) s1 = -INFINITY;
        if (qc < Q) {
          if (MODE == kModeFused) {
            if (tok0 < ctx) sm_logits[(tok0 - tok_begin) * Q + qc] = s0;
            if (tok1 < ctx) sm_logits[(tok1 - tok_begin) * Q + qc] = s1;
          } else {
            if (rec0) p.metric[slot0 * Q + qc] = s0;  // parked raw logit, finalised below
            if (rec1) p.metric[slot1 * Q + qc] = s1;
          }
        }
        float mx = fmaxf(s0, s1);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        const float m_new = fmaxf(m_run[nt][i], mx);  // finite: every block has >= 1 valid token
        const float alpha = __expf(m_run[nt][i] - m_new);
        const float p0 = __expf(s0 - m_new), p1 = __expf(s1 - m_new);
        float rs = p0 + p1;
        rs += __shfl_xor_sync(0xffffffffu, rs, 4);
        rs += __shfl_xor_sync(0xffffffffu, rs, 8);
        rs += __shfl_xor_sync(0xffffffffu, rs, 16);
        l_run[nt][i] = l_run[nt][i] * alpha + rs;
        m_run[nt][i] = m_new;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          oacc[nt][dt][i] *= alpha;
          oacc[nt][dt][2 + i] *= alpha;
        }
        pexp[i] = p0;
        pexp[2 + i] = p1;
      }
      // transpose P through shared memory: psw[nt][q][t]
      uint16_t* pq = psw + nt * 8 * 16;
      pq[(2 * c) * 16 + g] = to_bits16<BF16>(pexp[0]);
      pq[(2 * c + 1) * 16 + g] = to_bits16<BF16>(pexp[1]);
      pq[(2 * c) * 16 + g + 8] = to_bits16<BF16>(pexp[2]);
      pq[(2 * c + 1) * 16 + g + 8] = to_bits16<BF16>(pexp[3]);
    }
    __syncwarp();

    // tokens beyond the context may hold NaNs: zero them (reference :370-379)
    if (b == nblk - 1 && (ctx & (kB - 1)) != 0) {
      const int vcnt = ctx - b * kB;  // valid tokens in this block
      const int t0 = 4 * c;
      const uint32_t mlo = (t0 < vcnt ? 0xffffu : 0u) | (t0 + 1 < vcnt ? 0xffff0000u : 0u);
      const uint32_t mhi = (t0 + 2 < vcnt ? 0xffffu : 0u) | (t0 + 3 < vcnt ? 0xffff0000u : 0u);
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        va[dt].x &= mlo; va[dt].y &= mhi;
        vb[dt].x &= mlo; vb[dt].y &= mhi;
      }
    }

    // ---- O^T += V^T P^T ----
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint2 pb = *reinterpret_cast<const uint2*>(psw + nt * 8 * 16 + g * 16 + 4 * c);
#pragma unroll
      for (int dt = 0; dt < 8; ++dt)
        mma16816<BF16>(oacc[nt][dt], va[dt].x, vb[dt].x, va[dt].y, vb[dt].y, pb.x, pb.y);
    }
    __syncwarp();
  }

  // ---- merge the 4 warps ----
  if (g == 0) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        sm_m[wid * NQ + nt * 8 + 2 * c + i] = m_run[nt][i];
        sm_l[wid * NQ + nt * 8 + 2 * c + i] = l_run[nt][i];
      }
  }
  __syncthreads();
  if (tid < NQ) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, sm_m[w * NQ + tid]);
    float L = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) L += sm_l[w * NQ + tid] * __expf(sm_m[w * NQ + tid] - M);
    sm_M[tid] = M;
    sm_inv[tid] = __fdividef(1.f, L + 1e-6f);  // reference :298
    if (MODE == kModeV2 && tid < Q) {          // reference :316-326
      const int64_t idx = ((int64_t)s * p.num_heads + h * Q + tid) * p.max_partitions + part;
      p.max_logits[idx] = M;
      p.exp_sums[idx] = L;
    }
  }
  __syncthreads();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int col = nt * 8 + 2 * c + i;
      const float wgt = __expf(m_run[nt][i] - sm_M[col]);
      float* dst = so + (wid * NQ + col) * kOStride;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        dst[16 * dt + g] = oacc[nt][dt][i] * wgt;
        dst[16 * dt + g + 8] = oacc[nt][dt][2 + i] * wgt;
      }
    }
  __syncthreads();
  {
    const int d = tid;  // 128 threads <-> 128 head dims
    for (int q = 0; q < Q; ++q) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) acc += so[(w * NQ + q) * kOStride + d];
      acc *= sm_inv[q];
      const int qh = h * Q + q;
      if (MODE == kModeV2) {
        reinterpret_cast<uint16_t*>(p.tmp_out)[(((int64_t)s * p.num_heads + qh) * p.max_partitions + part) * kD + d] =
            to_bits16<BF16>(acc);
      } else {
        reinterpret_cast<uint16_t*>(p.out)[((int64_t)s * p.num_heads + qh) * kD + d] = to_bits16<BF16>(acc);
      }
    }
  }

  // ---- finalise the recorded probabilities ----
  if (MODE == kModeV1 && p.record) {
    const int n = (tok_end - tok_begin) * Q;
    for (int i = tid; i < n; i += kAttnThreads) {
      const int t = tok_begin + i / Q, q = i % Q;
      const int64_t slot = (int64_t)bt[t / kB] * kB + t % kB;
      if (p.kv_position[slot] <= max_pos) {
        const float lg = p.metric[slot * Q + q];
        p.metric[slot * Q + q] = __expf(lg - sm_M[q]) * sm_inv[q];  // reference :294-313
      }
    }
  }
  if (MODE == kModeFused) {
    // metrics[slot] += sum_q p^2   (vllm/kvcompress/metrics.py:429-439, use_l2)
    for (int t = tok_begin + tid; t < tok_end; t += kAttnThreads) {
      const int64_t slot = (int64_t)bt[t / kB] * kB + t % kB;
      if (p.kv_position[slot] <= max_pos) {
        float acc = 0.f;
        for (int q = 0; q < Q; ++q) {
          const float pr = __expf(sm_logits[(t - tok_begin) * Q + q] - sm_M[q]) * sm_inv[q];
          acc += pr * pr;
        }
        p.metric[slot] += acc;
      }
    }
  }
}


// -----------------------------------------------------------------------------
// Streaming kernel for the compressed regime (many (seq, head) items, short contexts):
// ONE WARP owns a whole (sequence, kv head) item, so there is no cross-warp merge, no CTA
// barrier and no partial-output traffic.  K/V blocks are staged by TMA bulk copies
// (cp.async.bulk, 4 KB K + 4 KB V per block, completion on a per-stage mbarrier) into a
// private ring of NSTAGE stages per warp; lane 0 runs the ring NSTAGE-1 blocks ahead of the
// math, so each warp keeps (NSTAGE-1) * 8 KB in flight without holding registers, and the SM
// keeps warps_per_SM * that.  MMA fragments are read from the staged blocks with the same
// vector shapes as the direct-from-global kernel above (16 B per lane for K, 8 B for V).
// grid = ceil(S * H / 4) CTAs of 4 independent warps.
// -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// CTA = kConsumers consumer warps (one (seq, head) item each) + 1 producer warp.
constexpr int kConsumers = 7;
constexpr int kStreamThreads = (kConsumers + 1) * kWarp;
constexpr int kUnitBytes = kB * kD;  // 2 KB: half a K block (8 of 16 chunks) or half a V block (64 rows)
constexpr int kNUnit = 6;            // ring slots per consumer: 12 KB, 5 units (10 KB) of look-ahead
// units of block b in consumption order: K lo, K hi, V lo, V hi

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Warp-specialised streaming kernel.  The producer warp's lane w feeds consumer w: it walks the
// item's block table and issues one cp.async.bulk per 2 KB unit into the consumer's private ring
// as soon as the slot's "empty" mbarrier says its fragments are in registers.  A consumer warp does
// no address generation for global memory at all: wait(full) -> LDS fragments -> arrive(empty) ->
// MMA.  No CTA-wide barrier, no cross-warp merge.
template <bool BF16, int NT, int MODE, bool ALIBI>
__global__ void __launch_bounds__(kStreamThreads, 2) paged_attention_stream_kernel(AttnParams p, float* scratch,
                                                                                   int scratch_tokens,
                                                                                   int smem_logit_tokens, int pdl, int pair) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool kWrite = (MODE == kModeFusedWrite || MODE == kModeFusedWriteMove);
  constexpr bool kMove = (MODE == kModeFusedWriteMove);
  // Programmatic dependent launch (kvc_set_attention_pdl, include/kvc_b200.h).  pdl <= 1: every warp waits for
  // the stream's previous kernel (completion + memory flush) before its first global access, and only then
  // lets the NEXT kernel start: nothing launched behind this grid can run before everything in front of it is
  // visible.  pdl == 2: tables / context lengths / cache blocks / parked tokens were not written by the
  // previous kernel (it was a projection kernel, or the previous layer's attention): the producers stream
  // them right away and only the consumers wait, before q / new k / new v.  pdl == 3: those inputs are
  // independent too; the wait sits at the exit.  All no-ops without the launch attribute.
  if (pdl <= 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int wid = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  constexpr int kPsBytes = NT * 8 * 16 * 2;
  const int H = p.num_kv_heads, Q = p.Q;
  // per-warp smem: ring | P^T | full/empty barriers | (fused, short contexts) raw logits [Q][smem_logit_tokens]
  const int kPerWarp = kNUnit * kUnitBytes + kPsBytes + 2 * kNUnit * 8 + 32 + smem_logit_tokens * Q * 4;
  const int num_items = p.num_kv_heads * p.num_seqs;
  auto ring_of = [&](int w) { return smem_raw + (size_t)w * kPerWarp; };
  auto full_of = [&](int w) { return smem_u32(ring_of(w) + kNUnit * kUnitBytes + kPsBytes); };
  auto empty_of = [&](int w) { return full_of(w) + kNUnit * 8; };
  auto patched_of = [&](int w) { return full_of(w) + 2 * kNUnit * 8; };  // deferred move stored (kMove)

  if (wid < kConsumers && lane == 0) {
#pragma unroll
    for (int i = 0; i < kNUnit; ++i) {
      mbar_init(full_of(wid) + 8 * i, 1);
      mbar_init(empty_of(wid) + 8 * i, 1);
    }
    if (kMove) mbar_init(patched_of(wid), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // =========================== producer warp ===========================
  if (wid == kConsumers && pair) {
    // PAIR mode: whole 4 KB blocks per bulk copy (K lo + hi, V lo + hi are adjacent in global memory and in the
    // ring): half as many copies and barrier transactions; three producer lanes per ring, lane j owns slot pair j
    const int w = lane / 3, j = lane - 3 * w;
    const int item = blockIdx.x * kConsumers + w;
    if (w >= kConsumers || item >= num_items) return;
    const int* __restrict__ bt = p.block_tables + (int64_t)item * p.max_num_blocks_per_seq;
    const int bt_spec0 = bt[0], bt_spec1 = (p.max_num_blocks_per_seq > 1) ? bt[1] : 0;
    const int ctx = p.context_lens[item];
    const int64_t pdp = kMove ? p.pend_dst[item] : -1;
    const int pd_phys = pdp >= 0 ? (int)(pdp >> 4) : -1;
    bool patched = pd_phys < 0;
    const int nblk = (ctx + kB - 1) / kB;
    if (nblk == 0) return;
    const uint16_t* __restrict__ kc = reinterpret_cast<const uint16_t*>(p.k_cache);
    const uint16_t* __restrict__ vc = reinterpret_cast<const uint16_t*>(p.v_cache);
    const uint32_t ring = smem_u32(ring_of(w)), full = full_of(w), empty = empty_of(w);
    const unsigned lanes = __activemask();
    const int npairs = 2 * nblk;
    bool skip_last = false;
    if (kWrite) {
      const int64_t nslot = p.slot_mapping[item];
      skip_last = nslot >= 0 && (nslot & (kB - 1)) == 0 && ctx == (nblk - 1) * kB + 1;
    }
    // pair order: K(0); then per block b: K(b+1) (if any), V(b)
    auto decode = [&](int q, int& isv, int& blk) {
      if (q == 0) { isv = 0; blk = 0; return; }
      const int t = q - 1, b = t >> 1, r = t & 1;
      if (b < nblk - 1) { isv = r; blk = isv ? b : b + 1; }
      else { isv = 1; blk = nblk - 1; }
    };
    int q = j;
    int isv = 0, blk = 0, phys = 0;
    if (q < npairs) { decode(q, isv, blk); phys = (blk == 0) ? bt_spec0 : bt_spec1; }  // first pairs: blk <= 1
    constexpr int kNPair = kNUnit / 2;
    while (__any_sync(lanes, q < npairs)) {
      if (q < npairs) {
        const int sp = q % kNPair, round = q / kNPair;
        bool ready = true;
        if (round > 0) {  // the slot pair is free once its second unit has been released
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(empty + 8 * (2 * sp + 1)), "r"((uint32_t)((round - 1) & 1)) : "memory");
          ready = ok != 0;
        }
        if (kMove && ready && !patched && phys == pd_phys) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(patched_of(w)), "r"(0u) : "memory");
          patched = ok != 0;
          ready = patched;
        }
        if (ready) {
          if (skip_last && blk == nblk - 1) {
            mbar_arrive(full + 8 * (2 * sp));
          } else {
            const uint16_t* src = (isv ? vc : kc) + (int64_t)phys * p.kv_block_stride;
            mbar_expect_tx(full + 8 * (2 * sp), 2 * kUnitBytes);
            bulk_g2s(ring + 2 * sp * kUnitBytes, src, 2 * kUnitBytes, full + 8 * (2 * sp));
          }
          q += kNPair;
          if (q < npairs) { decode(q, isv, blk); phys = bt[blk]; }
        }
      }
    }
    return;
  }
  if (wid == kConsumers) {
    // four producer lanes per consumer ring: lane (w, j) issues units u = j, j+4, j+8, ... of ring w, so up
    // to 28 bulk copies go out per polling pass and no lane ever blocks (test_wait, not try_wait).
    const int w = lane >> 2, j = lane & 3;
    const int item = blockIdx.x * kConsumers + w;
    if (w >= kConsumers || item >= num_items) return;
    const int* __restrict__ bt = p.block_tables + (int64_t)item * p.max_num_blocks_per_seq;
    const int bt_spec0 = bt[0], bt_spec1 = (p.max_num_blocks_per_seq > 1) ? bt[1] : 0;  // speculative, with ctx
    const int ctx = p.context_lens[item];  // item == s * H + h
    // deferred move: the consumer stores the parked token to the cache in its prologue; units of the
    // destination block are not fetched before that store is visible (patched barrier)
    const int64_t pdp = kMove ? p.pend_dst[item] : -1;
    const int pd_phys = pdp >= 0 ? (int)(pdp >> 4) : -1;
    bool patched = pd_phys < 0;
    const int nblk = (ctx + kB - 1) / kB;
    if (nblk == 0) return;
    const uint16_t* __restrict__ kc = reinterpret_cast<const uint16_t*>(p.k_cache);
    const uint16_t* __restrict__ vc = reinterpret_cast<const uint16_t*>(p.v_cache);
    const uint32_t ring = smem_u32(ring_of(w)), full = full_of(w), empty = empty_of(w);
    const unsigned lanes = __activemask();
    const int nunits = 4 * nblk;
    // Fused cache write, new token opens a fresh block (the steady state of continual compression: every
    // head sits on a block boundary after each compression): that block's only valid token is the one this
    // step writes, and the consumer patches it into the staged copy from registers -- nothing to fetch.
    // Its 4 units are released with a plain arrive; whatever the ring slot still holds is masked like any
    // other slot beyond the context length.
    bool skip_last = false;
    if (kWrite) {
      const int64_t nslot = p.slot_mapping[item];
      skip_last = nslot >= 0 && (nslot & (kB - 1)) == 0 && ctx == (nblk - 1) * kB + 1;
    }
    // Unit order = the consumer's software pipeline: K(0) lo/hi; then per block b: K(b+1) lo/hi (if any), V(b) lo/hi.
    auto decode = [&](int u, int& isv, int& hf, int& blk) {
      if (u < 2) { isv = 0; hf = u; blk = 0; return; }
      const int t = u - 2, b = t >> 2, r = t & 3;
      if (b < nblk - 1) { isv = r >> 1; hf = r & 1; blk = isv ? b : b + 1; }
      else { isv = 1; hf = (t - 4 * (nblk - 1)) & 1; blk = nblk - 1; }
    };
    int u = j;
    int isv = 0, hf = 0, blk = 0, phys = 0;
    if (u < nunits) { decode(u, isv, hf, blk); phys = (blk == 0) ? bt_spec0 : bt_spec1; }  // first unit: blk <= 1
    while (__any_sync(lanes, u < nunits)) {
      if (u < nunits) {
        const int slot = u % kNUnit, round = u / kNUnit;
        bool ready = true;
        if (round > 0) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(empty + 8 * slot), "r"((uint32_t)((round - 1) & 1)) : "memory");
          ready = ok != 0;
        }
        if (kMove && ready && !patched && phys == pd_phys) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(patched_of(w)), "r"(0u) : "memory");
          patched = ok != 0;
          ready = patched;
        }
        if (ready) {
          if (skip_last && blk == nblk - 1) {
            mbar_arrive(full + 8 * slot);
          } else {
            const uint16_t* src = (isv ? vc : kc) + (int64_t)phys * p.kv_block_stride + hf * (kUnitBytes / 2);
            mbar_expect_tx(full + 8 * slot, kUnitBytes);
            bulk_g2s(ring + slot * kUnitBytes, src, kUnitBytes, full + 8 * slot);
          }
          u += 4;
          if (u < nunits) { decode(u, isv, hf, blk); phys = bt[blk]; }  // table entry of the lane's next unit
        }
      }
    }
    return;
  }

  // =========================== consumer warps ===========================
  const int g = lane >> 2, c = lane & 3;
  const int item = blockIdx.x * kConsumers + wid;
  if (item >= num_items) return;
  KVC_TR(0, kvc_globaltimer());
  KVC_TR(1, clock64());
  const int s = item / H, h = item - s * H;
  // Everything the warp needs from tables that no kernel of this step's attention chain writes is requested at
  // once, so that the warp pays one memory round trip here instead of one per dependent load later: context
  // length, the first 32 block-table entries (lane b holds entry b: the sweep and the metric epilogue take them
  // by shuffle instead of a dependent load per block), the step's slot for the new token, the recording bound.
  const int* __restrict__ bt = p.block_tables + (int64_t)item * p.max_num_blocks_per_seq;
  const int my_bt = (lane < p.max_num_blocks_per_seq) ? bt[lane] : 0;
  const int ctx = p.context_lens[item];
  // destination of a deferred move (see below); loaded with the context length so the two latencies overlap
  const int64_t pd = kMove ? p.pend_dst[item] : -1;
  int64_t nslot_early = -1;
  if (kWrite) nslot_early = p.slot_mapping[item];
  const int max_pos = p.last_position[s] - p.kv_metric_buffer_len[s];
  auto phys_of = [&](int b) -> int64_t {  // b is warp-uniform
    return b < kWarp ? (int64_t)__shfl_sync(0xffffffffu, my_bt, b) : (int64_t)bt[b];
  };
  const int nblk = (ctx + kB - 1) / kB;
  uint16_t* outp = reinterpret_cast<uint16_t*>(p.out) + ((int64_t)s * p.num_heads + h * Q) * kD;
  if (ctx <= 0) {
    for (int i = lane; i < Q * kD; i += kWarp) outp[i] = 0;
    if (kWrite && p.pend_slot != nullptr && lane == 0) p.pend_slot[item] = -1;
    return;
  }
  unsigned char* ring = ring_of(wid);
  // Deferred move: the token this head parked last step was granted an evicted slot by the compression that ran
  // since (execute_cache_moves_deferred left only its destination in pend_dst).  It is stored to the cache
  // here, while the warp would otherwise just wait for its first units; the producer holds the destination
  // block's units back until the store is visible, so this step's attention already reads the moved token and
  // the sweep below is exactly the plain fused step.  The 144 partial-sector stores make L2 fetch the block's
  // sectors -- the very sectors the bulk copy of that block needs next, so DRAM sees them once -- and only the
  // write-back of the block is extra traffic.
  if (kMove) {
    if (pd >= 0) {
      const uint16_t* __restrict__ pk = p.pend_kv + (int64_t)item * (2 * kD);
      uint4 kx = make_uint4(0, 0, 0, 0);
      if (lane < 16) kx = *reinterpret_cast<const uint4*>(pk + lane * 8);
      uint16_t vx[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) vx[r] = pk[kD + lane + 32 * r];
      const int o = (int)(pd & (kB - 1));
      uint16_t* kd = reinterpret_cast<uint16_t*>(const_cast<void*>(p.k_cache)) + (pd >> 4) * p.kv_block_stride;
      uint16_t* vd = reinterpret_cast<uint16_t*>(const_cast<void*>(p.v_cache)) + (pd >> 4) * p.kv_block_stride;
      if (lane < 16) *reinterpret_cast<uint4*>(kd + lane * (kB * 8) + o * 8) = kx;
#pragma unroll
      for (int r = 0; r < 4; ++r) vd[(lane + 32 * r) * kB + o] = vx[r];
      asm volatile("fence.proxy.async.global;" ::: "memory");  // generic stores before the async-proxy read of the block
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(patched_of(wid));
  }
  uint16_t* psw = reinterpret_cast<uint16_t*>(ring + kNUnit * kUnitBytes);
  const uint32_t full = full_of(wid), empty = empty_of(wid);
  // fused mode parks raw logits transposed, [q][token], in an L2-resident scratch row of this item;
  // v1 mode parks them in kv_metric_out itself
  float* lg0 = nullptr;
  float* lg1 = nullptr;
  constexpr bool kFusedMetrics = (MODE == kModeFused || kWrite);
  // short contexts: logits stay in this warp's shared memory and the recording rule is evaluated
  // block by block during the sweep (a token that fails it gets -inf as its parked logit, i.e. contributes
  // exactly 0), so the epilogue has no dependent global loads; long contexts park them in the global scratch
  // row instead
  const bool smem_logits = kFusedMetrics && smem_logit_tokens > 0;
  const int lg_stride = smem_logits ? smem_logit_tokens : scratch_tokens;
  float* lg_base = nullptr;
  if (kFusedMetrics) {
    float* base = smem_logits ? reinterpret_cast<float*>(ring + kNUnit * kUnitBytes + kPsBytes + 2 * kNUnit * 8 + 32)
                              : scratch + (int64_t)item * Q * scratch_tokens;
    lg_base = base;
    lg0 = base + (2 * c) * lg_stride;  // column q = 2c (nt = 0)
    lg1 = lg0 + lg_stride;
  }

  // fused cache write (kvcompress_reshape_and_cache folded in, csrc/kvcompress_cache_kernels.cu:44-88):
  // the step's new token of this head is stored to the cache for later steps, and patched into the
  // staged copy of the last block below (the TMA read of that block may race with the store).
  // pdl == 2: everything read so far (tables, context lengths, cache blocks of earlier steps) predates the
  // stream's previous kernel by the caller's declaration; the step's q / new k / new v may be its output.
  KVC_TR(2, clock64());
  if (pdl == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
  KVC_TR(3, clock64());
  uint4 newk = make_uint4(0, 0, 0, 0);   // lane j < 16: chunk j of the new K
  uint16_t newv[4] = {0, 0, 0, 0};       // rows d = lane + 32 r of the new V
  int new_o = -1;                        // block offset of the new token, -1 = none
  int64_t new_slot = -1;
  if (kWrite) {
    const int64_t nslot = nslot_early;
    if (nslot >= 0) {
      new_o = (int)(nslot & (kB - 1));
      const uint16_t* nk = reinterpret_cast<const uint16_t*>(p.new_k) + (int64_t)s * p.new_k_stride + h * kD;
      const uint16_t* nv = reinterpret_cast<const uint16_t*>(p.new_v) + (int64_t)s * p.new_v_stride + h * kD;
      if (lane < 16) newk = *reinterpret_cast<const uint4*>(nk + lane * 8);
#pragma unroll
      for (int r = 0; r < 4; ++r) newv[r] = nv[lane + 32 * r];
      new_slot = nslot;  // the stores to the cache are issued after the sweep (nothing here waits on them)
    }
  }

  // Q as the B operand of S^T = K Q^T in the standard m16n8k16 layout: k-step ks covers d = 16 ks .. 16 ks + 15,
  // b0 = Q[g][16 ks + 2c, +1], b1 = Q[g][16 ks + 8 + 2c, +1]
  uint32_t qf[NT][8][2];
  float slope[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int ql = nt * 8 + g;
    const uint16_t* qp = reinterpret_cast<const uint16_t*>(p.q) + (int64_t)s * p.q_stride + (int64_t)(h * Q + ql) * kD;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qf[nt][ks][0] = (ql < Q) ? *reinterpret_cast<const uint32_t*>(qp + 16 * ks + 2 * c) : 0u;
      qf[nt][ks][1] = (ql < Q) ? *reinterpret_cast<const uint32_t*>(qp + 16 * ks + 8 + 2 * c) : 0u;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int qc = nt * 8 + 2 * c + i;
      slope[nt][i] = (ALIBI && qc < Q) ? p.alibi_slopes[h * Q + qc] : 0.f;
    }
  }
  float m_run[NT][2], l_part[NT][2];  // running max per column; per-lane PARTIAL sums (reduced at the end)
  float oacc[NT][8][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    m_run[nt][0] = m_run[nt][1] = -INFINITY;
    l_part[nt][0] = l_part[nt][1] = 0.f;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
      for (int i = 0; i < 4; ++i) oacc[nt][dt][i] = 0.f;
  }
  // per-lane fragment offsets inside a unit (bytes): K chunk c of token g; V row g, tokens 4c..4c+3
  // ldmatrix.x4 row address of this lane inside a K half: matrix m = lane / 8 -> chunk (m >> 1), tokens 8 * (m & 1) + lane % 8
  const int koff = (lane >> 4) * (kB * 16) + (((lane >> 3) & 1) * 8 + (lane & 7)) * 16;
  const int voff = g * (kB * 2) + c * 8;
  int slot = 0;
  uint32_t phase = 0;
  const int full_mask = pair ? ~1 : ~0;  // pair mode: the two units of a block complete on the first one's barrier
  auto next_unit = [&]() -> const unsigned char* {  // wait for the next unit of the ring
    mbar_wait(full + 8 * (slot & full_mask), phase);
    return ring + slot * kUnitBytes;
  };
  auto release_unit = [&]() {  // fragments of the current unit are in registers on every lane
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + 8 * slot);
    if (++slot == kNUnit) { slot = 0; phase ^= 1u; }
  };

  // S^T(kblk) = K Q^T: consume the two K halves of block kblk, issue two independent MMA chains
  // KL (compile time): kblk is the item's last block (the only one the step's new token can be patched into)
  auto consume_k = [&](auto KL, int kblk, float (&sacc)[2][NT][4]) {
    constexpr bool klast = decltype(KL)::value;
    (void)kblk;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const unsigned char* kU = next_unit();
      if (kWrite && klast && new_o >= 0) {  // patch the new token into the staged K half
        if ((lane >> 3) == hf && lane < 16)
          *reinterpret_cast<uint4*>(const_cast<unsigned char*>(kU) + (lane & 7) * (kB * 16) + new_o * 16) = newk;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic write, slot later refilled by TMA
        __syncwarp();
      }
      // four k-steps per half; one conflict-free ldmatrix.x4 per step: (chunk 2j, tok 0-7), (chunk 2j, tok 8-15),
      // (chunk 2j+1, tok 0-7), (chunk 2j+1, tok 8-15) = a0..a3 of the m16n8k16 A fragment
      const uint32_t kaddr = smem_u32(kU) + koff;
      uint32_t ka[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(ka[j][0]), "=r"(ka[j][1]), "=r"(ka[j][2]), "=r"(ka[j][3])
                     : "r"(kaddr + j * (2 * kB * 16)));
      release_unit();
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sacc[hf][nt][0] = sacc[hf][nt][1] = sacc[hf][nt][2] = sacc[hf][nt][3] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma16816<BF16>(sacc[hf][nt], ka[j][0], ka[j][1], ka[j][2], ka[j][3], qf[nt][4 * hf + j][0], qf[nt][4 * hf + j][1]);
      }
    }
  };

  float sacc[2][NT][4];  // logits of the block whose softmax is next
  // The sweep body exists in three compile-time flavours so that the steady-state loop carries none of the
  // last-block work (new-token patches, masking of slots beyond the context, NaN scrubbing): LAST = this is the
  // item's last block, KL = the block whose K is consumed one ahead (b + 1) is the last one.
  auto sweep_block = [&](auto LAST_T, auto KL_T, int b) {
    constexpr bool last = decltype(LAST_T)::value;
    const int tok0 = b * kB + g, tok1 = tok0 + 8;
    bool rec0 = false, rec1 = false;
    int64_t slot0 = 0, slot1 = 0;
    if (MODE == kModeV1 && p.record) {  // independent of the staged data: issued before the waits
      const int64_t phys = phys_of(b);
      slot0 = phys * kB + g;
      slot1 = slot0 + 8;
      rec0 = tok0 < ctx && p.kv_position[slot0] <= max_pos;
      rec1 = tok1 < ctx && p.kv_position[slot1] <= max_pos;
    }
    int pos0 = 0x7fffffff, pos1 = 0x7fffffff;
    if (smem_logits) {  // recording rule of this block's two tokens (loads overlap the math below)
      const int64_t phys = phys_of(b);
      if (tok0 < ctx) pos0 = p.kv_position[phys * kB + g];
      if (tok1 < ctx) pos1 = p.kv_position[phys * kB + g + 8];
    }
    // software pipeline: the next block's QK^T chains are in flight while this block's softmax runs
    float snext[2][NT][4];
    if (!last) consume_k(KL_T, b + 1, snext);
    // ---- logits + online softmax ----
    bool rescale = false;
    float alpha[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float pexp[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float s0 = (sacc[0][nt][i] + sacc[1][nt][i]) * p.scale;
        float s1 = (sacc[0][nt][2 + i] + sacc[1][nt][2 + i]) * p.scale;
        if (ALIBI) {
          s0 += slope[nt][i] * (float)(tok0 - ctx + 1);
          s1 += slope[nt][i] * (float)(tok1 - ctx + 1);
        }
        if (last) {
          if (tok0 >= ctx) s0 = -INFINITY;
          if (tok1 >= ctx) s1 = -INFINITY;
        }
        const int qc = nt * 8 + 2 * c + i;
        if (qc < Q) {
          if (kFusedMetrics) {
            float* lg = (i ? lg1 : lg0) + nt * 8 * lg_stride;
            lg[tok0] = s0;  // slots beyond ctx exist in the scratch row (padded to a block multiple)
            lg[tok1] = s1;
          } else {
            if (rec0) p.metric[slot0 * Q + qc] = s0;
            if (rec1) p.metric[slot1 * Q + qc] = s1;
          }
        }
        float mx = fmaxf(s0, s1);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        const float m_new = fmaxf(m_run[nt][i], mx);
        alpha[nt][i] = __expf(m_run[nt][i] - m_new);
        rescale |= (m_new != m_run[nt][i]);
        const float p0 = __expf(s0 - m_new), p1 = __expf(s1 - m_new);
        l_part[nt][i] = l_part[nt][i] * alpha[nt][i] + (p0 + p1);
        m_run[nt][i] = m_new;
        pexp[i] = p0;
        pexp[2 + i] = p1;
      }
      uint16_t* pq = psw + nt * 8 * 16 + (2 * c) * 16 + g;
      pq[0] = to_bits16<BF16>(pexp[0]);
      pq[16] = to_bits16<BF16>(pexp[1]);
      pq[8] = to_bits16<BF16>(pexp[2]);
      pq[24] = to_bits16<BF16>(pexp[3]);
    }
    // the running max rarely moves after the first blocks: skip the 64 multiplies when it did not
    if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          oacc[nt][dt][0] *= alpha[nt][0];
          oacc[nt][dt][1] *= alpha[nt][1];
          oacc[nt][dt][2] *= alpha[nt][0];
          oacc[nt][dt][3] *= alpha[nt][1];
        }
    }
    // ---- O^T += V^T P^T, one V half (four d-tiles) at a time ----
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const unsigned char* vU = next_unit();
      if (kWrite && last && new_o >= 0) {  // rows 64*hf + lane, 64*hf + 32 + lane of the new V
        uint16_t* vw = reinterpret_cast<uint16_t*>(const_cast<unsigned char*>(vU));
        vw[lane * kB + new_o] = newv[2 * hf];
        vw[(lane + 32) * kB + new_o] = newv[2 * hf + 1];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
      }
      const unsigned char* vS = vU + voff;
      uint2 va[4], vb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // rows 16*(4*hf+j) + g (+8) of the block = rows 16*j + g (+8) of this half
        va[j] = *reinterpret_cast<const uint2*>(vS + j * (16 * kB * 2));
        vb[j] = *reinterpret_cast<const uint2*>(vS + j * (16 * kB * 2) + 8 * kB * 2);
      }
      release_unit();  // its __syncwarp also orders this block's P stores before the P loads below
      if (last && (ctx & (kB - 1)) != 0) {  // tokens beyond the context may hold NaNs (reference :370-379)
        const int vcnt = ctx - b * kB;
        const int t0 = 4 * c;
        const uint32_t mlo = (t0 < vcnt ? 0xffffu : 0u) | (t0 + 1 < vcnt ? 0xffff0000u : 0u);
        const uint32_t mhi = (t0 + 2 < vcnt ? 0xffffu : 0u) | (t0 + 3 < vcnt ? 0xffff0000u : 0u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          va[j].x &= mlo; va[j].y &= mhi;
          vb[j].x &= mlo; vb[j].y &= mhi;
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint2 pb = *reinterpret_cast<const uint2*>(psw + nt * 8 * 16 + g * 16 + 4 * c);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma16816<BF16>(oacc[nt][4 * hf + j], va[j].x, vb[j].x, va[j].y, vb[j].y, pb.x, pb.y);
      }
    }
    __syncwarp();  // P buffer is rewritten next iteration
    if (smem_logits) {  // recording rule (reference :124, :309): rarely fails (protected window, masked tail)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (nt * 8 + 2 * c + i < Q) {
            float* lg = (i ? lg1 : lg0) + nt * 8 * lg_stride;
            if (!(pos0 <= max_pos)) lg[tok0] = -INFINITY;
            if (!(pos1 <= max_pos)) lg[tok1] = -INFINITY;
          }
    }
    if (!last) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) sacc[hf][nt][i] = snext[hf][nt][i];
    }
  };
  using T_ = std::true_type;
  using F_ = std::false_type;
  KVC_TR(4, clock64());
  if (nblk == 1) {
    consume_k(T_{}, 0, sacc);
  } else {
    consume_k(F_{}, 0, sacc);
    KVC_TR(5, clock64());
    for (int b = 0; b < nblk - 2; ++b) sweep_block(F_{}, F_{}, b);
    sweep_block(F_{}, T_{}, nblk - 2);
  }
  sweep_block(T_{}, F_{}, nblk - 1);
  KVC_TR(6, clock64());

  // ---- output: O / (L + 1e-6); lane (g,c) holds O[d = 16 dt + g (+8)][q = 2c, 2c+1] ----
  float inv[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float l = l_part[nt][i];
      l += __shfl_xor_sync(0xffffffffu, l, 4);
      l += __shfl_xor_sync(0xffffffffu, l, 8);
      l += __shfl_xor_sync(0xffffffffu, l, 16);
      inv[nt][i] = __fdividef(1.f, l + 1e-6f);
      const int qc = nt * 8 + 2 * c + i;
      if (qc < Q) {
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          outp[qc * kD + 16 * dt + g] = to_bits16<BF16>(oacc[nt][dt][i] * inv[nt][i]);
          outp[qc * kD + 16 * dt + g + 8] = to_bits16<BF16>(oacc[nt][dt][2 + i] * inv[nt][i]);
        }
      }
    }
  KVC_TR(8, clock64());
  if (kMove && pd >= 0 && lane == 0) p.pend_dst[item] = -1;  // applied (the producer read it long ago)
  if (kWrite && new_slot >= 0) {
    // reference csrc/kvcompress_cache_kernels.cu:55-88: K chunks, V elements, metric := head bias
    if (p.pend_kv != nullptr && new_o == 0) {
      // deferred append: the token sits alone in a fresh block, which the next compression most likely
      // drops again (the token is moved into an evicted slot): 2 x 128 partial-sector stores now and a
      // 4 KB block read then are replaced by 512 B of compact traffic each way
      uint16_t* pk = p.pend_kv + (int64_t)item * (2 * kD);
      if (lane < 16) *reinterpret_cast<uint4*>(pk + lane * 8) = newk;
#pragma unroll
      for (int r = 0; r < 4; ++r) pk[kD + lane + 32 * r] = newv[r];
      if (lane == 0) p.pend_slot[item] = new_slot;
    } else {
      uint16_t* kcw = reinterpret_cast<uint16_t*>(const_cast<void*>(p.k_cache)) + (new_slot >> 4) * p.kv_block_stride;
      uint16_t* vcw = reinterpret_cast<uint16_t*>(const_cast<void*>(p.v_cache)) + (new_slot >> 4) * p.kv_block_stride;
      if (lane < 16) *reinterpret_cast<uint4*>(kcw + lane * (kB * 8) + new_o * 8) = newk;
#pragma unroll
      for (int r = 0; r < 4; ++r) vcw[(lane + 32 * r) * kB + new_o] = newv[r];
      if (p.pend_slot != nullptr && lane == 0) p.pend_slot[item] = -1;
    }
    if (lane == 0) p.metric[new_slot] = p.head_bias[h];  // ordered before the reductions below by __syncwarp
  } else if (kWrite && p.pend_slot != nullptr && lane == 0) {
    p.pend_slot[item] = -1;
  }
  KVC_TR(9, clock64());
  if (smem_logits) {
    // metrics[slot] += sum_q p^2 from the warp's own shared-memory logits, token-major: a lane finalises four
    // consecutive tokens of all Q columns (16-byte shared-memory reads) and issues ONE 16-byte fire-and-forget
    // reduction for them (every slot belongs to exactly one warp; red.global.add.v4.f32, sm_90+).  Tokens
    // that fail the recording rule or lie beyond the context hold -inf logits: they add exactly 0.
    __syncwarp();
    float* mi = reinterpret_cast<float*>(psw);  // the P buffer is free now: [NT*8][2] = (M, 1 / (sum + 1e-6))
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (g == 0) {
          mi[(nt * 8 + 2 * c + i) * 2] = m_run[nt][i];
          mi[(nt * 8 + 2 * c + i) * 2 + 1] = inv[nt][i];
        }
    __syncwarp();
    const int ntok = nblk * kB;  // <= 32 blocks in this mode
    for (int t0 = 0; t0 < ntok; t0 += 4 * kWarp) {
      const int t4 = t0 + 4 * lane;
      const bool valid = t4 < ntok;
      const int64_t phys = __shfl_sync(0xffffffffu, my_bt, valid ? (t4 >> 4) : 0);
      if (valid) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        bool any = false;
        for (int q = 0; q < Q; ++q) {
          const float4 l = *reinterpret_cast<const float4*>(lg_base + q * lg_stride + t4);
          const float M = mi[2 * q], iv = mi[2 * q + 1];
          const float p0 = __expf(l.x - M) * iv, p1 = __expf(l.y - M) * iv;
          const float p2 = __expf(l.z - M) * iv, p3 = __expf(l.w - M) * iv;
          a0 += p0 * p0; a1 += p1 * p1; a2 += p2 * p2; a3 += p3 * p3;
          any |= (l.x != -INFINITY) | (l.y != -INFINITY) | (l.z != -INFINITY) | (l.w != -INFINITY);
        }
        if (any)
          asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                       ::"l"(p.metric + phys * kB + (t4 & (kB - 1))), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
      }
    }
  } else if (kFusedMetrics || (MODE == kModeV1 && p.record)) {
    // per-q max / inverse sum to every lane: column q lives in lanes with c == (q % 8) / 2
    __syncwarp();
    float* mi = reinterpret_cast<float*>(psw);  // reuse the P buffer: [NT*8][2] = (M, inv)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (g == 0) {
          mi[(nt * 8 + 2 * c + i) * 2] = m_run[nt][i];
          mi[(nt * 8 + 2 * c + i) * 2 + 1] = inv[nt][i];
        }
    __syncwarp();
    constexpr int U = 8;  // loads are issued in batches of U per lane so that their latencies overlap
    if (MODE == kModeV1) {
      const int n = ctx * Q;
      for (int i0 = 0; i0 < n; i0 += U * kWarp) {
        int64_t idx[U];
        int pos[U];
        float lg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * kWarp + lane;
          idx[u] = -1;
          if (i < n) {
            const int t = i / Q;
            idx[u] = ((int64_t)bt[t / kB] * kB + t % kB) * Q + i % Q;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          pos[u] = (idx[u] >= 0) ? p.kv_position[idx[u] / Q] : 0x7fffffff;
          lg[u] = (idx[u] >= 0) ? p.metric[idx[u]] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (idx[u] >= 0 && pos[u] <= max_pos) {
            const int q = (i0 + u * kWarp + lane) % Q;
            p.metric[idx[u]] = __expf(lg[u] - mi[q * 2]) * mi[q * 2 + 1];
          }
        }
      }
    } else {
      // token-major like the shared-memory path: a lane finalises four consecutive tokens (one 16-byte position
      // load, Q independent 16-byte logit loads from the L2-resident scratch row, all issued before the first use),
      // two such groups per pass, and issues one 16-byte reduction per group
      const float* lgi = scratch + (int64_t)item * Q * scratch_tokens;
      const int ntok = nblk * kB;
      constexpr int G = (NT == 1) ? 2 : 1;   // register budget: G * NT * 8 float4 of logits in flight per lane
      for (int t0 = 0; t0 < ntok; t0 += G * 4 * kWarp) {
        int t4[G];
        int64_t slot4[G];
        int4 pos4[G];
        float4 lgv[G][NT * 8];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          t4[u] = t0 + 4 * (lane + kWarp * u);
          const bool valid = t4[u] < ntok;
          const int b = valid ? (t4[u] >> 4) : 0;
          // b differs between lanes here: every lane executes the shuffle, entries beyond 32 come from the table
          const int sh = __shfl_sync(0xffffffffu, my_bt, b & (kWarp - 1));
          const int64_t phys = (b < kWarp) ? (int64_t)sh : (int64_t)bt[b];
          slot4[u] = valid ? phys * kB + (t4[u] & (kB - 1)) : -1;
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
          pos4[u] = make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
          if (slot4[u] >= 0) {
            pos4[u] = *reinterpret_cast<const int4*>(p.kv_position + slot4[u]);
#pragma unroll
            for (int q = 0; q < NT * 8; ++q)
              if (q < Q) lgv[u][q] = *reinterpret_cast<const float4*>(lgi + (int64_t)q * scratch_tokens + t4[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
          if (slot4[u] >= 0) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int q = 0; q < NT * 8; ++q)
              if (q < Q) {
                const float M = mi[2 * q], iv = mi[2 * q + 1];
                const float p0 = __expf(lgv[u][q].x - M) * iv, p1 = __expf(lgv[u][q].y - M) * iv;
                const float p2 = __expf(lgv[u][q].z - M) * iv, p3 = __expf(lgv[u][q].w - M) * iv;
                a0 += p0 * p0; a1 += p1 * p1; a2 += p2 * p2; a3 += p3 * p3;
              }
            // recording rule (reference :124, :309); tokens beyond the context hold -inf logits: exactly 0
            a0 = (pos4[u].x <= max_pos) ? a0 : 0.f;
            a1 = (pos4[u].y <= max_pos) ? a1 : 0.f;
            a2 = (pos4[u].z <= max_pos) ? a2 : 0.f;
            a3 = (pos4[u].w <= max_pos) ? a3 : 0.f;
            if (a0 != 0.f || a1 != 0.f || a2 != 0.f || a3 != 0.f)
              asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
                           ::"l"(p.metric + slot4[u]), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
          }
        }
      }
    }
  }
  KVC_TR(10, clock64());
  // inputs declared independent of the previous kernel: still order this grid's completion after it
  if (pdl == 3) asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef KVC_ATTN_TRACE
  if (p.trace != nullptr && lane == 0) {
    p.trace[(size_t)item * 16 + 7] = kvc_globaltimer();
    // pack the exit clock with the SM id: clocks of different SMs are not comparable
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[(size_t)item * 16 + 6] = (p.trace[(size_t)item * 16 + 6] & 0xffffffffffffull) | ((unsigned long long)smid << 48);
  }
#endif
}

// -----------------------------------------------------------------------------
// v2 partition kernel on the 5th-generation tensor cores (tcgen05 + TMEM) for LONG per-head contexts
// (uncompressed or lightly compressed caches: thousands of tokens per head, few (sequence, head) items).
// grid (num_kv_heads, num_seqs, partitions), 128 threads; a CTA owns one 512-token partition of one
// (sequence, kv head) and serves all Q query heads of the GQA group (K / V read once):
//   S^T[128 tok x 16 q] = Ktile[128 tok x 128 d] . Q^T      8 x tcgen05.mma (M=128, N=16, K=16) per 128-token tile,
//                                                          fp32 accumulators in TMEM (16 columns per tile)
//   softmax over the partition: one thread = one token row per tile (tcgen05.ld.32x32b.x16), column
//                               max / sum by warp shuffles + a 4-warp merge in shared memory
//   O^T[128 d x 16 q] += Vtile^T[128 d x 128 tok] . P^T     8 x tcgen05.mma per tile, accumulated over the tiles
// Operands are staged global -> registers -> shared memory in the canonical no-swizzle K-major core-matrix
// layout (8 rows x 16 B; the convention tools/tcgen05_probe.cu and kvc_prefill.cu verified on the B200):
//   K tile   A operand, row = token:  the cache block layout [16 j][16 t][8] is already core-matrix shaped, a
//            (block, j) pair is a contiguous 256 B run on both sides           LBO 2048 B, SBO 128 B
//   V tile   A operand, row = d:      a 16 B piece = 8 tokens of one row d                same LBO / SBO
//   Q, P     B operands, 16 rows (query heads):                                LBO  256 B, SBO 128 B
// K and V tiles share two 32 KB buffers (all QK^T products of the partition are in TMEM before the first V tile
// is staged); tokens beyond the context are masked to -inf in S and zeroed in V (reference :370-379).
// Outputs are the reference's v2 buffers (exp_sums, max_logits, tmp_out, raw logits parked in kv_metric_out);
// paged_attention_reduce_kernel below finishes.  Arithmetic: fp16/bf16 operands, fp32 accumulate, probabilities
// rounded to the cache type before PV like the reference (:356-357).
// -----------------------------------------------------------------------------
constexpr int kTcTile = 128;                      // tokens per MMA tile (8 cache blocks)
constexpr int kTcTileBytes = kTcTile * kD * 2;    // 32 KB
constexpr int kTcNQ = 16;                         // UMMA N: query heads padded to 16

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // sm_100 descriptor version
  return d;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(128) paged_attention_tc_kernel(AttnParams p) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char* sBuf = tc_smem;                                   // [2][32 KB] K tiles, then V tiles
  unsigned char* sQ = tc_smem + 2 * kTcTileBytes;                  // 4 KB
  unsigned char* sP = sQ + kTcNQ * kD * 2;                         // [4][4 KB]
  float* red = reinterpret_cast<float*>(sP + 4 * kTcNQ * kTcTile * 2);  // [4 warps][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 4 * kTcNQ);   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int h = blockIdx.x, s = blockIdx.y, part = blockIdx.z;
  const int H = p.num_kv_heads, Q = p.Q;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seq_head = s * H + h;
  const int ctx = p.context_lens[seq_head];
  const int nblk = (ctx + kB - 1) / kB;
  const int blk_begin = part * (kPartition / kB);
  if (blk_begin >= nblk) return;  // reference :139-142 (CTA-uniform, before any allocation)
  const int blk_end = min(nblk, blk_begin + kPartition / kB);
  const int tok_begin = blk_begin * kB;
  const int ntiles = (blk_end - blk_begin + 7) / 8;
  const int* __restrict__ bt = p.block_tables + (int64_t)seq_head * p.max_num_blocks_per_seq;
  const int max_pos = p.last_position[s] - p.kv_metric_buffer_len[s];
  const unsigned char* __restrict__ kc = reinterpret_cast<const unsigned char*>(p.k_cache);
  const unsigned char* __restrict__ vc = reinterpret_cast<const unsigned char*>(p.v_cache);
  const int64_t block_bytes = p.kv_block_stride * 2;

  __shared__ int s_bt[kPartition / kB];   // the partition's block ids: no dependent table load per copy
  if (tid < kPartition / kB) s_bt[tid] = (blk_begin + tid < blk_end) ? bt[blk_begin + tid] : 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- Q as the B operand: row n = query head, 16 chunks of 8 d ----
  for (int i = tid; i < kTcNQ * 16; i += 128) {
    const int n = i >> 4, c = i & 15;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < Q)
      v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.q) + (int64_t)s * p.q_stride +
                                          (int64_t)(h * Q + n) * kD + c * 8);
    *reinterpret_cast<uint4*>(sQ + c * 256 + n * 16) = v;
  }
  // Tiles stream through two 32 KB buffers in the order K(0) .. K(n-1), V(0) .. V(n-1); stream index i lives in
  // buffer i & 1.  Staging is asynchronous (cp.async, 16 B per copy, no registers): the copies of tile i+1 are in
  // flight while tile i is multiplied, and the first V tiles stream in under the softmax.
  //   K tile: smem (j, tok) = j * 2048 + tok * 16        <-  block (tok / 16): j * 256 + (tok % 16) * 16
  //   V tile: smem (tok / 8, d) = (tok / 8) * 2048 + d * 16  <-  block: d * 32 + (tok % 16 / 8) * 16
  auto issue_tile = [&](int i) {
    unsigned char* dst = sBuf + (i & 1) * kTcTileBytes;
    const bool is_v = i >= ntiles;
    const int t = is_v ? i - ntiles : i;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int idx = it * 128 + tid;          // 2048 pieces of 16 B
      int bi, src_off, dst_off;
      if (!is_v) {
        const int run = idx >> 4, piece = idx & 15;  // run = (block in tile) * 16 + j
        bi = run >> 4;
        const int j = run & 15;
        src_off = j * 256 + piece * 16;
        dst_off = j * 2048 + bi * 256 + piece * 16;
      } else {
        bi = idx >> 8;
        const int rem = idx & 255, d = rem >> 1, half = rem & 1;   // 256 pieces per block: (d, half)
        src_off = d * 32 + half * 16;
        dst_off = (bi * 2 + half) * 2048 + d * 16;
      }
      const int b = blk_begin + t * 8 + bi;
      if (b < blk_end) {
        const unsigned char* src = (is_v ? vc : kc) + (int64_t)s_bt[b - blk_begin] * block_bytes + src_off;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + dst_off)), "l"(src) : "memory");
      } else {
        *reinterpret_cast<uint4*>(dst + dst_off) = make_uint4(0, 0, 0, 0);   // beyond the context: zero operand
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // tokens >= ctx of the context's last block are zeroed in the staged V tile (reference :370-379: 0 * NaN)
  auto fix_v_tail = [&](int i) {
    const int t = i - ntiles;
    const int b_last = nblk - 1;   // the only block that can hold tokens beyond the context
    const int bi = b_last - (blk_begin + t * 8);
    if ((ctx & (kB - 1)) == 0 || bi < 0 || bi >= 8) return;
    unsigned char* dst = sBuf + (i & 1) * kTcTileBytes;
    for (int rem = tid; rem < 256; rem += 128) {
      const int d = rem >> 1, half = rem & 1;
      const int valid = ctx - (b_last * kB + half * 8);
      if (valid < 8) {
        uint4* ptr = reinterpret_cast<uint4*>(dst + (bi * 2 + half) * 2048 + d * 16);
        uint4 v = *ptr;
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t lo = (2 * e < valid) ? 0xffffu : 0u, hi = (2 * e + 1 < valid) ? 0xffff0000u : 0u;
          w[e] &= lo | hi;
        }
        *ptr = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  };
  auto publish = [&]() {  // generic-proxy shared-memory writes -> tensor core, CTA-wide
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  const int nstream = 2 * ntiles;
  __syncthreads();   // s_bt
  issue_tile(0);
  issue_tile(1);   // nstream >= 2
  publish();       // Q tile, TMEM slot
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t kFmt = BF16 ? 1u : 0u;
  // instruction descriptor: D = f32, A = B = f16 / bf16, both K-major, N = 16, M = 128
  constexpr uint32_t idesc = (1u << 4) | (kFmt << 7) | (kFmt << 10) | ((uint32_t)(kTcNQ >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
  auto mma = [&](uint32_t d_col, uint32_t a_base, uint32_t b_base, bool first_clears, int bar) {  // one thread
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t ad = tc_desc(a_base + ks * 4096, 2048, 128);
      const uint64_t bd = tc_desc(b_base + ks * 512, 256, 128);
      const uint32_t acc = (ks > 0 || !first_clears) ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + d_col), "l"(ad), "l"(bd),
          "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(&bars[bar]))
                 : "memory");
  };
  uint32_t phases = 0u, pending = 0u;   // per barrier: parity of the next wait; a commit not yet waited for
  auto wait_bar = [&](int b) {
    if (!((pending >> b) & 1u)) return;
    mbar_wait(smem_u32(&bars[b]), (phases >> b) & 1u);
    phases ^= 1u << b;
    pending &= ~(1u << b);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  constexpr uint32_t kOCol = 4 * kTcNQ;
  float mx[kTcNQ], inv[kTcNQ];

  for (int i = 0; i < nstream; ++i) {
    if (i == ntiles) {
      // every QK^T product of the partition is needed now; the first V tiles keep streaming in meanwhile
      wait_bar(0);
      wait_bar(1);
      // ---- softmax over the partition: thread = token row (tile t: token tok_begin + 128 t + tid) ----
      float sreg[4][kTcNQ];
      float mloc[kTcNQ];
#pragma unroll
      for (int q = 0; q < kTcNQ; ++q) mloc[q] = -INFINITY;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < ntiles) {
          uint32_t r[16];
          tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + t * kTcNQ, r);
          const int tok = tok_begin + t * kTcTile + tid;
          const bool live = tok < ctx;
          int64_t slot = 0;
          bool rec = false;
          if (live && p.record) {
            slot = (int64_t)s_bt[(tok >> 4) - blk_begin] * kB + (tok & 15);
            rec = p.kv_position[slot] <= max_pos;   // reference :309
          }
#pragma unroll
          for (int q = 0; q < kTcNQ; ++q) {
            float v = __uint_as_float(r[q]) * p.scale;
            if (p.alibi_slopes != nullptr && q < Q) v += p.alibi_slopes[h * Q + q] * (float)(tok - ctx + 1);  // :253
            v = live ? v : -INFINITY;
            sreg[t][q] = v;
            mloc[q] = fmaxf(mloc[q], v);
            if (rec && q < Q) p.metric[slot * Q + q] = v;   // raw logit, finalised by the reduce kernel
          }
        } else {
#pragma unroll
          for (int q = 0; q < kTcNQ; ++q) sreg[t][q] = -INFINITY;
        }
      }
#pragma unroll
      for (int q = 0; q < kTcNQ; ++q) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) mloc[q] = fmaxf(mloc[q], __shfl_xor_sync(0xffffffffu, mloc[q], d));
        if (lane == 0) red[warp * kTcNQ + q] = mloc[q];
      }
      __syncthreads();
      float lsum[kTcNQ];
#pragma unroll
      for (int q = 0; q < kTcNQ; ++q) {
        mx[q] = fmaxf(fmaxf(red[q], red[kTcNQ + q]), fmaxf(red[2 * kTcNQ + q], red[3 * kTcNQ + q]));
        lsum[q] = 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (t < ntiles) {
          const int r_ = tid;  // token row inside the tile
#pragma unroll
          for (int q = 0; q < kTcNQ; ++q) {
            const float pe = __expf(sreg[t][q] - mx[q]);   // the partition holds >= 1 live token: mx finite
            lsum[q] += pe;
            *reinterpret_cast<uint16_t*>(sP + t * (kTcNQ * kTcTile * 2) + (r_ >> 3) * 256 + q * 16 + (r_ & 7) * 2) =
                to_bits16<BF16>(pe);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kTcNQ; ++q) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) lsum[q] += __shfl_xor_sync(0xffffffffu, lsum[q], d);
        if (lane == 0) red[warp * kTcNQ + q] = lsum[q];
      }
      __syncthreads();   // red[] visible; the P tiles are published together with the first V tile below
#pragma unroll
      for (int q = 0; q < kTcNQ; ++q) {
        const float L = red[q] + red[kTcNQ + q] + red[2 * kTcNQ + q] + red[3 * kTcNQ + q];
        inv[q] = __fdividef(1.f, L + 1e-6f);  // reference :298
        if (tid == q && q < Q) {               // reference :316-326
          const int64_t idx = ((int64_t)s * p.num_heads + h * Q + q) * p.max_partitions + part;
          p.max_logits[idx] = mx[q];
          p.exp_sums[idx] = L;
        }
      }
    }
    // tile i has landed (this thread's copies; the CTA-wide publish below makes all of them visible)
    if (i + 1 < nstream) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (i >= ntiles) {
      __syncthreads();   // every thread's pieces of the V tile are in shared memory before the tail is patched
      fix_v_tail(i);
    }
    publish();
    if (tid == 0) {
      const uint32_t a_base = smem_u32(sBuf + (i & 1) * kTcTileBytes);
      if (i < ntiles) mma(i * kTcNQ, a_base, smem_u32(sQ), true, i & 1);                      // S^T(i) = K Q^T
      else mma(kOCol, a_base, smem_u32(sP + (i - ntiles) * (kTcNQ * kTcTile * 2)), i == ntiles, i & 1);  // O^T += V^T P^T
    }
    pending |= 1u << (i & 1);
    if (i + 2 < nstream) {
      wait_bar(i & 1);       // the MMAs of tile i have read the buffer
      issue_tile(i + 2);
    }
  }
  wait_bar(0);
  wait_bar(1);

  // ---- partial output: thread = head dim d ----
  {
    uint32_t r[16];
    tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + kOCol, r);
    uint16_t* to = reinterpret_cast<uint16_t*>(p.tmp_out);
#pragma unroll
    for (int q = 0; q < kTcNQ; ++q)
      if (q < Q)
        to[(((int64_t)s * p.num_heads + h * Q + q) * p.max_partitions + part) * kD + tid] =
            to_bits16<BF16>(__uint_as_float(r[q]) * inv[q]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// -----------------------------------------------------------------------------
// v2 reduce.  grid (num_kv_heads, num_seqs), 128 threads.  reference :525-652.
// -----------------------------------------------------------------------------
// grid (num_kv_heads, num_seqs, max_partitions): every CTA derives the final per-query-head (max, 1 / sum) from the
// partitions' statistics (Q x nparts values, one warp lane per partition), the CTA of partition 0 reduces the
// output, and EVERY CTA finalises the recorded probabilities of its own 512 tokens -- all Q columns of a token at
// once, 8 tokens per thread in flight -- instead of one CTA walking the whole context Q times behind dependent loads.
template <bool BF16>
__global__ void __launch_bounds__(kAttnThreads) paged_attention_reduce_kernel(AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* wts = reinterpret_cast<float*>(smem_raw);  // [Q][max_partitions] rescaled exp sums * 1 / sum
  __shared__ float s_M[16], s_inv[16];
  const int h = blockIdx.x, s = blockIdx.y, part = blockIdx.z, tid = threadIdx.x;
  const int lane = tid % kWarp, wid = tid / kWarp;
  const int H = p.num_kv_heads, Q = p.Q;
  const int seq_head = s * H + h;
  const int ctx = p.context_lens[seq_head];
  const int nparts = (ctx + kPartition - 1) / kPartition;
  if (part > 0 && (part >= nparts || !p.record)) return;   // CTA-uniform
  const int* __restrict__ bt = p.block_tables + (int64_t)seq_head * p.max_num_blocks_per_seq;
  const int max_pos = p.last_position[s] - p.kv_metric_buffer_len[s];
  // ---- per query head: M = max_j m_j, L = sum_j l_j exp(m_j - M)  (reference :606-614); warp w takes heads w, w+4, ..
  for (int q = wid; q < Q; q += kAttnWarps) {
    const int64_t base = ((int64_t)s * p.num_heads + h * Q + q) * p.max_partitions;
    float M = -INFINITY;
    for (int j = lane; j < nparts; j += kWarp) M = fmaxf(M, p.max_logits[base + j]);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, d));
    float L = 0.f;
    for (int j = lane; j < nparts; j += kWarp) {
      const float r = p.exp_sums[base + j] * expf(p.max_logits[base + j] - M);
      wts[q * p.max_partitions + j] = r;
      L += r;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) L += __shfl_xor_sync(0xffffffffu, L, d);
    if (lane == 0) {
      s_M[q] = M;
      s_inv[q] = __fdividef(1.f, L + 1e-6f);  // :614
    }
  }
  __syncthreads();
  if (part == 0) {
    const int d = tid;
    for (int q = 0; q < Q; ++q) {
      const int qh = h * Q + q;
      const int64_t base = ((int64_t)s * p.num_heads + qh) * p.max_partitions;
      const uint16_t* tp = reinterpret_cast<const uint16_t*>(p.tmp_out) + base * kD;
      float acc = 0.f;
      if (nparts == 1) {
        acc = from_bits16<BF16>(tp[d]);  // reference :555-562: plain copy
      } else {
        const float inv = s_inv[q];
        for (int j = 0; j < nparts; ++j)
          acc += from_bits16<BF16>(tp[(int64_t)j * kD + d]) * wts[q * p.max_partitions + j] * inv;  // :625
      }
      reinterpret_cast<uint16_t*>(p.out)[((int64_t)s * p.num_heads + qh) * kD + d] = to_bits16<BF16>(acc);
    }
  }
  if (p.record && part < nparts) {
    const int t_begin = part * kPartition, t_end = min(ctx, t_begin + kPartition);
    constexpr int U = 4;  // tokens per thread in flight (512 tokens / 128 threads)
    int64_t slot[U];
    int pos[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t_begin + u * kAttnThreads + tid;
      slot[u] = (t < t_end) ? (int64_t)bt[t / kB] * kB + t % kB : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) pos[u] = (slot[u] >= 0) ? p.kv_position[slot[u]] : 0x7fffffff;
    if (Q == 4) {  // one 16-byte access per token
      float4 lg[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pos[u] <= max_pos) lg[u] = *reinterpret_cast<const float4*>(p.metric + slot[u] * 4);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pos[u] <= max_pos) {
          float4 o;
          o.x = __expf(lg[u].x - s_M[0]) * s_inv[0];
          o.y = __expf(lg[u].y - s_M[1]) * s_inv[1];
          o.z = __expf(lg[u].z - s_M[2]) * s_inv[2];
          o.w = __expf(lg[u].w - s_M[3]) * s_inv[3];
          *reinterpret_cast<float4*>(p.metric + slot[u] * 4) = o;
        }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (pos[u] <= max_pos)
          for (int q = 0; q < Q; ++q) {
            const float lg = p.metric[slot[u] * Q + q];
            p.metric[slot[u] * Q + q] = __expf(lg - s_M[q]) * s_inv[q];
          }
    }
  }
}

#ifdef KVC_ATTN_TRACE
static unsigned long long* g_attn_trace = nullptr;
#endif
// partition kernel of paged attention v2: 1 = mma.sync fragments straight from global memory, 2 = tcgen05 / TMEM
// tiles (paged_attention_tc_kernel); 0 = the faster of the two on the B200 (currently 1)
static int g_v2_mode = -1;
static int attention_v2_mode() {
  if (g_v2_mode < 0) {
    const char* v = getenv("KVC_ATTN_V2_MODE");
    g_v2_mode = v ? atoi(v) : 0;
    if (g_v2_mode < 0 || g_v2_mode > 2) g_v2_mode = 0;
  }
  return g_v2_mode == 0 ? 1 : g_v2_mode;
}
static int g_pdl_mode = -1;  // -1 = not set yet: KVC_ATTN_PDL from the environment, else 1
static int attention_pdl_mode() {
  if (g_pdl_mode < 0) {
    const char* v = getenv("KVC_ATTN_PDL");
    g_pdl_mode = v ? atoi(v) : 1;
    if (g_pdl_mode < 0 || g_pdl_mode > 3) g_pdl_mode = 1;
  }
  return g_pdl_mode;
}

template <bool BF16, int NT, int MODE>
static size_t attn_smem_bytes(int fused_tokens, int Q) {
  constexpr int NQ = NT * 8;
  size_t b = sizeof(float) * (kAttnWarps * NQ * kOStride + 2 * kAttnWarps * NQ + 2 * NQ);
  b += sizeof(uint16_t) * kAttnWarps * NT * 8 * 16;
  if (MODE == kModeFused) b += sizeof(float) * (size_t)fused_tokens * Q;
  return b;
}

template <bool BF16, int NT, int MODE>
static int launch_attention(const AttnParams& p, int num_seqs, int partitions, int fused_tokens,
                            cudaStream_t st) {
  const size_t smem = attn_smem_bytes<BF16, NT, MODE>(fused_tokens, p.Q);
  auto kern = paged_attention_kernel<BF16, NT, MODE>;
  if (smem > 48 * 1024) {
    KVC_REQUIRE(smem <= 227 * 1024, KVC_ERR_UNSUPPORTED,
                "paged attention: %zu bytes of shared memory needed (max_context_len too large for the "
                "fused-metrics kernel)", smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "paged attention smem attr: %s", cudaGetErrorString(e));
  }
  kern<<<dim3(p.num_kv_heads, num_seqs, partitions), kAttnThreads, smem, st>>>(p);
  KVC_CHECK_LAUNCH("paged_attention_kernel");
  return KVC_OK;
}

template <int MODE>
static int dispatch_attention(const AttnParams& p, int num_seqs, int partitions, int fused_tokens,
                              int dtype, cudaStream_t st) {
  const int nt = (p.Q + 7) / 8;
  if (dtype == KVC_F16) {
    if (nt == 1) return launch_attention<false, 1, MODE>(p, num_seqs, partitions, fused_tokens, st);
    if (nt == 2) return launch_attention<false, 2, MODE>(p, num_seqs, partitions, fused_tokens, st);
  } else {
    if (nt == 1) return launch_attention<true, 1, MODE>(p, num_seqs, partitions, fused_tokens, st);
    if (nt == 2) return launch_attention<true, 2, MODE>(p, num_seqs, partitions, fused_tokens, st);
  }
  KVC_FAIL(KVC_ERR_UNSUPPORTED, "paged attention: %d query heads per kv head (max 16)", p.Q);
}


template <bool BF16, int NT, int MODE>
static int launch_stream(const AttnParams& p, int num_seqs, float* scratch, int scratch_tokens, cudaStream_t st) {
  // fused modes: keep the raw logits in shared memory when [Q][padded ctx] fits next to the ring
  // (<= 32 blocks so the recording bits fit one 64-bit register, and <= 3.5 KB per warp)
  int smem_logit_tokens = 0;
  constexpr size_t kPend = 0;
  if ((MODE == kModeFused || MODE == kModeFusedWrite || MODE == kModeFusedWriteMove) && scratch_tokens <= 32 * kB &&
      (size_t)scratch_tokens * p.Q * 4 + kPend <= 3584)
    smem_logit_tokens = scratch_tokens;
  const size_t smem = (size_t)kConsumers * (kNUnit * kUnitBytes + NT * 8 * 16 * 2 + 2 * kNUnit * 8 + 32 + kPend +
                                            (size_t)smem_logit_tokens * p.Q * 4);
  const int items = num_seqs * p.num_kv_heads;
  const dim3 grid((items + kConsumers - 1) / kConsumers);
  cudaError_t e;
  // launch mode: see kvc_set_attention_pdl (include/kvc_b200.h)
  const int pdl_mode = attention_pdl_mode();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kStreamThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_mode > 0 ? 1 : 0;
  // mode 3 lets consecutive launches overlap completely.  The fused modes park raw logits in the caller's
  // workspace when they do not fit in shared memory, and consecutive launches may be handed the same
  // workspace: the wait then has to come before the first logit is parked (mode 2).
  static const int pair = [] { const char* e_ = getenv("KVC_ATTN_PAIR"); return e_ ? atoi(e_) : 1; }();  // 4 KB copies
  int pdl = pdl_mode;
  const bool global_logits = (MODE == kModeFused || MODE == kModeFusedWrite || MODE == kModeFusedWriteMove) &&
                             smem_logit_tokens == 0;
  if (pdl == 3 && global_logits) pdl = 2;
  if (p.alibi_slopes) {
    auto kern = paged_attention_stream_kernel<BF16, NT, MODE, true>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, kern, p, scratch, scratch_tokens, smem_logit_tokens, pdl, pair);
  } else {
    auto kern = paged_attention_stream_kernel<BF16, NT, MODE, false>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, kern, p, scratch, scratch_tokens, smem_logit_tokens, pdl, pair);
  }
  if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "paged attention (stream) launch: %s", cudaGetErrorString(e));
  KVC_CHECK_LAUNCH("paged_attention_stream_kernel");
  return KVC_OK;
}

template <int MODE>
static int dispatch_stream(const AttnParams& p, int num_seqs, float* scratch, int scratch_tokens, int dtype,
                           cudaStream_t st) {
  const int nt = (p.Q + 7) / 8;
  if (dtype == KVC_F16) {
    if (nt == 1) return launch_stream<false, 1, MODE>(p, num_seqs, scratch, scratch_tokens, st);
    if (nt == 2) return launch_stream<false, 2, MODE>(p, num_seqs, scratch, scratch_tokens, st);
  } else {
    if (nt == 1) return launch_stream<true, 1, MODE>(p, num_seqs, scratch, scratch_tokens, st);
    if (nt == 2) return launch_stream<true, 2, MODE>(p, num_seqs, scratch, scratch_tokens, st);
  }
  KVC_FAIL(KVC_ERR_UNSUPPORTED, "paged attention: %d query heads per kv head (max 16)", p.Q);
}

// one warp per (seq, head) pays off when there are enough items to fill the machine
static bool use_stream_kernel(int num_seqs, int num_kv_heads, int max_context_len) {
  return (long long)num_seqs * num_kv_heads >= (long long)sm_count() * 4 || max_context_len <= 512;
}

static int fill_params(AttnParams& p, void* out, float* metric, const void* query,
                       const void* key_cache, const void* value_cache, int num_seqs, int num_heads,
                       int num_kv_heads, int head_size, float scale, const int* block_tables,
                       const int* context_lens, const int* kv_position, const int* last_position,
                       const int* kv_metric_buffer_len, int block_size, int max_num_blocks_per_seq,
                       int64_t q_stride, int64_t kv_block_stride, const float* alibi_slopes, int dtype,
                       int record) {
  KVC_REQUIRE(head_size == 128, KVC_ERR_UNSUPPORTED, "Unsupported head size: %d", head_size);  // :745-770
  KVC_REQUIRE(block_size == 16, KVC_ERR_UNSUPPORTED, "Unsupported block size: %d", block_size);
  KVC_REQUIRE(dtype == KVC_F16 || dtype == KVC_BF16, KVC_ERR_UNSUPPORTED,
              "Unsupported data type: %d (fp16/bf16 query and cache, kv_cache_dtype 'auto')", dtype);
  KVC_REQUIRE(out && query && key_cache && value_cache && block_tables && context_lens &&
                  kv_position && last_position && kv_metric_buffer_len,
              KVC_ERR_INVALID, "paged attention: null tensor");
  KVC_REQUIRE(!record || metric, KVC_ERR_INVALID, "paged attention: kv_metric_out is null");
  KVC_REQUIRE(num_seqs >= 0 && num_heads > 0 && num_kv_heads > 0 && num_heads % num_kv_heads == 0,
              KVC_ERR_INVALID, "paged attention: bad head counts %d/%d", num_heads, num_kv_heads);
  KVC_REQUIRE(q_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(query) & 15) == 0 &&
                  kv_block_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(key_cache) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(value_cache) & 15) == 0,
              KVC_ERR_INVALID, "paged attention: query / cache must be 16-byte aligned");
  p.out = out; p.metric = metric; p.exp_sums = nullptr; p.max_logits = nullptr; p.tmp_out = nullptr;
  p.q = query; p.k_cache = key_cache; p.v_cache = value_cache;
  p.block_tables = block_tables; p.context_lens = context_lens; p.kv_position = kv_position;
  p.last_position = last_position; p.kv_metric_buffer_len = kv_metric_buffer_len;
  p.alibi_slopes = alibi_slopes;
  p.num_heads = num_heads; p.num_kv_heads = num_kv_heads; p.Q = num_heads / num_kv_heads;
  p.scale = scale; p.max_num_blocks_per_seq = max_num_blocks_per_seq;
  p.q_stride = q_stride; p.kv_block_stride = kv_block_stride;
  p.record = record ? 1 : 0; p.max_partitions = 1; p.num_seqs = num_seqs;
  p.new_k = nullptr; p.new_v = nullptr; p.new_k_stride = 0; p.new_v_stride = 0;
  p.slot_mapping = nullptr; p.head_bias = nullptr;
  p.pend_kv = nullptr; p.pend_slot = nullptr; p.pend_dst = nullptr;
#ifdef KVC_ATTN_TRACE
  p.trace = g_attn_trace;
#endif
  return KVC_OK;
}

}  // namespace kvc

using namespace kvc;

extern "C" size_t kvc_paged_attention_fused_metrics_workspace_bytes(int num_seqs, int num_heads, int max_context_len);

extern "C" int kvc_paged_attention_v1(
    void* out, float* kv_metric_out, const void* query, const void* key_cache,
    const void* value_cache, int num_seqs, int num_heads, int num_kv_heads, int head_size,
    float scale, const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size, int max_context_len,
    int max_num_blocks_per_seq, int64_t q_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype, int record_kv_metrics, kvc_stream_t stream) {
  (void)max_context_len;  // the reference sizes shared memory from it (:733-739); not needed here
  AttnParams p;
  int rc = fill_params(p, out, kv_metric_out, query, key_cache, value_cache, num_seqs, num_heads,
                       num_kv_heads, head_size, scale, block_tables, context_lens, kv_position,
                       last_position, kv_metric_buffer_len, block_size, max_num_blocks_per_seq,
                       q_stride, kv_block_stride, alibi_slopes, dtype, record_kv_metrics);
  if (rc != KVC_OK) return rc;
  if (num_seqs == 0) return KVC_OK;
  if (use_stream_kernel(num_seqs, num_kv_heads, max_context_len))
    return dispatch_stream<kModeV1>(p, num_seqs, nullptr, 0, dtype, as_stream(stream));
  return dispatch_attention<kModeV1>(p, num_seqs, 1, 0, dtype, as_stream(stream));
}

extern "C" int kvc_paged_attention_v2(
    void* out, float* kv_metric_out, float* exp_sums, float* max_logits, void* tmp_out,
    float* tmp_kv_metric_out, const void* query, const void* key_cache, const void* value_cache,
    int num_seqs, int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size, int max_context_len,
    int max_num_blocks_per_seq, int64_t q_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype, int record_kv_metrics, kvc_stream_t stream) {
  (void)tmp_kv_metric_out;  // raw logits are parked in kv_metric_out itself
  AttnParams p;
  int rc = fill_params(p, out, kv_metric_out, query, key_cache, value_cache, num_seqs, num_heads,
                       num_kv_heads, head_size, scale, block_tables, context_lens, kv_position,
                       last_position, kv_metric_buffer_len, block_size, max_num_blocks_per_seq,
                       q_stride, kv_block_stride, alibi_slopes, dtype, record_kv_metrics);
  if (rc != KVC_OK) return rc;
  KVC_REQUIRE(exp_sums && max_logits && tmp_out, KVC_ERR_INVALID, "paged attention v2: null tmp buffer");
  KVC_REQUIRE(max_context_len > 0, KVC_ERR_INVALID, "paged attention v2: max_context_len <= 0");
  if (num_seqs == 0) return KVC_OK;
  p.exp_sums = exp_sums; p.max_logits = max_logits; p.tmp_out = tmp_out;
  p.max_partitions = (max_context_len + kPartition - 1) / kPartition;  // reference :914
  cudaStream_t st = as_stream(stream);
  // partition kernel: tcgen05 / TMEM tiles or mma.sync fragments (kvc_set_attention_v2_mode)
  const bool use_tc = attention_v2_mode() == 2;
  if (use_tc && p.Q <= kTcNQ) {
    const size_t tc_smem_bytes = 2 * kTcTileBytes + kTcNQ * kD * 2 + 4 * kTcNQ * kTcTile * 2 + 4 * kTcNQ * 4 + 64;
    cudaError_t e;
    const dim3 grid(num_kv_heads, num_seqs, p.max_partitions);
    if (dtype == KVC_F16) {
      e = cudaFuncSetAttribute(paged_attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)tc_smem_bytes);
      if (e == cudaSuccess) paged_attention_tc_kernel<false><<<grid, 128, tc_smem_bytes, st>>>(p);
    } else {
      e = cudaFuncSetAttribute(paged_attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)tc_smem_bytes);
      if (e == cudaSuccess) paged_attention_tc_kernel<true><<<grid, 128, tc_smem_bytes, st>>>(p);
    }
    if (e != cudaSuccess) KVC_FAIL(KVC_ERR_CUDA, "paged attention (tcgen05) smem attr: %s", cudaGetErrorString(e));
    KVC_CHECK_LAUNCH("paged_attention_tc_kernel");
  } else {
    rc = dispatch_attention<kModeV2>(p, num_seqs, p.max_partitions, 0, dtype, st);
    if (rc != KVC_OK) return rc;
  }
  const size_t smem = sizeof(float) * (size_t)p.max_partitions * p.Q;
  KVC_REQUIRE(smem <= 48 * 1024, KVC_ERR_UNSUPPORTED, "paged attention v2: %d partitions x %d query heads per kv head",
              p.max_partitions, p.Q);
  const dim3 rgrid(num_kv_heads, num_seqs, record_kv_metrics ? p.max_partitions : 1);
  if (dtype == KVC_F16)
    paged_attention_reduce_kernel<false><<<rgrid, kAttnThreads, smem, st>>>(p);
  else
    paged_attention_reduce_kernel<true><<<rgrid, kAttnThreads, smem, st>>>(p);
  KVC_CHECK_LAUNCH("paged_attention_reduce_kernel");
  return KVC_OK;
}

extern "C" int kvc_paged_attention_fused_metrics(
    void* out, float* kv_metrics, const void* query, const void* key_cache, const void* value_cache,
    int num_seqs, int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size, int max_context_len,
    int max_num_blocks_per_seq, int64_t q_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype, void* workspace, size_t workspace_bytes, kvc_stream_t stream) {
  AttnParams p;
  int rc = fill_params(p, out, kv_metrics, query, key_cache, value_cache, num_seqs, num_heads,
                       num_kv_heads, head_size, scale, block_tables, context_lens, kv_position,
                       last_position, kv_metric_buffer_len, block_size, max_num_blocks_per_seq,
                       q_stride, kv_block_stride, alibi_slopes, dtype, 1);
  if (rc != KVC_OK) return rc;
  KVC_REQUIRE(max_context_len > 0, KVC_ERR_INVALID, "fused attention: max_context_len <= 0");
  if (num_seqs == 0) return KVC_OK;
  const int padded = (max_context_len + kB - 1) / kB * kB;
  if (use_stream_kernel(num_seqs, num_kv_heads, max_context_len)) {
    const size_t need = kvc_paged_attention_fused_metrics_workspace_bytes(num_seqs, num_heads, max_context_len);
    KVC_REQUIRE(workspace && workspace_bytes >= need, KVC_ERR_WORKSPACE,
                "fused attention: workspace of %zu bytes required, got %zu", need, workspace_bytes);
    return dispatch_stream<kModeFused>(p, num_seqs, reinterpret_cast<float*>(workspace), padded, dtype,
                                       as_stream(stream));
  }
  return dispatch_attention<kModeFused>(p, num_seqs, 1, padded, dtype, as_stream(stream));
}

extern "C" size_t kvc_paged_attention_fused_metrics_workspace_bytes(int num_seqs, int num_heads,
                                                                    int max_context_len) {
  const int padded = (max_context_len + kB - 1) / kB * kB;
  if (num_seqs <= 0) return 0;
  return (size_t)num_seqs * num_heads * padded * sizeof(float);  // raw logits [S, Hq, padded]
}

extern "C" int kvc_paged_attention_fused_step(
    void* out, float* kv_metrics, const void* query, const void* key, const void* value, void* key_cache,
    void* value_cache, const int64_t* slot_mapping, const float* kv_metric_head_bias, int num_seqs,
    int num_heads, int num_kv_heads, int head_size, float scale, const int* block_tables,
    const int* context_lens, const int* kv_position, const int* last_position,
    const int* kv_metric_buffer_len, int block_size, int max_context_len, int max_num_blocks_per_seq,
    int64_t q_stride, int64_t key_stride, int64_t value_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype, void* workspace, size_t workspace_bytes, kvc_stream_t stream) {
  AttnParams p;
  int rc = fill_params(p, out, kv_metrics, query, key_cache, value_cache, num_seqs, num_heads,
                       num_kv_heads, head_size, scale, block_tables, context_lens, kv_position,
                       last_position, kv_metric_buffer_len, block_size, max_num_blocks_per_seq,
                       q_stride, kv_block_stride, alibi_slopes, dtype, 1);
  if (rc != KVC_OK) return rc;
  KVC_REQUIRE(key && value && slot_mapping && kv_metric_head_bias, KVC_ERR_INVALID,
              "fused attention step: null key / value / slot_mapping / head bias");
  KVC_REQUIRE(key_stride % 8 == 0 && value_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(key) & 15) == 0,
              KVC_ERR_INVALID, "fused attention step: key must be 16-byte aligned");
  KVC_REQUIRE(max_context_len > 0, KVC_ERR_INVALID, "fused attention step: max_context_len <= 0");
  if (num_seqs == 0) return KVC_OK;
  p.new_k = key; p.new_v = value; p.new_k_stride = key_stride; p.new_v_stride = value_stride;
  p.slot_mapping = slot_mapping; p.head_bias = kv_metric_head_bias;
  const int padded = (max_context_len + kB - 1) / kB * kB;
  const size_t need = kvc_paged_attention_fused_metrics_workspace_bytes(num_seqs, num_heads, max_context_len);
  KVC_REQUIRE(workspace && workspace_bytes >= need, KVC_ERR_WORKSPACE,
              "fused attention step: workspace of %zu bytes required, got %zu", need, workspace_bytes);
  return dispatch_stream<kModeFusedWrite>(p, num_seqs, reinterpret_cast<float*>(workspace), padded, dtype,
                                          as_stream(stream));
}

extern "C" int kvc_paged_attention_fused_step_deferred(
    void* out, float* kv_metrics, const void* query, const void* key, const void* value, void* key_cache,
    void* value_cache, const int64_t* slot_mapping, const float* kv_metric_head_bias, void* pend_kv,
    int64_t* pend_slot, int64_t* pend_dst, int num_seqs, int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position, const int* last_position,
    const int* kv_metric_buffer_len, int block_size, int max_context_len, int max_num_blocks_per_seq,
    int64_t q_stride, int64_t key_stride, int64_t value_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype, void* workspace, size_t workspace_bytes, kvc_stream_t stream) {
  AttnParams p;
  int rc = fill_params(p, out, kv_metrics, query, key_cache, value_cache, num_seqs, num_heads,
                       num_kv_heads, head_size, scale, block_tables, context_lens, kv_position,
                       last_position, kv_metric_buffer_len, block_size, max_num_blocks_per_seq,
                       q_stride, kv_block_stride, alibi_slopes, dtype, 1);
  if (rc != KVC_OK) return rc;
  KVC_REQUIRE(key && value && slot_mapping && kv_metric_head_bias, KVC_ERR_INVALID,
              "fused attention step: null key / value / slot_mapping / head bias");
  KVC_REQUIRE(pend_kv && pend_slot && (reinterpret_cast<uintptr_t>(pend_kv) & 15) == 0, KVC_ERR_INVALID,
              "fused attention step (deferred): pend_kv (16-byte aligned) and pend_slot are required");
  KVC_REQUIRE(key_stride % 8 == 0 && value_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(key) & 15) == 0,
              KVC_ERR_INVALID, "fused attention step: key must be 16-byte aligned");
  KVC_REQUIRE(max_context_len > 0, KVC_ERR_INVALID, "fused attention step: max_context_len <= 0");
  if (num_seqs == 0) return KVC_OK;
  p.new_k = key; p.new_v = value; p.new_k_stride = key_stride; p.new_v_stride = value_stride;
  p.slot_mapping = slot_mapping; p.head_bias = kv_metric_head_bias;
  p.pend_kv = reinterpret_cast<uint16_t*>(pend_kv); p.pend_slot = pend_slot; p.pend_dst = pend_dst;
  const int padded = (max_context_len + kB - 1) / kB * kB;
  const size_t need = kvc_paged_attention_fused_metrics_workspace_bytes(num_seqs, num_heads, max_context_len);
  KVC_REQUIRE(workspace && workspace_bytes >= need, KVC_ERR_WORKSPACE,
              "fused attention step: workspace of %zu bytes required, got %zu", need, workspace_bytes);
  if (pend_dst != nullptr)
    return dispatch_stream<kModeFusedWriteMove>(p, num_seqs, reinterpret_cast<float*>(workspace), padded, dtype,
                                                as_stream(stream));
  return dispatch_stream<kModeFusedWrite>(p, num_seqs, reinterpret_cast<float*>(workspace), padded, dtype,
                                          as_stream(stream));
}

#ifdef KVC_ATTN_TRACE
// diagnostic builds only (tools/attn_trace.py): timestamps of every consumer warp of the NEXT launches
extern "C" __attribute__((visibility("default"))) int kvc_debug_set_attention_trace(unsigned long long* buf) {
  kvc::g_attn_trace = buf;
  return 0;
}
#endif

extern "C" int kvc_set_attention_pdl(int mode) {
  KVC_REQUIRE(mode >= 0 && mode <= 3, KVC_ERR_INVALID, "attention PDL mode must be 0, 1, 2 or 3, got %d", mode);
  g_pdl_mode = mode;
  return KVC_OK;
}
extern "C" int kvc_get_attention_pdl(void) { return attention_pdl_mode(); }

extern "C" int kvc_set_attention_v2_mode(int mode) {
  KVC_REQUIRE(mode >= 0 && mode <= 2, KVC_ERR_INVALID, "attention v2 mode must be 0, 1 or 2, got %d", mode);
  g_v2_mode = mode;
  return KVC_OK;
}
extern "C" int kvc_get_attention_v2_mode(void) { attention_v2_mode(); return g_v2_mode; }
