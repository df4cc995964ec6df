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
#include "kvc_common.cuh"

namespace kvc {

constexpr int kAttnWarps = 4;
constexpr int kAttnThreads = kAttnWarps * kWarp;
constexpr int kD = 128;
constexpr int kB = 16;
constexpr int kPartition = 512;  // reference PARTITION_SIZE, :866
constexpr int kOStride = 132;    // padded d-stride of the cross-warp merge buffer (bank-conflict free)

enum AttnMode { kModeV1 = 0, kModeV2 = 1, kModeFused = 2 };

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
};

template <bool BF16>
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (BF16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
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
// v2 reduce.  grid (num_kv_heads, num_seqs), 128 threads.  reference :525-652.
// -----------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kAttnThreads) paged_attention_reduce_kernel(AttnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* wts = reinterpret_cast<float*>(smem_raw);  // [max_partitions] rescaled exp sums
  __shared__ float s_M, s_inv;
  const int h = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const int H = p.num_kv_heads, Q = p.Q;
  const int seq_head = s * H + h;
  const int ctx = p.context_lens[seq_head];
  const int nparts = (ctx + kPartition - 1) / kPartition;
  const int* __restrict__ bt = p.block_tables + (int64_t)seq_head * p.max_num_blocks_per_seq;
  const int max_pos = p.last_position[s] - p.kv_metric_buffer_len[s];
  for (int q = 0; q < Q; ++q) {
    const int qh = h * Q + q;
    const int64_t base = ((int64_t)s * p.num_heads + qh) * p.max_partitions;
    __syncthreads();
    if (tid == 0) {
      float M = -INFINITY;
      for (int j = 0; j < nparts; ++j) M = fmaxf(M, p.max_logits[base + j]);
      float L = 0.f;
      for (int j = 0; j < nparts; ++j) {
        const float r = p.exp_sums[base + j] * expf(p.max_logits[base + j] - M);  // :606-611
        wts[j] = r;
        L += r;
      }
      s_M = M;
      s_inv = __fdividef(1.f, L + 1e-6f);  // :614
    }
    __syncthreads();
    {
      const int d = tid;
      float acc = 0.f;
      const uint16_t* tp = reinterpret_cast<const uint16_t*>(p.tmp_out) + base * kD;
      if (nparts == 1) {
        acc = from_bits16<BF16>(tp[d]);  // reference :555-562: plain copy
      } else {
        for (int j = 0; j < nparts; ++j)
          acc += from_bits16<BF16>(tp[(int64_t)j * kD + d]) * wts[j] * s_inv;  // :625
      }
      reinterpret_cast<uint16_t*>(p.out)[((int64_t)s * p.num_heads + qh) * kD + d] = to_bits16<BF16>(acc);
    }
    if (p.record && nparts > 0) {
      const float M = s_M, inv = s_inv;
      for (int t = tid; t < ctx; t += kAttnThreads) {
        const int64_t slot = (int64_t)bt[t / kB] * kB + t % kB;
        if (p.kv_position[slot] <= max_pos) {
          const float lg = p.metric[slot * Q + q];
          p.metric[slot * Q + q] = __expf(lg - M) * inv;
        }
      }
    }
  }
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
  p.record = record ? 1 : 0; p.max_partitions = 1;
  return KVC_OK;
}

}  // namespace kvc

using namespace kvc;

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
  rc = dispatch_attention<kModeV2>(p, num_seqs, p.max_partitions, 0, dtype, st);
  if (rc != KVC_OK) return rc;
  const size_t smem = sizeof(float) * (size_t)p.max_partitions;
  if (dtype == KVC_F16)
    paged_attention_reduce_kernel<false><<<dim3(num_kv_heads, num_seqs), kAttnThreads, smem, st>>>(p);
  else
    paged_attention_reduce_kernel<true><<<dim3(num_kv_heads, num_seqs), kAttnThreads, smem, st>>>(p);
  KVC_CHECK_LAUNCH("paged_attention_reduce_kernel");
  return KVC_OK;
}

extern "C" int kvc_paged_attention_fused_metrics(
    void* out, float* kv_metrics, const void* query, const void* key_cache, const void* value_cache,
    int num_seqs, int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size, int max_context_len,
    int max_num_blocks_per_seq, int64_t q_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype, kvc_stream_t stream) {
  AttnParams p;
  int rc = fill_params(p, out, kv_metrics, query, key_cache, value_cache, num_seqs, num_heads,
                       num_kv_heads, head_size, scale, block_tables, context_lens, kv_position,
                       last_position, kv_metric_buffer_len, block_size, max_num_blocks_per_seq,
                       q_stride, kv_block_stride, alibi_slopes, dtype, 1);
  if (rc != KVC_OK) return rc;
  KVC_REQUIRE(max_context_len > 0, KVC_ERR_INVALID, "fused attention: max_context_len <= 0");
  if (num_seqs == 0) return KVC_OK;
  const int padded = (max_context_len + kB - 1) / kB * kB;
  return dispatch_attention<kModeFused>(p, num_seqs, 1, padded, dtype, as_stream(stream));
}
