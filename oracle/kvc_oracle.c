/*
 * kvc_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the KV-Compress hot path of
 * IsaacRe/vllm-kvcompress (reference @ 5e2a639).  It exists so that the CUDA
 * product path (vllm-kvcompress_b200/csrc) can be checked bit-for-bit on the
 * same inputs, and so that bench.py can time "the reference algorithm on the
 * host cores" next to the GPU numbers.  Nothing in the product path may link,
 * import or call this file: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 8c).  This restatement is pinned against
 *   (1) tests/golden/*.npz -- outputs of the reference's own pure-Python
 *       oracles (vllm/_custom_ops.py ref_schedule_cache_evictions,
 *       ref_count_block_evictions, ref_schedule_t1_cache_moves,
 *       ref_execute_cache_moves and tests/kernels/test_kvcompress_attention.py
 *       ref_single_query_cached_kv_attention) executed in the build container
 *       by tests/golden/make_golden.py, and
 *   (2) on the GPU box, the reference's own CUDA kernels compiled unmodified
 *       into oracle/_ref/_kvc_ref.so (oracle/Makefile.ref).
 *
 * One reference CUDA thread == one loop iteration here; loops over
 * independent heads / sequences carry an OpenMP pragma so the CPU baseline
 * can use every host core (compile with -fopenmp; without it the file is a
 * scalar port).
 *
 * All index tensors are int32 ("int") exactly as the reference kernels
 * reinterpret them (csrc/kvcompress_eviction_kernels.cu:527-542).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define KVC_EXPORT __attribute__((visibility("default")))

KVC_EXPORT int kvc_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the benchmark's CPU arm sets the thread count itself */
KVC_EXPORT void kvc_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------
 * K2  truncate_cache_evictions_kernel
 * reference: csrc/kvcompress_eviction_kernels.cu:158-188
 * one reference thread per global head g in [0, S*L*H).
 * ---------------------------------------------------------------------- */
KVC_EXPORT void kvc_oracle_truncate_cache_evictions(
    int* evicted_kv_indices,        /* [max_evicted_kv] */
    int* evicted_logical_indices,   /* [max_evicted_kv] */
    int* evicted_kv_count,          /* [S, L, H] */
    const int* evicted_kv_offsets,  /* [S, L, H] */
    const int* hanging_token_count, /* [S, L, H] */
    int total_heads, int block_size, int max_evicted_kv,
    int null_eviction_index) {
#pragma omp parallel for schedule(static)
  for (int g = 0; g < total_heads; ++g) {
    const int evicted_kv = evicted_kv_count[g];
    const int hanging_kv = hanging_token_count[g];
    int truncated;
    if (evicted_kv < hanging_kv) {
      truncated = 0; /* :174-175 */
    } else {
      truncated = evicted_kv - (evicted_kv - hanging_kv) % block_size; /* :177 */
    }
    evicted_kv_count[g] = truncated;
    const int start = evicted_kv_offsets[g] + truncated; /* :182 */
    const int end = (g + 1 >= total_heads) ? max_evicted_kv
                                           : evicted_kv_offsets[g + 1]; /* :183 */
    for (int i = start; i < end; ++i) {
      evicted_kv_indices[i] = g;
      evicted_logical_indices[i] = null_eviction_index;
    }
  }
}

/* ------------------------------------------------------------------------
 * K1  schedule_cache_evictions_kernel (+ K2 when truncate != 0)
 * reference: csrc/kvcompress_eviction_kernels.cu:27-156, launcher :492-555
 * Python twin: vllm/_custom_ops.py:774-931 (ref_schedule_cache_evictions)
 *
 * One reference thread per sequence (per (sequence, layer) when
 * evict_evenly_per_layer).  context_lens / hanging_token_count are [S, L, H]
 * (the wrapper transposes context_lens, _custom_ops.py:995).  block indices in
 * sorted_indices address the *masked* block arrays (layer_by_block etc.), not
 * physical cache blocks.
 *
 * Note on evict_evenly_per_layer: in the reference every (seq, layer) thread
 * zeroes the counts of all heads of its sequence before walking (:80-83),
 * which races with the other layers' increments; the intended (and, for a
 * single warp in lock-step, the observed) behaviour is "zero first, then walk
 * each non-control layer independently with budget E / L".  That is what is
 * restated here.
 * ---------------------------------------------------------------------- */
static void k1_walk(int* evicted_kv_indices, int* evicted_logical_indices,
                    int* evicted_kv_count, int* remaining_kv_count,
                    const int* evicted_kv_offsets, const int* sorted_indices,
                    const int* seq_block_offsets, const int* layer_by_block,
                    const int* head_by_block,
                    const int* virtual_block_num_by_block,
                    const int* context_lens, const int* kv_position,
                    const int* last_position, const int* protected_window_size,
                    int num_seqs, int num_layers, int num_kv_heads,
                    int total_blocks, int block_size, int seq_idx,
                    int layer_filter /* -1 = all layers */, int blocks_to_evict) {
  const int max_evictable_position =
      last_position[seq_idx] - protected_window_size[seq_idx]; /* :63 */
  const int output_seq_stride = num_layers * num_kv_heads;
  const int output_head_offset = seq_idx * output_seq_stride;
  const int seq_start = seq_block_offsets[seq_idx];
  const int seq_end = (seq_idx + 1 >= num_seqs) ? total_blocks
                                                : seq_block_offsets[seq_idx + 1]; /* :73 */
  if (blocks_to_evict == 0) return; /* :85-87 */
  const int* seq_sorted = sorted_indices + (int64_t)seq_start * block_size;
  const int seq_total_tokens = (seq_end - seq_start) * block_size;
  int evicted_block_count = 0;
  for (int i = 0; i < seq_total_tokens; ++i) { /* :117 */
    const int token_idx = seq_sorted[i];
    const int block_idx = token_idx / block_size;
    const int layer_idx = layer_by_block[block_idx];
    if (layer_filter >= 0 && layer_idx != layer_filter) continue; /* :121-123 */
    const int head_idx = head_by_block[block_idx];
    const int virtual_block_num = virtual_block_num_by_block[block_idx];
    const int block_offset = token_idx % block_size;
    const int virtual_token_idx = virtual_block_num * block_size + block_offset;
    const int layer_head_idx = layer_idx * num_kv_heads + head_idx;
    const int output_kv_idx = output_head_offset + layer_head_idx;
    /* skip empty slots and the protected window, :136-139 */
    if (virtual_token_idx >= context_lens[output_kv_idx] ||
        kv_position[token_idx] > max_evictable_position) {
      continue;
    }
    const int off = evicted_kv_offsets[output_kv_idx] + evicted_kv_count[output_kv_idx]++;
    evicted_kv_indices[off] = output_kv_idx;       /* :144 */
    evicted_logical_indices[off] = virtual_token_idx; /* :145 */
    if (--remaining_kv_count[output_kv_idx] == 0) { /* :149 */
      if (++evicted_block_count >= blocks_to_evict) return;
      remaining_kv_count[output_kv_idx] = block_size;
    }
  }
}

KVC_EXPORT void kvc_oracle_schedule_cache_evictions(
    int* evicted_kv_indices,            /* [max_evicted_kv] */
    int* evicted_logical_indices,       /* [max_evicted_kv] */
    int* evicted_kv_count,              /* [S, L, H] */
    int* remaining_kv_count,            /* [S, L, H] in/out scratch (clone of hanging) */
    const int* evicted_kv_offsets,      /* [S, L, H] */
    const int* sorted_indices,          /* [total_blocks * B] */
    const int* seq_block_offsets,       /* [S] */
    const int* layer_by_block,          /* [total_blocks] */
    const int* head_by_block,           /* [total_blocks] */
    const int* virtual_block_num_by_block, /* [total_blocks] */
    const int* evicted_blocks_per_seq,  /* [S] */
    const int* context_lens,            /* [S, L, H] */
    const int* hanging_token_count,     /* [S, L, H] */
    const int* kv_position,             /* [total_blocks, B] */
    const int* last_position,           /* [S] */
    const int* protected_window_size,   /* [S] */
    int num_seqs, int num_layers, int num_kv_heads, int total_blocks,
    int block_size, int evict_evenly_per_layer, const int* control_layers,
    int num_control_layers, int max_evicted_kv, int null_eviction_index,
    int truncate) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int s = 0; s < num_seqs; ++s) {
    const int stride = num_layers * num_kv_heads;
    for (int i = 0; i < stride; ++i) evicted_kv_count[s * stride + i] = 0; /* :80-83 */
    if (!evict_evenly_per_layer) {
      k1_walk(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
              remaining_kv_count, evicted_kv_offsets, sorted_indices,
              seq_block_offsets, layer_by_block, head_by_block,
              virtual_block_num_by_block, context_lens, kv_position,
              last_position, protected_window_size, num_seqs, num_layers,
              num_kv_heads, total_blocks, block_size, s, -1,
              evicted_blocks_per_seq[s]);
    } else {
      for (int l = 0; l < num_layers; ++l) {
        int is_control = 0;
        for (int c = 0; c < num_control_layers; ++c)
          if (control_layers[c] == l) is_control = 1; /* :55-61 */
        if (is_control) continue;
        k1_walk(evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
                remaining_kv_count, evicted_kv_offsets, sorted_indices,
                seq_block_offsets, layer_by_block, head_by_block,
                virtual_block_num_by_block, context_lens, kv_position,
                last_position, protected_window_size, num_seqs, num_layers,
                num_kv_heads, total_blocks, block_size, s, l,
                evicted_blocks_per_seq[s] / num_layers /* :76-78 */);
      }
    }
  }
  if (truncate) { /* launcher macro :463-472 */
    kvc_oracle_truncate_cache_evictions(
        evicted_kv_indices, evicted_logical_indices, evicted_kv_count,
        evicted_kv_offsets, hanging_token_count,
        num_seqs * num_layers * num_kv_heads, block_size, max_evicted_kv,
        null_eviction_index);
  }
}

/* ------------------------------------------------------------------------
 * K3  count_block_evictions_kernel
 * reference: csrc/kvcompress_eviction_kernels.cu:190-221, launcher :640-664
 * Python twin: vllm/_custom_ops.py:1022-1062
 * ---------------------------------------------------------------------- */
KVC_EXPORT void kvc_oracle_count_block_evictions(
    int* evicted_block_count,       /* [S, L, H] */
    int* evicted_logical_indices,   /* [total_kvs] in/out */
    const int* evicted_kv_offsets,  /* [S, L, H] */
    const int* hanging_token_count, /* [S, L, H] */
    int total_heads, int total_kvs, int block_size, int null_value) {
#pragma omp parallel for schedule(static)
  for (int g = 0; g < total_heads; ++g) {
    const int start = evicted_kv_offsets[g];
    const int end = (g + 1 >= total_heads) ? total_kvs : evicted_kv_offsets[g + 1];
    int evicted_blocks = 0;
    for (int i = start; i < end; i += block_size) { /* :201-207 */
      if (evicted_logical_indices[i] != null_value) {
        evicted_blocks++;
      } else {
        break;
      }
    }
    evicted_block_count[g] = evicted_blocks;
    if (evicted_blocks > 0) { /* :210-219 */
      const int last_block_end = start + evicted_blocks * block_size;
      for (int i = last_block_end - block_size + hanging_token_count[g];
           i < last_block_end; ++i) {
        evicted_logical_indices[i] = null_value;
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * K4  single_tier_schedule_cache_moves_kernel
 * reference: csrc/kvcompress_eviction_kernels.cu:223-289, launcher :698-726
 * Python twin: vllm/_custom_ops.py:1108-1154
 * block_tables [L, S, H, MB], context_lens [L, S, H]; everything else [S, L, H].
 * ---------------------------------------------------------------------- */
KVC_EXPORT void kvc_oracle_schedule_t1_cache_moves(
    int* cache_moves_idx,               /* [max_evicted_kv, 2] */
    int* cache_moves_count,             /* [S, L, H] */
    const int* evicted_logical_indices, /* [max_evicted_kv] */
    const int* evicted_kv_count,        /* [S, L, H] */
    const int* evicted_kv_offsets,      /* [S, L, H] */
    const int* block_tables,            /* [L, S, H, MB] */
    const int* context_lens,            /* [L, S, H] */
    int num_seqs, int num_layers, int num_kv_heads, int max_num_blocks_per_seq,
    int block_size) {
  const int total = num_seqs * num_layers * num_kv_heads;
#pragma omp parallel for schedule(dynamic, 8)
  for (int g = 0; g < total; ++g) {
    const int seq_idx = g / (num_layers * num_kv_heads);
    const int layer_idx = (g / num_kv_heads) % num_layers;
    const int head_idx = g % num_kv_heads;
    const int lsh = (layer_idx * num_seqs + seq_idx) * num_kv_heads + head_idx;
    const int64_t bt_off = (int64_t)lsh * max_num_blocks_per_seq;
    const int off = evicted_kv_offsets[g];
    const int evicted_kv_cnt = evicted_kv_count[g];
    int move_count = 0, evict_count = 0, done = 0;
    for (int i = 0; i < evicted_kv_cnt; ++i) { /* :252 */
      const int src_kv_idx = context_lens[lsh] - 1 - i;
      const int src_kv_stop_idx =
          evicted_logical_indices[off + evicted_kv_cnt - 1 - evict_count];
      const int dst_kv_idx = evicted_logical_indices[off + move_count];
      if (dst_kv_idx >= src_kv_idx) { /* :260-263 */
        cache_moves_count[g] = move_count;
        done = 1;
        break;
      }
      if (src_kv_idx <= src_kv_stop_idx) { /* :265-268 */
        ++evict_count;
        continue;
      }
      const int src_blk = block_tables[bt_off + src_kv_idx / block_size];
      const int src_phys = src_blk * block_size + src_kv_idx % block_size;
      const int dst_blk = block_tables[bt_off + dst_kv_idx / block_size];
      const int dst_phys = dst_blk * block_size + dst_kv_idx % block_size;
      const int64_t p = ((int64_t)off + (move_count++)) * 2; /* :280 */
      cache_moves_idx[p] = dst_phys;
      cache_moves_idx[p + 1] = src_phys;
    }
    if (!done) cache_moves_count[g] = move_count; /* :286 */
  }
}

/* ------------------------------------------------------------------------
 * K5  execute_cache_moves_kernel
 * reference: csrc/kvcompress_eviction_kernels.cu:359-435, launcher :858-929
 * Python twin: vllm/_custom_ops.py:1182-1216
 * Bit copy; cache element width `elem_size` bytes (1, 2 or 4).
 *   k_cache [NB, D/x, B, x], v_cache [NB, D, B], kv_metrics [NB, B] f32,
 *   kv_position [NB, B] i32.
 * Moves of one head are applied in list order; heads are independent
 * (the reference requires disjoint dst across heads, :358).
 * ---------------------------------------------------------------------- */
KVC_EXPORT void kvc_oracle_execute_cache_moves(
    void* k_cache, void* v_cache, float* kv_metrics, int* kv_position,
    const int* cache_moves_idx,    /* [max_evicted_kv, 2] (dst, src) */
    const int* cache_moves_count,  /* [S, L, H] */
    const int* evicted_kv_offsets, /* [S, L, H] */
    int total_heads, int head_size, int vec_size /* x */, int block_size,
    int elem_size) {
  uint8_t* kc = (uint8_t*)k_cache;
  uint8_t* vc = (uint8_t*)v_cache;
  const int64_t block_stride = (int64_t)block_size * head_size; /* :378 */
  const int64_t k_stride = (int64_t)block_size * vec_size;      /* :379 */
#pragma omp parallel for schedule(dynamic, 4)
  for (int g = 0; g < total_heads; ++g) {
    const int cnt = cache_moves_count[g];
    const int off = evicted_kv_offsets[g];
    for (int i = off; i < off + cnt; ++i) {
      const int dst_idx = cache_moves_idx[2 * (int64_t)i];
      const int src_idx = cache_moves_idx[2 * (int64_t)i + 1];
      const int64_t src_block_start = (int64_t)(src_idx / block_size) * block_stride;
      const int src_o = src_idx % block_size;
      const int64_t dst_block_start = (int64_t)(dst_idx / block_size) * block_stride;
      const int dst_o = dst_idx % block_size;
      kv_metrics[dst_idx] = kv_metrics[src_idx];   /* :410 */
      kv_position[dst_idx] = kv_position[src_idx]; /* :411 */
      /* K: D/x groups of x contiguous elements at stride B*x, :413-424 */
      for (int64_t j = 0; j < block_stride; j += k_stride) {
        memcpy(kc + (dst_block_start + (int64_t)dst_o * vec_size + j) * elem_size,
               kc + (src_block_start + (int64_t)src_o * vec_size + j) * elem_size,
               (size_t)vec_size * elem_size);
      }
      /* V: D single elements at stride B, :426-431 */
      for (int64_t j = 0; j < block_stride; j += block_size) {
        memcpy(vc + (dst_block_start + dst_o + j) * elem_size,
               vc + (src_block_start + src_o + j) * elem_size, (size_t)elem_size);
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * K6  single_tier_reshape_and_cache_kernel  (kv_cache_dtype == "auto" only)
 * reference: csrc/kvcompress_cache_kernels.cu:27-89, launcher :111-139
 * key/value [T, H, D] with token strides key_stride/value_stride (elements);
 * slot_mapping [T, H] int64 (<0 => padding, skipped).
 * ---------------------------------------------------------------------- */
KVC_EXPORT void kvc_oracle_reshape_and_cache(
    const void* key, const void* value, void* key_cache, void* value_cache,
    float* kv_metrics, const int64_t* slot_mapping,
    const float* kv_metric_head_bias, int num_tokens, int num_heads,
    int head_size, int block_size, int x, int64_t key_stride,
    int64_t value_stride, int elem_size) {
  const uint8_t* k = (const uint8_t*)key;
  const uint8_t* v = (const uint8_t*)value;
  uint8_t* kc = (uint8_t*)key_cache;
  uint8_t* vc = (uint8_t*)value_cache;
#pragma omp parallel for schedule(static)
  for (int t = 0; t < num_tokens; ++t) {
    for (int i = 0; i < num_heads * head_size; ++i) {
      const int head_idx = i / head_size;
      const int64_t slot_idx = slot_mapping[(int64_t)t * num_heads + head_idx];
      if (slot_idx < 0) continue; /* :49-52 */
      if (i % head_size == 0) kv_metrics[slot_idx] = kv_metric_head_bias[head_idx]; /* :55-57 */
      const int64_t block_idx = slot_idx / block_size;
      const int64_t block_offset = slot_idx % block_size;
      const int head_offset = i % head_size;
      const int x_idx = head_offset / x;
      const int x_offset = head_offset % x;
      const int64_t tgt_k = block_idx * (head_size / x) * block_size * x +
                            (int64_t)x_idx * block_size * x + block_offset * x + x_offset; /* :69-72 */
      const int64_t tgt_v = block_idx * head_size * block_size +
                            (int64_t)head_offset * block_size + block_offset; /* :73-75 */
      memcpy(kc + tgt_k * elem_size, k + ((int64_t)t * key_stride + i) * elem_size, (size_t)elem_size);
      memcpy(vc + tgt_v * elem_size, v + ((int64_t)t * value_stride + i) * elem_size, (size_t)elem_size);
    }
  }
}

/* ------------------------------------------------------------------------
 * K7/K8  KV-Compress paged attention, dense fp32 restatement.
 * reference arithmetic: csrc/attention/kvcompress_attention_kernels.cu
 *   :117-147 (per-(seq, kv head) context and block table), :207-314 (QK^T,
 *   masking, softmax with 1/(sum + 1e-6), metric write rule), :329-450 (PV),
 *   :553-651 (v2 reduce == same mathematical result);
 * test twin: tests/kernels/test_kvcompress_attention.py:41-145.
 * 16-bit inputs (fp16 when !is_bf16) are widened to fp32; everything is
 * accumulated in fp32/double, so this is *more* accurate than the reference
 * kernel (which multiplies P and V in half2, dtype_float16.cuh:433-457); the
 * parity budget is atol 1e-3 (test_kvcompress_attention.py:356-357).
 *
 * Metric rule (:124, :305-313): p is written to
 * kv_metric_out[slot * Q + q_off] iff record_kv_metrics and
 * kv_position[slot] <= last_position[s] - kv_metric_buffer_len[s].
 * ---------------------------------------------------------------------- */
static float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t f;
  if (exp == 0) {
    if (man == 0) {
      f = sign;
    } else { /* subnormal */
      exp = 127 - 15 + 1;
      while ((man & 0x400u) == 0) {
        man <<= 1;
        exp--;
      }
      man &= 0x3ffu;
      f = sign | (exp << 23) | (man << 13);
    }
  } else if (exp == 31) {
    f = sign | 0x7f800000u | (man << 13);
  } else {
    f = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float out;
  memcpy(&out, &f, 4);
  return out;
}
static float bf16_to_float(uint16_t h) {
  uint32_t f = (uint32_t)h << 16;
  float out;
  memcpy(&out, &f, 4);
  return out;
}
static inline float to_f32(uint16_t h, int is_bf16) {
  return is_bf16 ? bf16_to_float(h) : half_to_float(h);
}

KVC_EXPORT void kvc_oracle_paged_attention(
    float* out,                  /* [S, Hq, D] fp32 */
    float* kv_metric_out,        /* [NB, B, Q] fp32, only touched where the rule fires */
    const uint16_t* query,       /* [S, Hq, D] with row stride q_stride (elements) */
    const uint16_t* key_cache,   /* [NB, D/x, B, x] */
    const uint16_t* value_cache, /* [NB, D, B] */
    const int* block_tables,     /* [S, H, MB] */
    const int* context_lens,     /* [S, H] */
    const int* kv_position,      /* [NB, B] */
    const int* last_position,    /* [S] */
    const int* kv_metric_buffer_len, /* [S] */
    const float* alibi_slopes,   /* [Hq] or NULL */
    int num_seqs, int num_heads, int num_kv_heads, int head_size,
    int block_size, int x, int max_num_blocks_per_seq, int64_t q_stride,
    float scale, int is_bf16, int record_kv_metrics) {
  const int Q = num_heads / num_kv_heads;
  const int64_t blk_stride = (int64_t)block_size * head_size;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int s = 0; s < num_seqs; ++s) {
    for (int qh = 0; qh < num_heads; ++qh) {
      const int h = qh / Q;
      const int q_off = qh % Q;
      const int ctx = context_lens[s * num_kv_heads + h];
      const int* bt = block_tables + ((int64_t)s * num_kv_heads + h) * max_num_blocks_per_seq;
      const int max_pos = last_position[s] - kv_metric_buffer_len[s];
      const float slope = alibi_slopes ? alibi_slopes[qh] : 0.f;
      float* o = out + ((int64_t)s * num_heads + qh) * head_size;
      for (int d = 0; d < head_size; ++d) o[d] = 0.f;
      if (ctx <= 0) continue;
      float* logits = (float*)malloc(sizeof(float) * (size_t)ctx);
      double* acc = (double*)calloc((size_t)head_size, sizeof(double));
      const uint16_t* q = query + (int64_t)s * q_stride + (int64_t)qh * head_size;
      float qk_max = -3.402823466e+38f;
      for (int t = 0; t < ctx; ++t) {
        const int64_t blk = bt[t / block_size];
        const int o_in = t % block_size;
        double dot = 0.0;
        for (int d = 0; d < head_size; ++d) {
          const int64_t ki = blk * blk_stride + (int64_t)(d / x) * block_size * x + (int64_t)o_in * x + d % x;
          dot += (double)to_f32(q[d], is_bf16) * (double)to_f32(key_cache[ki], is_bf16);
        }
        float qk = scale * (float)dot;
        if (slope != 0.f) qk += slope * (float)(t - ctx + 1); /* :253 */
        logits[t] = qk;
        if (qk > qk_max) qk_max = qk;
      }
      double exp_sum = 0.0;
      for (int t = 0; t < ctx; ++t) {
        logits[t] = expf(logits[t] - qk_max);
        exp_sum += logits[t];
      }
      const float inv_sum = 1.f / ((float)exp_sum + 1e-6f); /* :298 */
      for (int t = 0; t < ctx; ++t) {
        const int64_t blk = bt[t / block_size];
        const int o_in = t % block_size;
        const float p = logits[t] * inv_sum;
        const int64_t slot = blk * block_size + o_in;
        if (record_kv_metrics && kv_position[slot] <= max_pos) /* :305-313 */
          kv_metric_out[slot * Q + q_off] = p;
        for (int d = 0; d < head_size; ++d) {
          const int64_t vi = blk * blk_stride + (int64_t)d * block_size + o_in;
          acc[d] += (double)p * (double)to_f32(value_cache[vi], is_bf16);
        }
      }
      for (int d = 0; d < head_size; ++d) o[d] = (float)acc[d];
      free(logits);
      free(acc);
    }
  }
}

/* ------------------------------------------------------------------------
 * Decode-time metric aggregation (reference: vllm/kvcompress/metrics.py:337-342
 * clear_temp_metrics + :429-439 aggregate_decode with use_l2):
 *   metrics[slot] += sum_q temp[slot, q]^2   for every slot of the cache,
 * where temp was zeroed before the forward pass, so only slots written by the
 * attention kernels this step contribute.  The fused CUDA op accumulates the
 * same sum for the slots it visits; this is the dense check.
 * ---------------------------------------------------------------------- */
KVC_EXPORT void kvc_oracle_aggregate_decode(float* metrics, const float* temp_metrics,
                                            int64_t num_slots, int Q, int use_l2) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < num_slots; ++i) {
    float s = 0.f;
    for (int q = 0; q < Q; ++q) {
      const float t = temp_metrics[i * Q + q];
      s += use_l2 ? t * t : t;
    }
    metrics[i] += s;
  }
}

/* ------------------------------------------------------------------------
 * Prefill KV-metric initialisation
 * reference: vllm/attention/backends/flash_attn.py:1122-1211
 *   _naive_kvc_attention (:1122-1164): per prompt, queries = the last
 *   min(prompt_len, n_observed) tokens, processed in chunks of
 *   max_observed_block_size; output[start:end] = sum over chunks of the
 *   chunk's metrics.
 *   _naive_kvc_masked_attention (:1167-1211): logits = scale *
 *   float(einsum in the INPUT dtype) (so the dot product is rounded to
 *   fp16/bf16 first, :1178); causal mask triu(diagonal=q_offset+1) (:1176-1179);
 *   fp32 softmax over the keys (:1180); squared iff use_l2 (:1181-1182);
 *   metric mask tril(diagonal=q_offset-kv_metric_buffer_len) (:1184);
 *   sum over the chunk's queries (:1186); use_average: * (k+1)/n_queries
 *   (:1187-1195); use_maxpool: max_pool1d(kernel 7, padding 3) (:1196-1202).
 * and CompressionMetrics.aggregate_prefill, vllm/kvcompress/metrics.py:395-427.
 * Here the key tensor is [T, H, D] (the reference repeats it to Hq heads
 * first, flash_attn.py:988-991: query head hq reads kv head hq / Q).
 * ---------------------------------------------------------------------- */
static float round_half(float x) { return (float)(_Float16)x; }
static float round_bf16(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u += 0x7fffu + ((u >> 16) & 1u); /* RNE */
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

KVC_EXPORT void kvc_oracle_prefill_kv_metrics(
    float* kv_metric_out,      /* [T, Hq] */
    const uint16_t* query,     /* [T, Hq, D] */
    const uint16_t* key,       /* [T, H, D] */
    const int* seq_start_loc,  /* [num_seqs + 1] */
    const int* kv_metric_buffer_len, /* [num_seqs] */
    int T, int num_seqs, int Hq, int H, int D, float scale, int n_observed,
    int max_observed_block_size, int use_l2, int use_average, int use_maxpool,
    int is_bf16) {
  const int Q = Hq / H;
  (void)T;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int i = 0; i < num_seqs; ++i) {
    for (int hq = 0; hq < Hq; ++hq) {
      const int start = seq_start_loc[i], end = seq_start_loc[i + 1];
      const int P = end - start;
      if (P <= 0) continue;
      const int h = hq / Q;
      const int nobs_total = P < n_observed ? P : n_observed;
      const int start_trunc = end - nobs_total;
      float* logits = (float*)malloc(sizeof(float) * (size_t)P);
      float* chunk = (float*)malloc(sizeof(float) * (size_t)P);
      float* pooled = (float*)malloc(sizeof(float) * (size_t)P);
      for (int k = 0; k < P; ++k) kv_metric_out[(size_t)(start + k) * Hq + hq] = 0.f;
      for (int l = start_trunc; l < end; l += max_observed_block_size) {
        const int lend = l + max_observed_block_size < end ? l + max_observed_block_size : end;
        const int nq = lend - l;
        const int q_offset = l - start;
        for (int k = 0; k < P; ++k) chunk[k] = 0.f;
        for (int qi = 0; qi < nq; ++qi) {
          const uint16_t* qv = query + ((size_t)(l + qi) * Hq + hq) * D;
          float mx = -INFINITY;
          for (int k = 0; k < P; ++k) {
            const uint16_t* kv = key + ((size_t)(start + k) * H + h) * D;
            float acc = 0.f;
            for (int d = 0; d < D; ++d) acc += to_f32(qv[d], is_bf16) * to_f32(kv[d], is_bf16);
            float s = scale * (is_bf16 ? round_bf16(acc) : round_half(acc));
            /* additive mask of finfo(dtype).min: the softmax weight underflows to exactly 0 */
            if (k - qi >= q_offset + 1) s = -INFINITY;
            logits[k] = s;
            if (s > mx) mx = s;
          }
          float sum = 0.f;
          for (int k = 0; k < P; ++k) {
            logits[k] = expf(logits[k] - mx);
            sum += logits[k];
          }
          for (int k = 0; k < P; ++k) {
            float pr = logits[k] / sum;
            if (use_l2) pr = pr * pr;
            if (k - qi <= q_offset - kv_metric_buffer_len[i]) chunk[k] += pr;
          }
        }
        if (use_average)
          for (int k = 0; k < P; ++k) chunk[k] *= (float)(k + 1) / (float)nq;
        for (int k = 0; k < P; ++k) {
          float v = chunk[k];
          if (use_maxpool) {
            const int j0 = k - 3 < 0 ? 0 : k - 3, j1 = k + 3 > P - 1 ? P - 1 : k + 3;
            for (int j = j0; j <= j1; ++j) v = chunk[j] > v ? chunk[j] : v;
          }
          pooled[k] = v;
        }
        for (int k = 0; k < P; ++k) kv_metric_out[(size_t)(start + k) * Hq + hq] += pooled[k];
      }
      free(logits);
      free(chunk);
      free(pooled);
    }
  }
}

/* CompressionMetrics.aggregate_prefill, vllm/kvcompress/metrics.py:395-427 */
KVC_EXPORT void kvc_oracle_aggregate_prefill(float* metrics, const float* prefill_metrics,
                                             const int64_t* slot_mapping, int T, int Hq, int H) {
  const int Q = Hq / H;
  for (int t = 0; t < T; ++t)
    for (int h = 0; h < H; ++h) {
      float s = 0.f;
      for (int q = 0; q < Q; ++q) s += prefill_metrics[(size_t)t * Hq + h * Q + q];
      const int64_t slot = slot_mapping[(size_t)t * H + h];
      if (slot >= 0) metrics[slot] += s;
    }
}
