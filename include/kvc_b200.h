/*
 * kvc_b200.h -- C ABI of libkvc_b200.so: the B200-native (sm_100a) KV-Compress
 * eviction / compaction / paged-attention hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream; no torch types.  Each one replaces exactly one op of
 * the reference's torch C++ extension (csrc/torch_bindings.cpp) and keeps that
 * op's argument meaning, tensor layouts, in-place behaviour and error cases.
 * vllm-kvcompress_b200/csrc/torch_bindings.cpp re-exports them under the
 * reference's own dispatcher schemas (torch.ops._C / _C_cache_ops / _C_kvc_ops),
 * so the fork's Python (vllm/_custom_ops.py) calls them unchanged; INTEGRATION.md
 * shows the binding a maintainer adds.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless a parameter says "host".
 *  - All index tensors are int32 and contiguous (the reference kernels
 *    reinterpret_cast<int*> them, csrc/kvcompress_eviction_kernels.cu:527-542);
 *    slot_mapping is int64 (csrc/kvcompress_cache_kernels.cu:36).
 *  - Nothing is allocated and nothing synchronises: launches are asynchronous on
 *    `stream` (reference: at::cuda::getCurrentCUDAStream(), :551-552).  Ops that
 *    need scratch take it from the caller (`workspace`, size from the matching
 *    *_workspace_bytes function).
 *  - Return value: KVC_OK (0) or a negative kvc_status; kvc_last_error() gives the
 *    message of the last failure on the calling thread.  Unsupported
 *    block/head/vector sizes fail with KVC_ERR_UNSUPPORTED exactly where the
 *    reference raises TORCH_CHECK(false, "Unsupported ...").
 *  - [S,L,H] = (sequence, layer, kv head) row-major; [L,S,H] likewise.
 *    B = KV block size (tokens per block); one physical block holds B tokens of
 *    ONE kv head of ONE layer (vllm/attention/ops/paged_attn.py:256-284):
 *      k_cache [NB, D/x, B, x]   (x = 16 / sizeof(cache element))
 *      v_cache [NB, D, B]
 *      kv_metrics [NB, B] f32, kv_position [NB, B] i32.
 */
#ifndef KVC_B200_H_
#define KVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(KVC_BUILDING_LIB)
#define KVC_API __attribute__((visibility("default")))
#else
#define KVC_API
#endif

/* cudaStream_t without dragging cuda_runtime.h into C callers */
typedef struct CUstream_st* kvc_stream_t;

typedef enum kvc_status {
  KVC_OK = 0,
  KVC_ERR_UNSUPPORTED = -1, /* reference: TORCH_CHECK(false, "Unsupported ...") */
  KVC_ERR_INVALID = -2,     /* null pointer / negative size / bad alignment */
  KVC_ERR_WORKSPACE = -3,   /* workspace missing or too small */
  KVC_ERR_CUDA = -4         /* a CUDA runtime call or launch failed */
} kvc_status;

/* cache element type of k_cache / v_cache / query */
typedef enum kvc_dtype {
  KVC_F16 = 0,
  KVC_BF16 = 1,
  KVC_F32 = 2,
  KVC_U8 = 3 /* fp8_e5m2 stored as bytes: bit-copy ops only */
} kvc_dtype;

/* Message for the last non-OK status returned on this thread ("" if none). */
KVC_API const char* kvc_last_error(void);
/* Library/ABI version, bumped when a signature changes. */
KVC_API int kvc_abi_version(void);
/* Compiled-for architecture string, e.g. "sm_100a". */
KVC_API const char* kvc_arch(void);

/* ---------------------------------------------------------------------------
 * schedule_cache_evictions (+ truncate)                                    K1/K2
 * replaces _C_kvc_ops::schedule_cache_evictions
 *   csrc/torch_bindings.cpp:374-386, csrc/kvcompress_eviction.h:5-28,
 *   kernels csrc/kvcompress_eviction_kernels.cu:27-188, launcher :492-555.
 * Walks, per sequence, its metric-sorted slot list; appends (head id, logical
 * index) of every evictable slot to that head's segment until
 * evicted_blocks_per_seq[s] whole blocks are freed.  Results are bit-identical
 * to the reference's one-thread-per-sequence walk; here the walk is a chunked
 * multi-bin rank (histogram -> scan -> stable scatter).
 *   evicted_kv_indices       [max_evicted_kv] out
 *   evicted_logical_indices  [max_evicted_kv] out
 *   evicted_kv_count         [S,L,H] out
 *   remaining_kv_count       [S,L,H] in/out scratch; caller passes a copy of
 *                            hanging_token_count (vllm/_custom_ops.py:987)
 *   evicted_kv_offsets       [S,L,H]
 *   sorted_indices           [total_blocks*B] slot ids sorted by (seq, metric)
 *   seq_block_offsets        [S]
 *   layer_by_block, head_by_block, virtual_block_num_by_block [total_blocks]
 *   evicted_blocks_per_seq   [S]
 *   context_lens             [S,L,H]   (NOTE: [S,L,H] here, _custom_ops.py:995)
 *   hanging_token_count      [S,L,H]
 *   kv_position              [total_blocks, B]
 *   last_position, protected_window_size [S]
 *   control_layers           [num_control_layers] or NULL
 * block_size in {1,2,4,16} else KVC_ERR_UNSUPPORTED (:474-490).
 * ------------------------------------------------------------------------- */
KVC_API size_t kvc_schedule_cache_evictions_workspace_bytes(
    int num_seqs, int num_layers, int num_kv_heads, int total_blocks,
    int block_size);

KVC_API int kvc_schedule_cache_evictions(
    int* evicted_kv_indices, int* evicted_logical_indices,
    int* evicted_kv_count, int* remaining_kv_count,
    const int* evicted_kv_offsets, const int* sorted_indices,
    const int* seq_block_offsets, const int* layer_by_block,
    const int* head_by_block, const int* virtual_block_num_by_block,
    const int* evicted_blocks_per_seq, const int* context_lens,
    const int* hanging_token_count, const int* kv_position,
    const int* last_position, const int* protected_window_size,
    int num_seqs, int num_layers, int num_kv_heads, int total_blocks,
    int block_size, int evict_evenly_per_layer, const int* control_layers,
    int num_control_layers, int max_evicted_kv, int null_eviction_index,
    int truncate, void* workspace, size_t workspace_bytes,
    kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * truncate_cache_evictions                                                   K2
 * replaces _C_kvc_ops::truncate_cache_evictions
 *   csrc/torch_bindings.cpp:389-394, csrc/kvcompress_eviction.h:30-38,
 *   kernel csrc/kvcompress_eviction_kernels.cu:158-188, launcher :585-610.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_truncate_cache_evictions(
    int* evicted_kv_indices, int* evicted_logical_indices,
    int* evicted_kv_count, const int* evicted_kv_offsets,
    const int* hanging_token_count, int num_seqs, int num_layers,
    int num_kv_heads, int block_size, int max_evicted_kv,
    int null_eviction_index, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * count_block_evictions                                                      K3
 * replaces _C_kvc_ops::count_block_evictions
 *   csrc/torch_bindings.cpp:397-401, csrc/kvcompress_eviction.h:40-46,
 *   kernel csrc/kvcompress_eviction_kernels.cu:190-221, launcher :640-664.
 * total_kvs = numel(evicted_logical_indices).
 * ------------------------------------------------------------------------- */
KVC_API int kvc_count_block_evictions(
    int* evicted_block_count, int* evicted_logical_indices,
    const int* evicted_kv_offsets, const int* hanging_token_count,
    int num_seqs, int num_layers, int num_kv_heads, int total_kvs,
    int block_size, int null_value, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * schedule_t1_cache_moves                                                    K4
 * replaces _C_kvc_ops::schedule_t1_cache_moves
 *   csrc/torch_bindings.cpp:404-409, csrc/kvcompress_eviction.h:48-56,
 *   kernel csrc/kvcompress_eviction_kernels.cu:223-289, launcher :698-726.
 *   cache_moves_idx [max_evicted_kv, 2] (dst, src) physical slot ids, out
 *   cache_moves_count [S,L,H] out
 *   evicted_logical_indices [max_evicted_kv] (ascending per head segment)
 *   evicted_kv_count, evicted_kv_offsets [S,L,H]
 *   block_tables [L,S,H,MB], context_lens [L,S,H]
 * ------------------------------------------------------------------------- */
KVC_API int kvc_schedule_t1_cache_moves(
    int* cache_moves_idx, int* cache_moves_count,
    const int* evicted_logical_indices, const int* evicted_kv_count,
    const int* evicted_kv_offsets, const int* block_tables,
    const int* context_lens, int num_seqs, int num_layers, int num_kv_heads,
    int max_num_blocks_per_seq, int block_size, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * execute_cache_moves                                                        K5
 * replaces _C_kvc_ops::execute_cache_moves
 *   csrc/torch_bindings.cpp:412-417, csrc/kvcompress_eviction.h:69-78,
 *   kernel csrc/kvcompress_eviction_kernels.cu:359-435, launcher :858-929.
 * Bit-copies K, V, metric and position of every scheduled (dst, src) pair.
 * head_size in {1,2,4,128}, vec_size (x) in {1,2,8}, block_size in {1,2,4,16}
 * else KVC_ERR_UNSUPPORTED (:793-856).  blocks_per_head / threads_per_head are
 * launch hints in the reference (cache_engine.py:149-150); results never depend
 * on them and this implementation ignores them.
 * Algorithmic bytes per move (D=128, 2-byte cache): 2*(2*128*2+4+4)+8 = 1048.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_execute_cache_moves(
    void* k_cache, void* v_cache, float* kv_metrics, int* kv_position,
    const int* cache_moves_idx, const int* cache_moves_count,
    const int* evicted_kv_offsets, int num_seqs, int num_layers,
    int num_kv_heads, int64_t num_blocks, int head_size, int vec_size,
    int block_size, int dtype /* kvc_dtype */, int blocks_per_head,
    int threads_per_head, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * kvcompress_reshape_and_cache                                               K6
 * replaces _C_cache_ops::kvcompress_reshape_and_cache
 *   csrc/torch_bindings.cpp:354-362, csrc/kvcompress_cache.h:5-15,
 *   kernel csrc/kvcompress_cache_kernels.cu:27-89, launcher :111-139.
 *   key, value [T, H, D] with token strides key_stride / value_stride (elements)
 *   slot_mapping [T, H] int64 (<0 = padding)
 *   kv_metric_head_bias [H] f32: metric of every written slot is set to it.
 * kv_cache_dtype "auto" only (dtype = element type of key AND cache).
 * Batches of >= 512 tokens in the production layout (fp16 / bf16, D = 128, B = 16, x = 8) are
 * written block-wise: 16 batch tokens of a head that map to the 16 slots of one block in order
 * become one 4 KB block write (V transposed in shared memory); everything else, and every other
 * shape, goes element-wise like the reference.  Same bytes in the cache either way.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_reshape_and_cache(
    const void* key, const void* value, void* key_cache, void* value_cache,
    float* kv_metrics, const int64_t* slot_mapping,
    const float* kv_metric_head_bias, int num_tokens, int num_heads,
    int head_size, int block_size, int x, int64_t key_stride,
    int64_t value_stride, int dtype /* kvc_dtype */, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * kvcompress_paged_attention_v1 / v2                                     K7 / K8
 * replaces _C::kvcompress_paged_attention_v1 / _v2
 *   csrc/torch_bindings.cpp:53-80, csrc/ops.h:18-37, 67-90,
 *   kernels csrc/attention/kvcompress_attention_kernels.cu:95-651,
 *   launchers :683-771, :878-978.
 * Single-query (decode) attention over per-(seq, kv head) block tables and
 * context lengths; additionally writes each KV's softmax probability per query
 * head to kv_metric_out[slot*Q + q_off] iff record_kv_metrics and
 * kv_position[slot] <= last_position[s] - kv_metric_buffer_len[s] (:124,:305-313).
 *   out [S, Hq, D]; query [S, Hq, D] row stride q_stride (elements)
 *   kv_metric_out [NB, B, Q] f32
 *   block_tables [S, H, MB], context_lens [S, H]
 *   alibi_slopes [Hq] f32 or NULL
 * v2 extra buffers (caller-allocated, same shapes as the reference):
 *   exp_sums, max_logits [S, Hq, P] f32, tmp_out [S, Hq, P, D] (query dtype),
 *   P = ceil(max_context_len / 512); tmp_kv_metric_out is accepted for ABI
 *   compatibility and never touched.
 * head_size must be 128 (:745-770), block_size 16, dtype F16/BF16,
 * kv_cache_dtype "auto" (the only combination the KVC engine can run end to
 * end: K5 supports B in {1,2,4,16}, the reference attention B in {8,16,32}).
 * ------------------------------------------------------------------------- */
KVC_API int kvc_paged_attention_v1(
    void* out, float* kv_metric_out, const void* query, const void* key_cache,
    const void* value_cache, int num_seqs, int num_heads, int num_kv_heads,
    int head_size, float scale, const int* block_tables,
    const int* context_lens, const int* kv_position, const int* last_position,
    const int* kv_metric_buffer_len, int block_size, int max_context_len,
    int max_num_blocks_per_seq, int64_t q_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype /* kvc_dtype */,
    int record_kv_metrics, kvc_stream_t stream);

KVC_API int kvc_paged_attention_v2(
    void* out, float* kv_metric_out, float* exp_sums, float* max_logits,
    void* tmp_out, float* tmp_kv_metric_out, const void* query,
    const void* key_cache, const void* value_cache, int num_seqs,
    int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size,
    int max_context_len, int max_num_blocks_per_seq, int64_t q_stride,
    int64_t kv_block_stride, const float* alibi_slopes,
    int dtype /* kvc_dtype */, int record_kv_metrics, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Additive (not in the reference's op list): decode attention that folds the
 * per-step metric aggregation into the kernel.
 * replaces the chain  CompressionMetrics.clear_temp_metrics()  (zero [NB,B,Q]),
 * paged_attention_kvc_v1(record=True), CompressionMetrics.aggregate_decode()
 * (metrics += (temp**2).sum(-1) over the WHOLE cache)
 *   vllm/kvcompress/metrics.py:337-342, 429-439; llm_engine.py:1621-1634
 * with one launch that does  kv_metrics[slot] += sum_q p[slot,q]^2  for the
 * slots it visits (same recording rule as above).  No [NB,B,Q] buffers.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_paged_attention_fused_metrics(
    void* out, float* kv_metrics /* [NB,B] f32, accumulated in place */,
    const void* query, const void* key_cache, const void* value_cache,
    int num_seqs, int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size,
    int max_context_len, int max_num_blocks_per_seq, int64_t q_stride,
    int64_t kv_block_stride, const float* alibi_slopes,
    int dtype /* kvc_dtype */, void* workspace, size_t workspace_bytes,
    kvc_stream_t stream);
/* scratch for the parked raw logits: num_seqs * num_heads * ceil16(max_context_len) floats */
KVC_API size_t kvc_paged_attention_fused_metrics_workspace_bytes(
    int num_seqs, int num_heads, int max_context_len);

/* ---------------------------------------------------------------------------
 * Additive: one decode layer of the hot path in ONE launch = kvcompress_reshape_and_cache
 * (csrc/kvcompress_cache_kernels.cu:27-89: write the step's new K/V of every (seq, kv head)
 * to its slot, metric := head bias) + paged attention + `metrics += sum_q p^2`
 * (call sites vllm/attention/backends/flash_attn.py:871 and :1067-1091,
 * vllm/kvcompress/metrics.py:429-439).  context_lens must already count the new token
 * (batch_append_slots ran).  key/value [S, H, D] with token strides key_stride/value_stride;
 * slot_mapping [S, H] int64 for THIS layer (< 0 = no new token).  Workspace as for
 * kvc_paged_attention_fused_metrics.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_paged_attention_fused_step(
    void* out, float* kv_metrics, const void* query, const void* key,
    const void* value, void* key_cache, void* value_cache,
    const int64_t* slot_mapping, const float* kv_metric_head_bias, int num_seqs,
    int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size,
    int max_context_len, int max_num_blocks_per_seq, int64_t q_stride,
    int64_t key_stride, int64_t value_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype /* kvc_dtype */, void* workspace,
    size_t workspace_bytes, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Additive: deferred append.  In the steady state of continual compression every head sits on a
 * block boundary after each compression (block.py:184-210), so the step's new token opens a fresh
 * block -- 128 + 16 partial-sector stores (a 4.6 KB fill + 4.6 KB write-back in DRAM) -- and the next
 * compression moves exactly that token into an evicted slot and frees the block again
 * (kvcompress_eviction_kernels.cu:252-288 emits the newest surviving token first), reading the 4 KB
 * block back for 512 useful bytes.  With these two entry points such a token is parked instead in a
 * compact side buffer and stored once, where it finally belongs:
 *   kvc_paged_attention_fused_step_deferred: as kvc_paged_attention_fused_step; a new token whose
 *     slot is the first of its block goes to pend_kv [S, H, 2*D] (K | V, cache dtype) with
 *     pend_slot [S, H] = slot (int64); every other head gets pend_slot = -1 and the ordinary store.
 *     Metric (head bias) and position of the slot are written as usual.
 *   kvc_execute_cache_moves_deferred: as kvc_execute_cache_moves with pend_kv [L, S, H, 2*D] /
 *     pend_slot [L, S, H]: a move whose source is the head's pending slot copies from pend_kv; a
 *     pending token that no move consumes is stored to its own slot.  On completion the cache is
 *     bit-identical to what reshape_and_cache + execute_cache_moves leave.  Must run before anything
 *     else reads K/V of the pending slots (the fused step kernel itself patches the staged block from
 *     registers).  Production layout only (2-byte cache, D=128, B=16, x=8).
 *   Deferred MOVE (pend_dst [L, S, H] int64, optional in both calls, initialised to -1): when the head's only
 *     move takes the parked token, kvc_execute_cache_moves_deferred moves metric and position only and
 *     records the destination slot in pend_dst; the next kvc_paged_attention_fused_step_deferred stores the
 *     parked K/V there in its consumer warps' prologue, holds the TMA fetch of that block back until the store
 *     is visible (so its own attention already reads the token; the sectors L2 fetches for the partial stores
 *     are the ones the block's bulk copy needs next) and resets pend_dst.  Between the two calls the cache
 *     lacks that token's K/V at the destination.
 *   kvc_flush_deferred: stores every parked token where it belongs (pend_dst if set, else pend_slot) and
 *     clears both markers: afterwards the cache is canonical.  pend arrays flattened [num_heads].
 * ------------------------------------------------------------------------- */
KVC_API int kvc_paged_attention_fused_step_deferred(
    void* out, float* kv_metrics, const void* query, const void* key,
    const void* value, void* key_cache, void* value_cache,
    const int64_t* slot_mapping, const float* kv_metric_head_bias,
    void* pend_kv, int64_t* pend_slot, int64_t* pend_dst /* nullable */, int num_seqs,
    int num_heads, int num_kv_heads, int head_size, float scale,
    const int* block_tables, const int* context_lens, const int* kv_position,
    const int* last_position, const int* kv_metric_buffer_len, int block_size,
    int max_context_len, int max_num_blocks_per_seq, int64_t q_stride,
    int64_t key_stride, int64_t value_stride, int64_t kv_block_stride,
    const float* alibi_slopes, int dtype /* kvc_dtype */, void* workspace,
    size_t workspace_bytes, kvc_stream_t stream);

KVC_API int kvc_execute_cache_moves_deferred(
    void* k_cache, void* v_cache, float* kv_metrics, int* kv_position,
    const int* cache_moves_idx, const int* cache_moves_count,
    const int* evicted_kv_offsets, const void* pend_kv, const int64_t* pend_slot,
    int64_t* pend_dst /* nullable */, int num_seqs, int num_layers, int num_kv_heads,
    int64_t num_blocks, int head_size, int vec_size, int block_size,
    int dtype /* kvc_dtype */, kvc_stream_t stream);

KVC_API int kvc_flush_deferred(
    void* k_cache, void* v_cache, const void* pend_kv, int64_t* pend_slot,
    int64_t* pend_dst /* nullable */, int64_t num_heads, int head_size, int vec_size,
    int block_size, int dtype /* kvc_dtype */, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Launch mode of the streaming decode-attention kernel (process-wide, read at every launch; initial
 * value from the environment variable KVC_ATTN_PDL, default 1).  PDL = programmatic dependent launch
 * (cudaLaunchAttributeProgrammaticStreamSerialization): the kernel's CTAs may become resident while the
 * previous kernel of the stream drains.
 *   0  plain stream-ordered launches.
 *   1  PDL, always safe (default; use it for the drop-in sequence kvcompress_reshape_and_cache ->
 *      paged attention and right after any kernel that touches tables or the cache): every warp executes
 *      griddepcontrol.wait before its first global access and only then lets the next kernel start;
 *      what overlaps the previous kernel's tail is CTA launch, barrier setup and shared-memory carving.
 *   2  PDL, caller declares that block tables, context lengths, slot mapping, cache blocks and parked
 *      tokens read by this launch were NOT written by the immediately preceding kernel of the stream
 *      (it was the model's QKV projection / rotary kernel, or the previous layer's attention, which only
 *      touches its own layer's blocks): the TMA producers stream cache blocks at once; the consumers
 *      execute griddepcontrol.wait before they read query / new key / new value.  The launch in front
 *      of a run of mode-2 launches must be mode 0/1 (engine.py: layer 0 = 1, layers >= 1 = 2).
 *   3  as 2, and query / key / value are not produced by the preceding kernel either (synthetic
 *      benchmarks with every layer's inputs resident): the wait moves to the kernel's end.  Falls back
 *      to 2 whenever raw logits are parked in the caller's workspace (contexts too long for shared
 *      memory), because consecutive launches may share that workspace.  Not usable by a model.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_set_attention_pdl(int mode);
KVC_API int kvc_get_attention_pdl(void);

/* Partition kernel of kvc_paged_attention_v2 (long per-head contexts; process-wide, initial value from the
 * environment variable KVC_ATTN_V2_MODE): 1 = mma.sync fragments loaded straight from the paged blocks;
 * 2 = tcgen05 / TMEM: 128-token UMMA tiles (S^T = K Q^T, O^T = V^T P^T, fp32 accumulators in tensor memory),
 * operands staged in the canonical K-major core-matrix layout; 0 (default) = the faster of the two on the B200.
 * Same buffers and results (within the fp16 tolerance) in every mode. */
KVC_API int kvc_set_attention_v2_mode(int mode);
KVC_API int kvc_get_attention_v2_mode(void);

/* ---------------------------------------------------------------------------
 * Additive: fused eviction selection straight from the paged metrics.
 * replaces CompressionMetrics.schedule_evictions (vllm/kvcompress/metrics.py:441-847:
 * mask gathers + 6 torch.sort + per-sequence host loop + count_block_evictions + final
 * sort) and, equivalently, sort_seq_metrics (:849-966) -> schedule_cache_evictions +
 * truncate -> host re-sort (vllm/kvcompress/scheduler.py:472-477).  Output is what
 * schedule_t1_cache_moves consumes:
 *   evicted_logical_indices [max_evicted_kv]: per head segment (evicted_kv_offsets)
 *       the evicted logical indices ascending, null_eviction_index after them
 *   evicted_kv_count [S,L,H] = (k-1)*B + hanging for k freed blocks, else 0
 *   evicted_block_count [S,L,H] = k
 * Inputs: kv_metrics/kv_position [NB,B]; block_tables [L,S,H,MB], context_lens [L,S,H]
 * (batch views, as for schedule_t1_cache_moves); hanging_token_count,
 * evicted_kv_offsets [S,L,H]; evicted_blocks_per_seq, last_position,
 * protected_window_size [S].  Candidate rule and completion logic = K1
 * (csrc/kvcompress_eviction_kernels.cu:136-153); ties by (metric, logical index).
 * Head-bias / use_average / attention-sink terms: kvc_select_evictions_ex below.
 * ------------------------------------------------------------------------- */
/* Launch shape of kvc_select_evictions[_ex] (process-wide): 0 = by shape (default): calls with >= 96
 * sequences, block size 16 and <= 512 slots per head run as ONE launch, a CTA per sequence, whose sort is
 * pruned to the keys at or below a per-sequence bound on the E-th completion key (see kvc_select.cu);
 * 1 = always the three launches (rank, threshold, emit); 2 = the single launch whenever the shape
 * allows.  Long contexts (more than 512 slots per head) take the sort-free radix path in modes 0 / 2 / 3: as
 * ONE cooperative launch (keys resident in shared memory, grid barriers between the digit passes) when every
 * head's CTA fits on the device at once, else -- and always in mode 3 -- as seven launches.  Results are
 * identical in every mode. */
KVC_API int kvc_set_select_mode(int mode);
KVC_API int kvc_get_select_mode(void);

/* Kernel choice of kvc_execute_cache_moves for the production layout (process-wide; initial value from the
 * environment variable KVC_MOVES_MODE): 0 (default) = the register-gather run kernel (post-prefill shapes) and
 * the one-move / run kernels (decode-time batches); 1 = post-prefill compaction (fewer than 4096 heads) runs the
 * TMA-staged kernel instead: whole source V blocks and the old destination block bulk-copied into shared
 * memory (cp.async.bulk), lane = token column copies, finished K / V blocks written back with one bulk store
 * each.  Results are identical in both modes; on the B200 the staged kernel is bound by the bytes it can keep in
 * flight from ~200 KB of shared memory and is the slower of the two. */
KVC_API int kvc_set_moves_mode(int mode);
KVC_API int kvc_get_moves_mode(void);

/* Workspace of kvc_select_evictions[_ex]: _workspace_bytes_ex is exact for a call's shape (contexts of more
 * than 512 slots per head take the sort-free radix path, which stores one 4-byte key per slot and 4 x 256 sums
 * per sequence); _workspace_bytes is a shape-agnostic upper bound (up to 1024 sequences per call). */
KVC_API size_t kvc_select_evictions_workspace_bytes(int64_t max_evicted_kv, int block_size);
KVC_API size_t kvc_select_evictions_workspace_bytes_ex(int64_t max_evicted_kv, int block_size, int num_seqs,
                                                       int max_num_blocks_per_seq);

KVC_API int kvc_select_evictions(
    int* evicted_logical_indices, int* evicted_kv_count, int* evicted_block_count,
    const float* kv_metrics, const int* kv_position, const int* block_tables,
    const int* context_lens, const int* hanging_token_count,
    const int* evicted_kv_offsets, const int* evicted_blocks_per_seq,
    const int* last_position, const int* protected_window_size, int num_seqs,
    int num_layers, int num_kv_heads, int max_num_blocks_per_seq, int block_size,
    int64_t max_evicted_kv, int null_eviction_index, void* workspace,
    size_t workspace_bytes, kvc_stream_t stream);

/* kvc_select_evictions with the optional terms of the live selection
 * (vllm/kvcompress/metrics.py:495-506, 539-544; KVHeadBias :45-81):
 *   use_average: metric / (last_position[s] - kv_position)            (:495-501)
 *   head_bias [L, H, num_bins] f32 + position_bins [num_bins] i32 (ascending bin starts):
 *       metric += head_bias[l, h, bin(kv_position)] * bias_weight      (:503-506)
 *   num_sinks: slots with kv_position < num_sinks are never evicted   (:542)
 * head_bias NULL / use_average 0 / num_sinks 0 = kvc_select_evictions. */
KVC_API int kvc_select_evictions_ex(
    int* evicted_logical_indices, int* evicted_kv_count, int* evicted_block_count,
    const float* kv_metrics, const int* kv_position, const int* block_tables,
    const int* context_lens, const int* hanging_token_count,
    const int* evicted_kv_offsets, const int* evicted_blocks_per_seq,
    const int* last_position, const int* protected_window_size, int num_seqs,
    int num_layers, int num_kv_heads, int max_num_blocks_per_seq, int block_size,
    int64_t max_evicted_kv, int null_eviction_index, const float* head_bias,
    const int* position_bins, int num_bins, float bias_weight, int use_average,
    int num_sinks, void* workspace, size_t workspace_bytes, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Additive: block-state device ops around a compression cycle (the reference does these
 * with ~30 small torch launches, vllm/kvcompress/block.py:184-210, 275-335;
 * block_manager.py:269-294; scheduler.py:131-141, 263-280).
 * kvc_prepare_compression: from context_lens [L,S,H] compute hanging_token_count [S,L,H]
 *   (ctx % B or B), evicted_kv_offsets [S,L,H] (exclusive cumsum of ceil(ctx/B)*B in
 *   (s,l,h) order) and, if evicted_blocks_per_seq != NULL, the per-sequence budget
 *   max(0, allocated blocks - max_cache_blocks_per_seq).  seq_slots_workspace: [S] ints.
 * kvc_free_and_append_slots: per head, context_lens -= max(evicted_block_count*B -
 *   (B - hanging), 0) (skipped when evicted_block_count == NULL), then if `append`:
 *   slot = block_tables[.., ctx / B] * B + ctx % B -> slot_mapping [L,S,H] (int64),
 *   kv_position[slot] = seq_len[s], ctx += 1.  A head whose block table is full (ctx / B ==
 *   max_num_blocks_per_seq; the reference raises an index error, block.py:303-321) appends nothing:
 *   its slot_mapping entry is -1 and *status (optional, int32 on the device) is set to 2.
 * ------------------------------------------------------------------------- */
KVC_API int kvc_prepare_compression(
    const int* context_lens, int* hanging_token_count, int* evicted_kv_offsets,
    int* evicted_blocks_per_seq, int* seq_slots_workspace, int num_seqs,
    int num_layers, int num_kv_heads, int block_size,
    int max_cache_blocks_per_seq, kvc_stream_t stream);

KVC_API int kvc_free_and_append_slots(
    int* context_lens, const int* block_tables, const int* evicted_block_count,
    const int* hanging_token_count, const int* seq_len, int64_t* slot_mapping,
    int* kv_position, int num_seqs, int num_layers, int num_kv_heads,
    int max_num_blocks_per_seq, int block_size, int append, int* status, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Additive: prefill KV-metric initialisation on the tensor cores (tcgen05 / TMEM).
 * replaces _naive_kvc_attention + _naive_kvc_masked_attention
 *   vllm/attention/backends/flash_attn.py:1122-1211 (call site :993-1007: a second dense
 *   pass over repeat_kv(key) in torch) and, when kv_metrics/slot_mapping are given,
 *   CompressionMetrics.aggregate_prefill (vllm/kvcompress/metrics.py:395-427) as well.
 * For each prompt i (tokens [seq_start_loc[i], seq_start_loc[i+1])): the last
 * n_obs = min(prompt_len, n_observed) queries attend causally to the prompt's keys;
 *   metric[k, hq] = sum_q [k <= pos(q) - kv_metric_buffer_len[i]] * softmax_k(scale * q.k)^2
 * (^2 iff use_l2), * (k+1)/n_obs iff use_average, max-pool(7, pad 3) along k iff use_maxpool.
 *   kv_metric_out [T, Hq] f32 or NULL
 *   kv_metrics [NB, B] f32 + slot_mapping [T, H] int64, both or neither:
 *       kv_metrics[slot_mapping[t,h]] += sum over the kv head's query heads (aggregate_prefill)
 *   query [T, Hq, D], key [T, H, D] (NOT repeated per query head), token strides in elements
 *   seq_start_loc [num_seqs+1], kv_metric_buffer_len [num_seqs]; max_prompt_len (host int)
 * The QK product is rounded to the input type before the fp32 scale exactly as the reference's
 * fp16/bf16 einsum is (flash_attn.py:1178).  head_size 128, F16/BF16, n_observed <= 4096 (one
 * observation chunk: the reference default max_observed_block_size).
 * query / key are read through TMA tensor maps built per call from the pointers and token
 * strides (so they may be views of a fused qkv projection output): both must be 16-byte
 * aligned with token strides that are multiples of 8 elements, heads contiguous within a token.
 * Three launches (row statistics, squared-probability column sums, max-pool + aggregation),
 * the last two as programmatic dependents of their predecessor.
 * ------------------------------------------------------------------------- */
KVC_API size_t kvc_prefill_kv_metrics_workspace_bytes(
    int num_tokens, int num_seqs, int num_heads, int num_kv_heads,
    int n_observed, int max_prompt_len);

KVC_API int kvc_prefill_kv_metrics(
    float* kv_metric_out, float* kv_metrics, const int64_t* slot_mapping,
    const void* query, const void* key, const int* seq_start_loc,
    const int* kv_metric_buffer_len, int num_tokens, int num_seqs,
    int num_heads, int num_kv_heads, int head_size, int64_t q_stride,
    int64_t k_stride, int max_prompt_len, float scale, int n_observed,
    int use_l2, int use_average, int use_maxpool, int dtype /* kvc_dtype */,
    void* workspace, size_t workspace_bytes, kvc_stream_t stream);

/* aggregate_prefill alone (vllm/kvcompress/metrics.py:395-427):
 * kv_metrics[slot_mapping[t,h]] += sum_q prefill_metrics[t, h*Q + q]. */
KVC_API int kvc_aggregate_prefill(
    float* kv_metrics, const float* prefill_metrics,
    const int64_t* slot_mapping, int num_tokens, int num_heads,
    int num_kv_heads, kvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Additive: device-side block allocator and block-state updates.
 * replaces ParallelBlockAllocator (vllm/kvcompress/block_manager.py:75-140: every allocate()
 * compacts the whole free mask, `block_numbers[free_mask][:n]`) and the BlockSpaceManagerKVC
 * paths on top of it, which materialise [L,S,H,MB] boolean masks and advanced-index copies of
 * the block tables (block.py:337-412): _add_sequence (:196-222), _append_to_sequence_batch
 * (:269-294), free_compressed_blocks (:466-530 + block.py:184-210), _remove_sequence_batch
 * (:243-262), plus CompressionMetrics.insert_metadata / remove_metadata (metrics.py:344-370).
 * State = the reference's own tensors, updated IN PLACE (no batch-view copies):
 *   free_mask [num_blocks] bool bytes (1 = free), free_count [1] i32, status [1] i32
 *   block_tables [L, max_seqs, H, MB] i32, context_lens [L, max_seqs, H] i32
 *   kv_position (= token_positions) [num_blocks, B] i32
 *   optional metadata [num_blocks] i32 each (NULL to skip): seq_index_by_block,
 *   layer_index_by_block, head_index_by_block, logical_block_num_by_block
 * Allocation order is the reference's: the lowest-numbered free blocks, ascending, handed out
 * in row-major order of (layer, batch position, head[, block]); block tables come out identical.
 * If fewer blocks are free than requested NOTHING is allocated and *status is set to 1
 * (reference: ValueError("Out of memory!") on the host, :108-110); the caller polls status.
 * seq_slots [n] = batch slot index of each sequence of the call (NULL = 0..n-1 for append/free).
 * ------------------------------------------------------------------------- */
KVC_API size_t kvc_alloc_workspace_bytes(int64_t num_blocks, int num_batch,
                                         int num_layers, int num_kv_heads);

/* free_all (:127-129): every block free, free_count = num_blocks, status = 0 */
KVC_API int kvc_alloc_init(uint8_t* free_mask, int* free_count, int* status,
                           int64_t num_blocks, kvc_stream_t stream);

/* _append_to_sequence_batch + decode slot mapping (block.py:303-321): every head of the batch
 * grows by one token at position seq_len[i]; heads on a block boundary get a new block.
 * slot_mapping [L, num_batch, H] int64 out. */
KVC_API int kvc_alloc_append_slots(
    uint8_t* free_mask, int* free_count, int* status, int64_t num_blocks,
    int* block_tables, int* context_lens, const int* seq_slots,
    const int* seq_len, int64_t* slot_mapping, int* kv_position,
    int* seq_index_by_block, int* layer_index_by_block,
    int* head_index_by_block, int* logical_block_num_by_block, int num_batch,
    int max_seqs, int num_layers, int num_kv_heads, int max_num_blocks_per_seq,
    int block_size, void* workspace, size_t workspace_bytes, kvc_stream_t stream);

/* _add_sequence for num_new prompts: seq_lens [num_new] tokens each, ceil(len/B) blocks per
 * head in (sequence, layer, head, block) order; context_lens = len; positions = logical index.
 * max_seq_len (host) = max of seq_lens. */
KVC_API int kvc_alloc_add_sequences(
    uint8_t* free_mask, int* free_count, int* status, int64_t num_blocks,
    int* block_tables, int* context_lens, const int* seq_slots,
    const int* seq_lens, int* kv_position, int* seq_index_by_block,
    int* layer_index_by_block, int* head_index_by_block,
    int* logical_block_num_by_block, int num_new, int max_seq_len, int max_seqs,
    int num_layers, int num_kv_heads, int max_num_blocks_per_seq, int block_size,
    void* workspace, size_t workspace_bytes, kvc_stream_t stream);

/* free_compressed_blocks: evicted_block_count [num_batch, L, H] trailing blocks of each head are
 * freed, context_lens -= max(k*B - (B - hanging), 0), seq_index_by_block[freed] = -1. */
KVC_API int kvc_alloc_free_compressed(
    uint8_t* free_mask, int* free_count, const int* block_tables,
    int* context_lens, const int* evicted_block_count, const int* seq_slots,
    int* seq_index_by_block, int num_batch, int max_seqs, int num_layers,
    int num_kv_heads, int max_num_blocks_per_seq, int block_size,
    kvc_stream_t stream);

/* _remove_sequence_batch: free every allocated block of the sequences, context_lens = 0. */
KVC_API int kvc_alloc_remove_sequences(
    uint8_t* free_mask, int* free_count, const int* block_tables,
    int* context_lens, const int* seq_slots, int* seq_index_by_block,
    int num_removed, int max_seqs, int num_layers, int num_kv_heads,
    int max_num_blocks_per_seq, int block_size, kvc_stream_t stream);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* KVC_B200_H_ */
