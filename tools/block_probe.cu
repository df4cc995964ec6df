// Probe: what DRAM throughput does the compaction's ACCESS PATTERN allow?  Random whole 4 KB cache blocks
// bulk-copied global -> shared (cp.async.bulk) with a fixed number of copies in flight per SM and no processing
// at all; optionally every n-th staged block is bulk-stored to another random block (the compaction's
// read : write mix).  Prints GB/s per configuration; the streaming-copy peak of MEASURED_PEAKS.json is the
// reference point.  Used for DESIGN.md section 5 (the pattern ceiling of execute_cache_moves).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/block_probe tools/block_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread per ring slot: issue -> wait -> (store) -> next; `slots` copies of 4 KB in flight per CTA, 1 CTA per SM
__global__ void __launch_bounds__(64, 1) block_stream(unsigned char* __restrict__ buf, const int* __restrict__ ids,
                                                      const int* __restrict__ wids, int per_slot, int slots,
                                                      int store_every, int chunk) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int t = threadIdx.x;
  if (t >= slots) return;
  const uint32_t bar = smem_u32(smem + (size_t)slots * chunk) + 8 * t;
  const uint32_t dst = smem_u32(smem) + t * chunk;
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const int* my = ids + ((size_t)blockIdx.x * slots + t) * per_slot;
  const int* myw = wids + ((size_t)blockIdx.x * slots + t) * per_slot;
  for (int u = 0; u < per_slot; ++u) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(buf + (size_t)my[u] * chunk), "r"(chunk), "r"(bar) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(bar), "r"(u & 1) : "memory");
    if (store_every > 0 && (u % store_every) == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(buf + (size_t)myw[u] * chunk), "r"(dst), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const size_t total = (size_t)8 << 30;  // 8 GiB: far beyond the 126 MB L2
  unsigned char* buf;
  cudaMalloc(&buf, total);
  cudaMemset(buf, 1, total);
  std::mt19937 rng(0);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int chunk : {4096, 2048, 512}) {
    const size_t nblocks = total / chunk;
    for (int store_every : {0, 4}) {
      for (int slots : {8, 16, 32, 48}) {
        if ((size_t)slots * chunk > 200 * 1024) continue;
        const int per_slot = std::max(8, (int)(((size_t)3 << 30) / chunk / ((size_t)sms * slots)));  // ~3 GiB read
        const size_t need = (size_t)sms * slots * per_slot;
        std::vector<int> ids(need), wids(need);
        for (auto& x : ids) x = (int)(rng() % nblocks);
        for (auto& x : wids) x = (int)(rng() % nblocks);
        int *d_ids, *d_w;
        cudaMalloc(&d_ids, need * 4);
        cudaMalloc(&d_w, need * 4);
        cudaMemcpy(d_ids, ids.data(), need * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_w, wids.data(), need * 4, cudaMemcpyHostToDevice);
        const size_t smem = (size_t)slots * chunk + slots * 8;
        cudaFuncSetAttribute(block_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e9;
        for (int it = 0; it < 4; ++it) {
          cudaEventRecord(e0);
          block_stream<<<sms, 64, smem>>>(buf, d_ids, d_w, per_slot, slots, store_every, chunk);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          best = std::min(best, ms);
        }
        const double rd = (double)need * chunk, wr = store_every ? rd / store_every : 0.0;
        cudaError_t e = cudaGetLastError();
        printf("{\"chunk\": %d, \"in_flight_per_sm_kb\": %d, \"store_every\": %d, \"read_gbs\": %.1f, \"total_gbs\": %.1f, \"ms\": %.3f%s}\n",
               chunk, slots * chunk / 1024, store_every, rd / best / 1e6, (rd + wr) / best / 1e6, best,
               e == cudaSuccess ? "" : ", \"error\": 1");
        cudaFree(d_ids);
        cudaFree(d_w);
      }
    }
  }
  return 0;
}
