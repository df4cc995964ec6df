// Probe: throughput of cp.async.bulk (non-tensor TMA) global->shared for random chunks, one ring per warp.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tma_probe tools/tma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NSLOT>
__global__ void tma_read(const unsigned char* __restrict__ buf, const int* __restrict__ ids, int chunks_per_warp,
                         int chunk_bytes, unsigned* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int wid = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int nw = blockDim.x / 32;
  unsigned char* ring = smem + (size_t)wid * (NSLOT * chunk_bytes + 64);
  const uint32_t bar0 = smem_u32(ring + NSLOT * chunk_bytes), ring0 = smem_u32(ring);
  if (lane == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int* my = ids + ((size_t)blockIdx.x * nw + wid) * chunks_per_warp;
  auto issue = [&](int u) {
    if (lane == 0) {
      const uint32_t bar = bar0 + 8 * (u % NSLOT);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(ring0 + (u % NSLOT) * chunk_bytes), "l"(buf + (size_t)my[u] * chunk_bytes), "r"(chunk_bytes), "r"(bar) : "memory");
    }
  };
  for (int u = 0; u < NSLOT - 1 && u < chunks_per_warp; ++u) issue(u);
  unsigned acc = 0;
  for (int u = 0; u < chunks_per_warp; ++u) {
    const uint32_t bar = bar0 + 8 * (u % NSLOT), par = (u / NSLOT) & 1;
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(par) : "memory");
    acc += *reinterpret_cast<const unsigned*>(ring + (u % NSLOT) * chunk_bytes + lane * 16);
    __syncwarp();
    if (u + NSLOT - 1 < chunks_per_warp) issue(u + NSLOT - 1);
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const size_t total = (size_t)4 << 30;
  unsigned char* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  unsigned* sink; cudaMalloc(&sink, 4);
  std::mt19937 rng(0);
  for (int chunk_kb : {2, 4, 8}) {
    for (int warps_per_sm : {4, 8, 16}) {
      const int chunk_bytes = chunk_kb << 10;
      const size_t nchunks = total / chunk_bytes;
      std::vector<int> ids(nchunks);
      std::iota(ids.begin(), ids.end(), 0);
      std::shuffle(ids.begin(), ids.end(), rng);
      const int nw = 4;                       // warps per CTA
      constexpr int NSLOT = 4;
      const int chunks_per_warp = (9 * 8 * 1024) / chunk_bytes;   // ~72 KB per warp, like one (seq, head) item
      const int ctas = 512;
      const size_t need = (size_t)ctas * nw * chunks_per_warp;
      ids.resize(need);
      int* d_ids; cudaMalloc(&d_ids, need * 4);
      cudaMemcpy(d_ids, ids.data(), need * 4, cudaMemcpyHostToDevice);
      const size_t smem_per_cta = (size_t)nw * (NSLOT * chunk_bytes + 64);
      // limit residency to warps_per_sm by padding dynamic smem
      size_t smem = std::max(smem_per_cta, (size_t)(227 * 1024) / (warps_per_sm / nw) - 1024);
      cudaFuncSetAttribute(tma_read<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e9;
      for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        tma_read<NSLOT><<<ctas, nw * 32, smem>>>(buf, d_ids, chunks_per_warp, chunk_bytes, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
      }
      cudaError_t e = cudaGetLastError();
      printf("bulk %d KB chunks, %2d warps/SM, ring %d: %8.1f GB/s  (%.1f us for %.0f MB) %s\n", chunk_kb, warps_per_sm, NSLOT,
             (double)need * chunk_bytes / best / 1e6, best * 1e3, (double)need * chunk_bytes / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
      cudaFree(d_ids);
    }
  }
  return 0;
}
