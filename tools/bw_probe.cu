// HBM read-bandwidth probe for the paged-cache access pattern: each work item reads N random
// chunks of C bytes (C = 2/4/8/64 KB) out of a 4 GiB buffer with 16 B loads, maximal MLP, no math.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bw_probe tools/bw_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>

__global__ void read_chunks(const uint4* __restrict__ buf, const int* __restrict__ chunk_ids, int chunks_per_cta,
                            int chunk_vec, unsigned* sink) {
  // one CTA (256 threads) streams its list of chunks
  unsigned acc = 0;
  const int* ids = chunk_ids + (size_t)blockIdx.x * chunks_per_cta;
  for (int i = 0; i < chunks_per_cta; i += 4) {
    uint4 v[4][1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t base = (size_t)ids[i + j] * chunk_vec;
      for (int t = threadIdx.x; t < chunk_vec; t += blockDim.x) {
        uint4 x;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "l"(buf + base + t));
        acc += x.x ^ x.y ^ x.z ^ x.w;
      }
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const size_t total = (size_t)4 << 30;  // 4 GiB
  uint4* buf; cudaMalloc(&buf, total); cudaMemset(buf, 1, total);
  unsigned* sink; cudaMalloc(&sink, 4);
  std::mt19937 rng(0);
  for (int chunk_kb : {2, 4, 8, 16, 64, 1024}) {
    for (int random = 1; random >= 0; --random) {
      const size_t chunk_bytes = (size_t)chunk_kb << 10;
      const size_t nchunks = total / chunk_bytes;
      const size_t read_chunks_n = std::min<size_t>(nchunks, (size_t)(1u << 30) / chunk_bytes);  // read 1 GiB
      std::vector<int> ids(nchunks);
      std::iota(ids.begin(), ids.end(), 0);
      if (random) std::shuffle(ids.begin(), ids.end(), rng);
      ids.resize(read_chunks_n);
      const int chunk_vec = chunk_bytes / 16;
      const int threads = chunk_vec >= 256 ? 256 : chunk_vec;  // 2 KB chunk = 128 x 16 B
      int chunks_per_cta = 16;
      while (read_chunks_n / chunks_per_cta < 148 * 8 && chunks_per_cta > 4) chunks_per_cta /= 2;
      const int ctas = read_chunks_n / chunks_per_cta;
      int* d_ids; cudaMalloc(&d_ids, ids.size() * 4);
      cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e9;
      for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        read_chunks<<<ctas, threads>>>(buf, d_ids, chunks_per_cta, chunk_vec, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
      }
      printf("chunk %5d KB %s: %8.1f GB/s  (%d CTAs x %d thr, %d chunks/CTA)\n", chunk_kb, random ? "random    " : "sequential",
             (double)ctas * chunks_per_cta * chunk_bytes / best / 1e6, ctas, threads, chunks_per_cta);
      cudaFree(d_ids);
    }
  }
  return 0;
}
