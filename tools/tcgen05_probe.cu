// Probe for the tcgen05 (UMMA) attention tiles on sm_100a: S^T[128 tok x 16 q] = K[128 x 128] Q^T with the
// K operand read straight from raw paged K blocks ([16 chunk j][16 tok][8 halves], 4 KB per block):
// one M=128 tile = the token-halves (tokens 8h..8h+7) of 16 blocks, so the 8-row core matrices are
// 4096 B apart along M (SBO) and 256 B apart along K (LBO).  Prints the max error against a host
// reference for both LBO/SBO interpretations.
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/tcgen05_probe tools/tcgen05_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)layout << 61;     // 0 = no swizzle (interleave), 6 = 32B, 4 = 64B, 2 = 128B
  return d;
}

__global__ void __launch_bounds__(128) probe(const __half* __restrict__ kblocks,  // [16 blocks][2048]
                                             const __half* __restrict__ q,        // [16 n][128 d]
                                             float* __restrict__ out,             // [2 h][128 r][16 n]
                                             int swap_lbo_sbo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* sA = reinterpret_cast<__half*>(smem);                 // 64 KB raw K blocks
  __half* sB = reinterpret_cast<__half*>(smem + 65536);         // 4 KB Q tile, canonical K-major no-swizzle
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 4096 + 16);
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  for (int i = tid; i < 16 * 2048; i += 128) sA[i] = kblocks[i];
  // B: address(n, d) = (d/8)*256 + (n/8)*128 + (n%8)*16 + (d%8)*2 bytes
  for (int i = tid; i < 16 * 128; i += 128) {
    const int n = i / 128, d = i % 128;
    sB[((d / 8) * 256 + (n / 8) * 128 + (n % 8) * 16) / 2 + (d % 8)] = q[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> tensor core (async proxy)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // instruction descriptor: D=F32 (1<<4), A=B=F16 (0), K-major both, N=16 (16>>3 at bit 17), M=128 (128>>4 at bit 24)
  const uint32_t idesc = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) {
    for (int h = 0; h < 2; ++h) {
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t a_addr = smem_u32(sA) + h * 128 + ks * 512;   // two 256 B chunks per k-step
        const uint32_t b_addr = smem_u32(sB) + ks * 512;
        uint64_t ad, bd;
        if (!swap_lbo_sbo) {
          ad = make_desc(a_addr, /*lbo (K dir)*/ 256, /*sbo (M dir)*/ 4096, 0);
          bd = make_desc(b_addr, 256, 128, 0);
        } else {
          ad = make_desc(a_addr, 4096, 256, 0);
          bd = make_desc(b_addr, 128, 256, 0);
        }
        const uint32_t acc = ks > 0;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem + h * 16), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // everyone waits for the MMAs
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + h * 16;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 16; ++n) out[(h * 128 + warp * 32 + lane) * 16 + n] = __uint_as_float(r[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
  std::vector<__half> hk(16 * 2048), hq(16 * 128);
  srand(1);
  for (auto& x : hk) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : hq) x = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half *dk, *dq; float* dout;
  cudaMalloc(&dk, hk.size() * 2); cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dout, 2 * 128 * 16 * 4);
  cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 65536 + 4096 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // host reference: row r of tile h = token (block r/8, offset 8h + r%8); K[blk][j][t][e], d = 8j + e
  std::vector<float> ref(2 * 128 * 16);
  for (int h = 0; h < 2; ++h) for (int r = 0; r < 128; ++r) for (int n = 0; n < 16; ++n) {
    const int blk = r / 8, t = 8 * h + r % 8;
    float acc = 0;
    for (int d = 0; d < 128; ++d) acc += __half2float(hk[blk * 2048 + (d / 8) * 128 + t * 8 + d % 8]) * __half2float(hq[n * 128 + d]);
    ref[(h * 128 + r) * 16 + n] = acc;
  }
  for (int swap = 0; swap < 2; ++swap) {
    cudaMemset(dout, 0, 2 * 128 * 16 * 4);
    probe<<<1, 128, smem>>>(dk, dq, dout, swap);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> ho(2 * 128 * 16);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < ho.size(); ++i) { double d = fabs(ho[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 0.05) ++bad; }
    printf("swap_lbo_sbo=%d: %s  max|err|=%.4f  mismatches=%d/%zu   out[0..3]=%.3f %.3f %.3f %.3f  ref=%.3f %.3f %.3f %.3f\n", swap,
           cudaGetErrorString(e), maxerr, bad, ho.size(), ho[0], ho[1], ho[2], ho[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
