// Micro-benchmark 2: tcgen05.mma issue/throughput with 1 or 2 issuing warps and 1 or 2 accumulators.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_speed2 scripts/umma_speed2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.u32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// NWARP issuing warps, each R MMAs; warp w uses accumulator (w * NACC + (i % NACC)) * 128
template <int NWARP, int NACC, int UNROLL>
__global__ void speed(int M, int N, int R, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = (uint64_t*)smem;   // [2]
  uint32_t* slot = (uint32_t*)(smem + 64);
  uint32_t* buf = (uint32_t*)(smem + 1024);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 128 * 1024 / 4; e += blockDim.x) buf[e] = 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = *slot;
  if (warp < NWARP) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t abase = smem_u32(buf) + warp * 32768, bbase = smem_u32(buf) + 65536 + warp * 32768;
    uint64_t ad[UNROLL], bd[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { ad[u] = make_desc(abase + u * 256, 2064, 128); bd[u] = make_desc(bbase + u * 256, 1040, 128); }
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      if (elect_one()) {
#pragma unroll 1
        for (int r = 0; r < R; r += UNROLL) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) mma(tbase + (uint32_t)((warp * NACC + (u % NACC)) * 128), ad[u], bd[u], idesc, (uint32_t)(r + u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + warp)) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar + warp)), "r"((uint32_t)(rep & 1)) : "memory");
      const long long t2 = clock64();
      if ((tid & 31) == 0) out[warp * 4 + rep] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}
template <int NWARP, int NACC, int UNROLL>
void run(int M, int N, long long* d) {
  const int R = 256;
  cudaFuncSetAttribute(speed<NWARP, NACC, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 128 * 1024);
  cudaMemset(d, 0, 64);
  speed<NWARP, NACC, UNROLL><<<1, 128, 1024 + 128 * 1024>>>(M, N, R, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("warps=%d acc/warp=%d unroll=%d M=%3d N=%3d: %s  %6.1f cyc per MMA per warp (warp1: %6.1f) -> %6.1f cyc per MMA overall\n", NWARP, NACC, UNROLL, M, N,
         cudaGetErrorString(e), (double)h[2] / R, (double)h[6] / R, (double)(h[2] > h[6] ? h[2] : h[6]) / (R * NWARP));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  const int Ns[5] = {16, 48, 64, 128, 256};
  for (int i = 0; i < 5; ++i) {
    const int N = Ns[i];
    run<1, 1, 1>(128, N, d);
    run<1, 1, 4>(128, N, d);
    run<1, 1, 8>(128, N, d);
    if (N <= 128) { run<1, 2, 4>(128, N, d); run<2, 1, 4>(128, N, d); run<2, 2, 4>(128, N, d); }
  }
  run<1, 1, 4>(64, 48, d);
  run<2, 1, 4>(64, 48, d);
  return 0;
}
