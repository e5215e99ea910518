// Micro-benchmark: cycles per tcgen05.mma for the operand layouts used by libadil_b200 (one CTA per SM, one issuing
// thread, R back-to-back MMAs on one accumulator, clock64 around issue .. commit completion).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_speed scripts/umma_speed.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7u) << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.u32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}
struct Cfg { int nacc; int kind;  /* 0 tf32, 1 bf16 */ int M, N, a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo, layout, a_step, b_step; int R; const char* name; };

__global__ void speed(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = (uint64_t*)smem;
  uint32_t* slot = (uint32_t*)(smem + 8);
  uint32_t* buf = (uint32_t*)(smem + 1024);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 200 * 1024 / 4; e += blockDim.x) buf[e] = 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
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
  if (warp == 0) {
    const uint32_t fmt = c.kind == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t abase = smem_u32(buf), bbase = smem_u32(buf) + 96 * 1024;
    uint64_t ad[4], bd[4];
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = make_desc(abase + ks * c.a_step, c.a_lbo, c.a_sbo, c.layout);
      bd[ks] = make_desc(bbase + ks * c.b_step, c.b_lbo, c.b_sbo, c.layout);
    }
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      long long t1 = 0;
      if (elect_one()) {
#pragma unroll 1
        for (int r = 0; r < c.R; r += 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (c.kind == 0)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                           ::"r"(tbase + (uint32_t)((ks % c.nacc) * 128)), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(r + ks)) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                           ::"r"(tbase + (uint32_t)((ks % c.nacc) * 128)), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"((uint32_t)(r + ks)) : "memory");
          }
        }
        t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"((uint32_t)(rep & 1)) : "memory");
      const long long t2 = clock64();
      if (blockIdx.x == 0 && t1 != 0) { out[2 * rep] = t1 - t0; out[2 * rep + 1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(speed, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 200 * 1024);
  const int R = 256;
  // strides as in libadil_b200: Sg = 1808 (B=100 -> Bp=112), Sd(TP=48) = 784, Sd(64) = 1040, Sv(synth,B=100) = 1680
  Cfg cfgs[] = {
    {1, 1, 64, 48, 1, 1, 128, 1808, 128, 1808, 0, 256, 256, R, "bf16 dD^T  M=64  N=48  A MN-major B MN-major (no swizzle)"},
    {1, 1, 128, 48, 1, 1, 128, 1808, 128, 1808, 0, 256, 256, R, "bf16 dD^T  M=128 N=48  A MN-major B MN-major (no swizzle)"},
    {1, 1, 128, 64, 0, 1, 1808, 128, 128, 784, 0, 3616, 256, R, "bf16 dv    M=128 N=64  A K-major  B MN-major (no swizzle)"},
    {1, 1, 128, 64, 0, 0, 1808, 128, 784, 128, 0, 3616, 1568, R, "bf16       M=128 N=64  A K-major  B K-major  (no swizzle)"},
    {1, 0, 128, 64, 0, 0, 1680, 128, 1040, 128, 0, 3360, 2080, R, "tf32 synth M=128 N=64  A K-major  B K-major  (no swizzle)"},
    {1, 1, 128, 64, 0, 0, 0, 1024, 0, 1024, 2, 32, 32, R, "bf16       M=128 N=64  K-major 128B swizzle (reference layout)"},
    {1, 0, 128, 64, 0, 0, 0, 1024, 0, 1024, 2, 32, 32, R, "tf32       M=128 N=64  K-major 128B swizzle (reference layout)"},
    {1, 1, 128, 64, 1, 1, 4096, 1024, 4096, 1024, 2, 2048, 2048, R, "bf16       M=128 N=64  MN-major 128B swizzle"},
    {1, 1, 128, 128, 0, 0, 0, 1024, 0, 1024, 2, 32, 32, R, "bf16       M=128 N=128 K-major 128B swizzle (reference layout)"},
    {1, 1, 128, 256, 0, 0, 0, 1024, 0, 1024, 2, 32, 32, R, "bf16       M=128 N=256 K-major 128B swizzle (reference layout)"},
    {1, 1, 128, 16, 0, 0, 0, 1024, 0, 1024, 2, 32, 32, R, "bf16       M=128 N=16  K-major 128B swizzle (reference layout)"},
  };
  for (int pass = 0; pass < 3; ++pass)
  for (auto& c : cfgs) {
    c.nacc = pass == 0 ? 1 : (pass == 1 ? 2 : 4);
    if (c.N > 128 && c.nacc > 1) continue;
    cudaMemset(d, 0, 64);
    speed<<<1, 128, 1024 + 200 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[6]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("nacc=%d %-70s %s  issue %6.1f cyc/mma  total %6.1f cyc/mma (3rd rep: %6.1f)\n", c.nacc, c.name, cudaGetErrorString(e),
           (double)h[2] / c.R, (double)h[3] / c.R, (double)h[5] / c.R);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
