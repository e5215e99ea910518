// Probe: tcgen05.mma with the A operand in TMEM (kind::f16, M=128), A written by tcgen05.st 32x32b; correctness + speed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/umma_probe_ts scripts/umma_probe_ts.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
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
// A: [128 x KT] row-major, B: [N x KT] row-major (n, k); out[m, n] = sum_k A[m,k] B[n,k].  KT = 16 (two k-steps).
__global__ void probe(const float* A, const float* B, float* out, int N, int R, long long* tim) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = (uint64_t*)smem;
  uint32_t* slot = (uint32_t*)(smem + 8);
  unsigned short* Bi = (unsigned short*)(smem + 128);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int KT = 32, SB = 128 * (N / 8) + 16;
  for (int e = tid; e < 32 * 1024 / 2; e += blockDim.x) Bi[e] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  for (int e = tid; e < N * KT; e += blockDim.x) {
    const int n = e / KT, k = e % KT;
    Bi[(k >> 3) * (SB >> 1) + (n >> 3) * 64 + (n & 7) * 8 + (k & 7)] = (unsigned short)(__float_as_uint(B[e]) >> 16);
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = *slot;
  const uint32_t a_tmem = tbase + 256;   // A at columns [256, 256+KT)
  {  // thread t <-> lane t writes its row of A
    uint32_t r[16];
    for (int k = 0; k < KT / 2; ++k) r[k] = (__float_as_uint(A[tid * KT + 2 * k]) >> 16) | (__float_as_uint(A[tid * KT + 2 * k + 1]) & 0xffff0000u);
    const uint32_t ta = a_tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t bd0 = make_desc(smem_u32(Bi), SB, 128);
    const uint64_t bstep = (uint64_t)((2 * SB) >> 4);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                       ::"r"(tbase), "r"(a_tmem + ks * 8), "l"(bd0 + ks * bstep), "r"(idesc), "r"((uint32_t)(r | ks)) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    if (tid == 0) tim[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(tbase + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[tid * N + c0 + j] = __uint_as_float(u[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}
int main() {
  const int KT = 32;
  for (int N : {16, 64}) {
    float *hA = new float[128 * KT], *hB = new float[N * KT], *ref = new float[128 * N], *got = new float[128 * N];
    for (int i = 0; i < 128 * KT; ++i) hA[i] = (float)((i * 7 + 3) % 13) - 6.f;
    for (int i = 0; i < N * KT; ++i) hB[i] = (float)((i * 5 + 1) % 11) - 5.f;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < KT; ++k) s += hA[m * KT + k] * hB[n * KT + k]; ref[m * N + n] = s; }
    float *dA, *dB, *dO; long long* dT;
    cudaMalloc(&dA, 128 * KT * 4); cudaMalloc(&dB, N * KT * 4); cudaMalloc(&dO, 128 * N * 4); cudaMalloc(&dT, 8);
    cudaMemcpy(dA, hA, 128 * KT * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 32 * 1024);
    for (int R : {1, 128}) {
      cudaMemset(dO, 0xff, 128 * N * 4);
      probe<<<1, 128, 128 + 32 * 1024>>>(dA, dB, dO, N, R, dT);
      cudaError_t e = cudaDeviceSynchronize();
      long long t; cudaMemcpy(&t, dT, 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(got, dO, 128 * N * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int i = 0; i < 128 * N; ++i) maxerr = fmax(maxerr, fabs((double)got[i] - (R == 1 ? 1.0 : 1.0) * ref[i]));
      printf("TS tf32 N=%d R=%d: %s maxerr(vs single product)=%g got[0..3]=%g %g %g %g ref=%g %g %g %g  cycles/mma=%.1f\n", N, R, cudaGetErrorString(e),
             maxerr, got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3], (double)t / (2.0 * R));
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
