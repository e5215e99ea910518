// Probe of tcgen05.mma.kind::f16 (bf16 operands) no-swizzle descriptors, K-major and MN-major, one shared image form:
//   off(r,c) = (c/8)*S_c + (r/8)*S_r + (r%8)*16 + (c%8)*2   [bytes]; K-major: r=m|n, c=k ; MN-major: r=k, c=m|n
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
struct Cfg { int a_mn, b_mn, N; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int a_Sc, b_Sc; };

__global__ void probe(const float* A, const float* B, float* out, Cfg c) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = (uint64_t*)smem;
  uint32_t* slot = (uint32_t*)(smem + 8);
  unsigned char* Ai = smem + 128;
  unsigned char* Bi = Ai + 32768;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 64 * 1024 / 4; e += blockDim.x) ((float*)Ai)[e] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  for (int e = tid; e < 128 * 16; e += blockDim.x) {
    int m = e / 16, k = e % 16;
    int r = c.a_mn ? k : m, cc = c.a_mn ? m : k;
    int off = (cc / 8) * c.a_Sc + (r / 8) * 128 + (r % 8) * 16 + (cc % 8) * 2;
    *(__nv_bfloat16*)(Ai + off) = __float2bfloat16(A[e]);
  }
  for (int e = tid; e < c.N * 16; e += blockDim.x) {
    int n = e / 16, k = e % 16;
    int r = c.b_mn ? k : n, cc = c.b_mn ? n : k;
    int off = (cc / 8) * c.b_Sc + (r / 8) * 128 + (r % 8) * 16 + (cc % 8) * 2;
    *(__nv_bfloat16*)(Bi + off) = __float2bfloat16(B[e]);
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = *slot;
  if (tid == 0) {
    uint64_t ad = make_desc(smem_u32(Ai), c.a_lbo, c.a_sbo), bd = make_desc(smem_u32(Bi), c.b_lbo, c.b_sbo);
    uint32_t idesc = make_idesc(128, c.N, c.a_mn, c.b_mn);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  uint32_t ok = 0; long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (tid == 0) printf("timeout\n"); break; }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp < 4) {
    for (int c0 = 0; c0 < c.N; c0 += 8) {
      uint32_t u[8];
      uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(u[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase));
}

int main() {
  const int N = 16;
  static float hA[128 * 16], hB[N * 16], ref[128 * N], got[128 * N];
  for (int i = 0; i < 128 * 16; ++i) hA[i] = (float)((i * 7 + 3) % 13) - 6.f;
  for (int i = 0; i < N * 16; ++i) hB[i] = (float)((i * 5 + 1) % 11) - 5.f;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 16; ++k) s += hA[m * 16 + k] * hB[n * 16 + k]; ref[m * N + n] = s; }
  float *dA, *dB, *dO;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dO, sizeof(got));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 64 * 1024);
  for (int a_mn = 0; a_mn < 2; ++a_mn) for (int b_mn = 0; b_mn < 2; ++b_mn) for (int variant = 0; variant < 2; ++variant) {
    Cfg c; c.a_mn = a_mn; c.b_mn = b_mn; c.N = N;
    // K-major operand image: r = m (16 groups of 8 -> 2048 B), c = k (2 groups): S_c = 2048+16 ; MN-major: r = k (2 groups -> 256 B), c = m: S_c = 256+16
    c.a_Sc = a_mn ? 272 : 2064; c.b_Sc = b_mn ? 272 : (128 * (N / 8) + 16);
    // variant 0 (CUTLASS reading): K-major LBO=S_c (k chunks), SBO=128 (row groups) ; MN-major SBO=S_c (MN groups), LBO=128 (k row groups)
    // variant 1: MN-major roles swapped
    if (!a_mn) { c.a_lbo = c.a_Sc; c.a_sbo = 128; } else if (variant == 0) { c.a_sbo = c.a_Sc; c.a_lbo = 128; } else { c.a_lbo = c.a_Sc; c.a_sbo = 128; }
    if (!b_mn) { c.b_lbo = c.b_Sc; c.b_sbo = 128; } else if (variant == 0) { c.b_sbo = c.b_Sc; c.b_lbo = 128; } else { c.b_lbo = c.b_Sc; c.b_sbo = 128; }
    if (variant == 1 && !a_mn && !b_mn) continue;
    cudaMemset(dO, 0xff, sizeof(got));
    probe<<<1, 128, 128 + 64 * 1024>>>(dA, dB, dO, c);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(got, dO, sizeof(got), cudaMemcpyDeviceToHost);
    double maxerr = 0; int nz = 0; for (int i = 0; i < 128 * N; ++i) { maxerr = fmax(maxerr, fabs((double)got[i] - ref[i])); nz += got[i] != 0.f; }
    printf("bf16 a_mn=%d b_mn=%d variant=%d : %s maxerr=%g nonzero=%d  got[0..3]=%g %g %g %g ref=%g %g %g %g\n", a_mn, b_mn, variant, cudaGetErrorString(e), maxerr, nz,
           got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
