// Cold vs warm instruction fetch of straight-line code on sm_100a (what a persistent kernel's prologue pays the first
// time it runs after other kernels have evicted its code from the instruction caches and from L2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/icache_probe scripts/icache_probe.cu
// Each probe kernel executes N straight-line, mostly independent FFMAs in ONE warp per CTA (grid = 148) and reports
// the global-timer span of CTA 0; "cold" = after a 512 MB memset (L2) and a polluter kernel with 128 KB of other code
// (SM instruction caches), "warm" = the same kernel launched again right away.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

template <int N, int SALT>
__global__ void chain(float* out, long long* span, float a, float b) {
  const long long t0 = gtime();
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = (float)(threadIdx.x + i + SALT);
#pragma unroll
  for (int i = 0; i < N; ++i) x[i & 7] = fmaf(x[i & 7], a, b + (float)(i % 3));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  const long long t1 = gtime();
  if (s == 123.456f) out[threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) { span[0] = t1 - t0; }
  if (blockIdx.x == 77 && threadIdx.x == 0) { span[1] = t1 - t0; }
}

__global__ void empty_kernel(float* out) { if (out == nullptr) printf("x"); }

template <int N>
void run(const char* name, float* out, long long* span, void* flush, size_t flush_bytes, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    long long h[2]; float ms;
    // cold
    cudaMemsetAsync(flush, rep, flush_bytes);
    chain<8192, 1><<<148, 32>>>(out, span + 2, 1.0001f, 0.5f);  // polluter: 128 KB of other code
    cudaEventRecord(e0);
    chain<N, 0><<<148, threads>>>(out, span, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaMemcpy(h, span, sizeof(h), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%s threads=%d N=%d (%d KB) cold: in-kernel %lld / %lld ns, events %.2f us", name, threads, N, N * 16 / 1024, h[0], h[1], ms * 1e3f);
    // warm
    cudaEventRecord(e0);
    chain<N, 0><<<148, threads>>>(out, span, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaMemcpy(h, span, sizeof(h), cudaMemcpyDeviceToHost);
    cudaEventElapsedTime(&ms, e0, e1);
    printf(" | warm: in-kernel %lld / %lld ns, events %.2f us\n", h[0], h[1], ms * 1e3f);
  }
}

int main() {
  float* out; long long* span; void* flush;
  const size_t fb = 512ull << 20;
  cudaMalloc(&out, 4096); cudaMalloc(&span, 64); cudaMalloc(&flush, fb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 4; ++rep) {
    float ms;
    cudaMemsetAsync(flush, rep, fb);
    chain<8192, 1><<<148, 32>>>(out, span + 2, 1.0001f, 0.5f);
    cudaEventRecord(e0); empty_kernel<<<148, 32>>>(out); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("empty kernel cold: events %.2f us", ms * 1e3f);
    cudaEventRecord(e0); empty_kernel<<<148, 32>>>(out); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf(" | warm: %.2f us", ms * 1e3f);
    cudaEventRecord(e0); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf(" | two events back to back: %.2f us\n", ms * 1e3f);
  }
  run<256>("chain", out, span, flush, fb, 32);
  run<1024>("chain", out, span, flush, fb, 32);
  run<4096>("chain", out, span, flush, fb, 32);
  run<4096>("chain", out, span, flush, fb, 512);
  run<8192>("chain", out, span, flush, fb, 32);
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
