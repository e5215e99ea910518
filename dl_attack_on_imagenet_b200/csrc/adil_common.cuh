// Shared device/host helpers for libadil_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "adil_b200.h"

namespace adil {

constexpr int kMaxC = ADIL_MAX_CHANNELS;

// Per-channel constants passed by value in kernel parameters (mean/std of demo_dL_attack.py:55).
struct ChannelConsts {
  float mean[kMaxC];
  float stdv[kMaxC];
  float rstd[kMaxC];  // fp32(1 / std): reciprocal seed for div_by_const
  int C;
  int hw;
  int use;  // 0: identity
};

// AdamW scalars of one update, evaluated on the host in double like torch/optim/adam.py:526-547 and rounded to
// fp32 where torch hands a Python scalar to an fp32 elementwise op.
struct AdamwDev {
  float decay;      // 1 - lr*wd
  float w1;         // 1 - beta1
  float beta2;      // beta2
  float w2;         // 1 - beta2
  float bc2_sqrt;   // sqrt(1 - beta2^t)
  float eps;        // eps
  float neg_step;   // -(lr / (1 - beta1^t))
  float rbc2_sqrt;  // fp32(1 / bc2_sqrt): reciprocal seed of the exact division below
  int lerp_hi;      // 1 when (1-beta1) >= 0.5: torch's lerp switches formula
};

// a / c for a loop-invariant divisor c whose correctly rounded reciprocal rc is known: one multiply and two FMAs give
// the correctly rounded IEEE quotient (Markstein), i.e. the same bits as __fdiv_rn / torch's true division, at a third
// of the instructions.  Inputs here are finite and far from the denormal range.
__device__ __forceinline__ float div_by_const(float a, float c, float rc) {
  const float q = __fmul_rn(a, rc);
  const float r = __fmaf_rn(-q, c, a);
  return __fmaf_rn(r, rc, q);
}

// One AdamW element update; op order and roundings follow the torch sequence
//   p.mul_(1-lr*wd); m.lerp_(g,1-b1); s.mul_(b2).addcmul_(g,g,1-b2); denom=(s.sqrt()/bc2_sqrt).add_(eps);
//   p.addcdiv_(m, denom, -step_size)                                        (adil.py:154,186)
__device__ __forceinline__ void adamw_update(float& p, float& m, float& s, float g, const AdamwDev& h) {
  p = __fmul_rn(p, h.decay);
  float diff = __fsub_rn(g, m);
  m = h.lerp_hi ? __fsub_rn(g, __fmul_rn(diff, __fsub_rn(1.0f, h.w1))) : __fmaf_rn(h.w1, diff, m);
  s = __fmul_rn(s, h.beta2);
  s = __fmaf_rn(__fmul_rn(h.w2, g), g, s);
  float denom = __fadd_rn(div_by_const(__fsqrt_rn(s), h.bc2_sqrt, h.rbc2_sqrt), h.eps);
  p = __fmaf_rn(h.neg_step, __fdiv_rn(m, denom), p);
}

// Dictionary variant (P*K elements per step, the bulk of the path's arithmetic): same update with the special-function
// unit -- sqrt.approx (rel. error 2^-23; a subnormal second moment flushes to 0, where eps dominates the denominator
// anyway), a multiply by the rounded reciprocal of bc2_sqrt and rcp.approx (1 ulp) * m instead of the IEEE sequences
// above (14 instead of 36 instructions per element).  The result differs from torch's
// by a few ulp of the UPDATE, i.e. < 1e-8 absolute on D at lr = 0.01 (bound 1e-5, tests hold it to 1e-6).
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float b) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return r;
}
__device__ __forceinline__ void adamw_update_fast(float& p, float& m, float& s, float g, const AdamwDev& h) {
  p = __fmul_rn(p, h.decay);
  const float diff = __fsub_rn(g, m);
  m = h.lerp_hi ? __fsub_rn(g, __fmul_rn(diff, __fsub_rn(1.0f, h.w1))) : __fmaf_rn(h.w1, diff, m);
  s = __fmaf_rn(__fmul_rn(h.w2, g), g, __fmul_rn(s, h.beta2));
  const float denom = __fmaf_rn(sqrt_approx(s), h.rbc2_sqrt, h.eps);
  p = __fmaf_rn(h.neg_step, __fmul_rn(m, rcp_approx(denom)), p);  // denom in [1e-8, ~1e19]: no range scaling needed
}

__device__ __forceinline__ float clamp1(float x) { return fminf(fmaxf(x, -1.0f), 1.0f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses (read-once / write-once data: x, g, out)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// 128-bit load of data this kernel also writes (no .nc), not kept in L1
__device__ __forceinline__ float4 ld_global4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// host-side error plumbing (adil_api.cu)
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();
bool is_host_pointer(const void* p);  // true for pageable / pinned host memory (index arrays may live on the host)
AdamwDev make_adamw(const adil_adamw_t* hp);
ChannelConsts make_consts(int C, int hw, const float* mean_host, const float* std_host, bool use);

// options of the backward entry points (adil_grad / adil_grad_dict_step flags)
struct GradOpts {
  int accumulate;      // ADIL_GRAD_ACCUMULATE_DD: dD2 += instead of dD2 =
  int keep_partials;   // ADIL_GRAD_KEEP_PARTIALS: leave the per-CTA code-gradient slabs in scratch (no reduction launch)
  int* nslabs_out;     // host: number of slabs written (keep_partials)
  const float* delta;  // l2 penalty: gx += l2_coef * delta (FMA path only)
  float l2_coef;
};

// FMA-path launchers (adil_fma.cu)
int launch_synth_fma(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2,
                     const float* v, const int64_t* v_index, float* codes_out, int B, int P, int K,
                     const ChannelConsts& cc, float eps, int flags, cudaStream_t st);
int launch_grad_fma(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                    const float* v, const int64_t* v_index, int B, int P, int K, const ChannelConsts& cc,
                    const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes, const GradOpts& opt,
                    cudaStream_t st);
int grad_fma_max_batch(int K, bool want_dD, bool want_dv);

// tcgen05-path launchers (adil_tc.cu)
bool tc_synth_ok(int B, int P, int K, int hw);
bool tc_grad_ok(int B, int P, int K, int hw, bool want_dD, bool want_dv, bool fused);
int launch_synth_tc(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2,
                    const float* v, const int64_t* v_index, float* codes_out, int B, int P, int K,
                    const ChannelConsts& cc, float eps, int flags, cudaStream_t st);
int launch_grad_tc(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                   const float* v, const int64_t* v_index, int B, int P, int K, const ChannelConsts& cc,
                   const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes, const GradOpts& opt,
                   cudaStream_t st);

// number of partial [B,K] slabs the grad kernels may write into scratch
constexpr int kMaxGradCtas = 592;  // 148 SMs x 4

}  // namespace adil
