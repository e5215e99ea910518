// tcgen05 (5th-gen tensor core) kernels of the ADiL hot path: warp-specialised, persistent, TMA-fed pipelines.
//
//   synth_kernel  out[b,p] = f(x[b,p] + sum_k v[b,k] D[p,k])                        (adil.py:24-27, demo:22-25)
//   grad_kernel   dD[p,k]  = sum_b gx[b,p] v[b,k] ;  dv[b,k] = sum_p gx[b,p] D[p,k]   (autograd of the above)
//                 optionally fused with AdamW(D) + clamp                             (adil.py:186,188)
//
// Numerics.  The contractions run on the tensor cores with fp32-grade split operands and fp32 accumulators in TMEM:
//   synthesis : 3xTF32.  x = hi + lo, hi = x & 0xffffe000 (exact in TF32), lo = x - hi; MMAs lo*hi + hi*lo + hi*hi.
//   backward  : bf16x3.  x = b0 + b1 + b2 (exact residual split, last term rounded); six MMAs (2,0)(0,2)(1,1)(1,0)
//               (0,1)(0,0).  bf16 because the gradient tile is contracted over images for dD and over pixels for dv:
//               one shared-memory image of a 16-bit operand can be read in both majors, a TF32 operand cannot.
//
// Roles (one persistent CTA per SM; hand-offs between roles are mbarriers):
//   synthesis (800 threads)
//     warps 0..15   workers: dictionary tile split (fp32 -> hi/lo images, UMMA canonical no-swizzle layout) and the
//                   epilogue (TMEM -> registers -> +x, clamps, Normalize -> shared-memory staging)
//     warp 16       MMA issuer (one elected lane issues tcgen05.mma and commits to an mbarrier) + dictionary-tile TMA
//     warps 17..24  I/O: every global access of the image rows (cp.async in, coalesced 128-bit streaming stores out)
//   backward (896 threads)
//     warps 0..15   workers: operand staging only (gradient rows -> three bf16 images, TMA-landed dictionary rows ->
//                   three bf16 images of D / std)
//     warps 16..23  epilogue: dD^T accumulator -> flat [pixel][atom] tile in shared memory -> AdamW + clamp with the
//                   moments prefetched from global memory into registers -> coalesced stores (or, plain dD output,
//                   into the stage that the loader writes out with one TMA bulk store)
//     warp 26 / 27  MMA issuer / loader (1-D TMA bulk copies of the contiguous D tiles, bulk stores of dD tiles);
//                   warps 24, 25 only place them on schedulers 2 and 3
// Tiles of TP pixels are assigned round-robin (tile = blockIdx.x + i * gridDim.x).  Per tile the workers stage the
// operands of tile i while the MMAs of tile i-1 run and the epilogue of tile i-2 finishes; operand images and
// accumulators are double-buffered.
//
// Shared-memory operand images (no swizzle): 128-byte core matrices of 8 "rows" x 16 bytes,
//     off(r, c) = (c / E) * S + (r / 8) * 128 + (r % 8) * 16 + (c % E) * sizeof(elem)      E = 16 / sizeof(elem)
// r = index along which 8 rows form a core matrix, c = contiguous index.  The same image serves a K-major operand
// (r = M/N index, c = contraction index: LBO = S, SBO = 128) and an MN-major operand (r = contraction index,
// c = M/N index: LBO = 128, SBO = S).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "adil_common.cuh"

namespace adil {

int launch_reduce_partials(float* dvb, const float* partial, int n, int nslabs, int K, int ld_out, cudaStream_t st,
                           int slab_step = 1);

namespace {

// Phase timing (debug builds only, -DADIL_TIMING): thread 0 of every CTA accumulates clock64 deltas per pipeline phase;
// the launchers print the per-CTA-tile averages to stderr.  Never compiled into the product library.  Synthesis
// kernel only: the counter arrays cost ~30 registers, and the backward kernel (capped at 72) spills under them and slows
// down several-fold -- it carries the cheap -DADIL_CHAIN stamps below instead.
#ifdef ADIL_TIMING
__device__ long long g_tim[16];
#define TIM_DECL long long tim_last = clock64(); long long tim_acc[12] = {0}
#define TIM(i) do { if (tid == 0) { const long long t_ = clock64(); tim_acc[i] += t_ - tim_last; tim_last = t_; } } while (0)
#define TIM_FLUSH(n) do { if (tid == 0) { for (int i_ = 0; i_ < 12; ++i_) atomicAdd((unsigned long long*)&g_tim[i_], (unsigned long long)tim_acc[i_]); atomicAdd((unsigned long long*)&g_tim[12], (unsigned long long)(n)); } } while (0)
__device__ long long g_stamp[8];
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_stamp[i] = gtime(); } while (0)

#else
#define TIM_DECL
#define TIM(i)
#define TIM_FLUSH(n)
#endif
// Hand-off chain of one tile (debug builds only, -DADIL_CHAIN; cheap enough not to disturb the schedule): global-timer
// stamps of CTA 5, tile 6, written by lane 0 of whichever warp passes the probe.
#ifdef ADIL_CHAIN
__device__ long long g_chain[16];
__device__ __forceinline__ long long gtime2() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define CHAIN(i, tile) do { if (blockIdx.x == 5 && (tile) == 6 && (threadIdx.x & 31) == 0) g_chain[i] = gtime2(); } while (0)
__device__ long long g_stamp[16];
#define STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_stamp[i] = gtime2(); } while (0)
__device__ long long g_sstamp[16];
#define SSTAMP(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_sstamp[i] = gtime2(); } while (0)
__device__ long long g_wstamp[16];  // backward kernel, CTA 0: lane 0 of whichever warp passes the probe
#define WSTAMP(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_wstamp[i] = gtime2(); } while (0)
#else
#define CHAIN(i, tile)
#define STAMP(i)
#define SSTAMP(i)
#define WSTAMP(i)
#endif

constexpr int NW = 16;              // worker warps
constexpr int NT = NW * 32;         // worker threads
constexpr int WARP_MMA = NW;        // issuer warp
constexpr int WARP_LOAD = NW + 1;   // loader warp
constexpr int NE = 8;               // backward: epilogue warps 16..23 (TMEM quadrant = warp & 3, pixel half = (warp - 16) / 4)
constexpr int NEP_MAX = 10;         // backward, fused step: warps 16..25 share the AdamW pass when it has at most 3 items per
                                    // thread (24, 25 only do that); with 4 items per thread: the eight epilogue warps
constexpr int WARP_EPI = NW;
// The active epilogue warps sit on schedulers 0 and 1 (quadrant = warp % 4 = scheduler) whenever K <= 64; the issuer
// and the loader go to schedulers 2 and 3 so that they do not queue behind them for issue slots.
constexpr int WARP_MMA_G = NW + NE + 2;   // backward: issuer warp (26)
constexpr int WARP_LOAD_G = NW + NE + 3;  // backward: loader warp (27)
constexpr int NTHREADS_GRAD = (NW + NE + 4) * 32;  // 896 threads -> at most 72 registers each
constexpr int NIO = 8;              // synthesis: warps 17..24 move the image rows (cp.async in, coalesced stores out)
constexpr int NTIO = NIO * 32;
constexpr int NTHREADS_SYNTH = NT + 32 + NTIO;  // 800 threads -> at most 80 registers each
constexpr int NS = 3;               // stages of the raw dictionary tiles
constexpr int NSX = 3;              // stages of the image-row tiles (synthesis)
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int HDR_BYTES = 256 + 2048;  // barriers + tmem slot | per-image x row offsets | per-image code rows

__host__ __device__ inline int rup(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp.  tcgen05.mma / cp.async.bulk are warp-uniform instructions (operands in uniform
// registers): under `elect.sync` ptxas issues them once, under a `lane == 0` test it wraps each one in a per-thread
// "waterfall" loop that costs ~100 cycles per instruction (measured: scripts/umma_speed*.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.u32 %0, 1;\n\t}\n" : "+r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait WITHOUT a suspend-time hint: the instruction itself blocks in hardware for an implementation-defined
// time.  With a hint ptxas emits PHASECHK + NANOSLEEP.SYNCS, which returns at once: measured 3.0 M spin iterations per
// launch -- 36 % of all executed instructions and 76 % utilisation of the XU pipe that AdamW's sqrt / rcp need.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (cudaErrorLaunchFailure), never as a hung GPU.
// try_wait suspends the warp in hardware for a while before it returns false, so the loop is cheap.
// (No message: a printf here kept threadIdx.x alive -- and spilled -- across the whole kernel, and as an out-of-line call it
// cost the fused step 9-18 us at K = 64 / 100 through the register allocation around the call sites.  A reload from
// local memory queues behind the stores of the AdamW pass: this kernel must stay free of spills.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (TMA store): full-line writes whatever the row pitch of the tile
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
// shared -> global bulk reduction (TMA): global[i] += shared[i] (fp32); chunks of a batch accumulate in stream order
__device__ __forceinline__ void bulk_red_add_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// `n` floats (1, 2 or 4: 4 / 8 / 16 bytes, both addresses aligned to the size)
__device__ __forceinline__ void cp_async_floats(void* smem_dst, const void* gsrc, int n) {
  if (n == 4) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
  } else if (n == 2) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Gather of `nrows` (<= ROWS) rows of `cpr` copies of CW floats each by one warp: lane l holds the global address of
// row (l % ROWS) in `myrow`; the rows are handed round by shuffle and land `dpitch` floats apart.  A rolled loop with a
// minimal body on purpose: code that runs once costs its instruction fetch (cold: ~0.5 us per KB of straight-line
// code, measured by unrolling this), and a fat body (parameter loads, width dispatch and address arithmetic per row)
// ran at 0.19 us per row.
template <int CW>
__device__ __forceinline__ void gather_rows(float* dst, int dpitch, const float* myrow, int nrows, int cpr, int lane) {
#pragma unroll 1
  for (int i = 0; i < nrows; ++i) {
    const float* src = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(myrow), i));
#pragma unroll 1
    for (int c = lane; c < cpr; c += 32) cp_async_floats(dst + c * CW, src + c * CW, CW);
    dst += dpitch;
  }
}
template <int ROWS>
__device__ __forceinline__ void gather_rows_cw(float* dst, int dpitch, const float* myrow, int nrows, int K, int cw, int lane) {
  nrows = nrows < ROWS ? nrows : ROWS;
  if (cw == 4) gather_rows<4>(dst, dpitch, myrow, nrows, K >> 2, lane);
  else if (cw == 2) gather_rows<2>(dst, dpitch, myrow, nrows, K >> 1, lane);
  else gather_rows<1>(dst, dpitch, myrow, nrows, K, lane);
}
// the mbarrier receives one arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier id 2 = worker-only barrier (hand-offs between roles use mbarriers)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// register rebalancing between warpgroups (all four warps of an aligned group of four must execute it)
#ifndef ADIL_V_SETMAXNREG  // (measured: no gain here -- the issuer spills at 24 registers, the epilogue warps do not need 96)
template <int R> __device__ __forceinline__ void reg_alloc() {}
template <int R> __device__ __forceinline__ void reg_dealloc() {}
#else
template <int R> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
#endif

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// UMMA shared-memory matrix descriptor, no swizzle (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;
}
// instruction descriptors, fp32 accumulate (cute/arch/mma_sm100_desc.hpp: InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// A operand in TMEM (lane = M row, one 32-bit column per TF32 element), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns <- 8 registers per thread (thread t <-> TMEM lane base+t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])),
               "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])),
               "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in TMEM (lane = M row, one 32-bit column per PAIR of bf16 contraction elements, even element low)
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* r) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread t <-> TMEM lane base+t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __fsub_rn(x, hi);  // exact
}
// x = t0 + t1 + t2 + O(2^-24 x): t0, t1 are the truncated high halves of x and of the (exact) residuals; the words
// returned hold the bf16 term in their UPPER 16 bits (ready for PRMT packing).
__device__ __forceinline__ void split_bf16x3(float x, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  w0 = __float_as_uint(x);
  const float r1 = __fsub_rn(x, __uint_as_float(w0 & 0xffff0000u));   // exact
  w1 = __float_as_uint(r1);
  const float r2 = __fsub_rn(r1, __uint_as_float(w1 & 0xffff0000u));  // exact
  w2 = __float_as_uint(r2) + 0x8000u;                                  // round the last term to nearest
}
__device__ __forceinline__ uint32_t pack_hi16(uint32_t lo_word, uint32_t hi_word) {
  return __byte_perm(lo_word, hi_word, 0x7632);  // {hi_word[31:16], lo_word[31:16]}
}

// e / d for the loop-invariant divisor d whose magic is `mul` = ceil(2^32 / d) (0: d == 1); exact for e * d < 2^32
__device__ __forceinline__ int div_magic_dev(int e, unsigned mul) { return mul ? (int)__umulhi((unsigned)e, mul) : e; }

__host__ __device__ inline int img_stride(int R) { return 128 * ((R + 7) / 8) + 16; }  // +16 de-phases the banks

// Channel constants of a tile that spans at most two channels: pixels [0, bnd) of the tile are channel c0.  Tiles are
// visited in increasing order, so the constants are cached in registers and reloaded only when a tile leaves the
// cached channel (twice per launch for 3 channels) -- a per-tile integer division or a scan over the channels costs
// more than 10 % of the kernel's instructions.
struct TileChan {
  int lo, hi;  // absolute pixel range [lo, hi) of channel c0; tile-local boundary bnd = hi - p0
  int bnd;
  float mean0, std0, rstd0, mean1, std1, rstd1;
  float nmr0, nmr1;  // -mean * (1/std): Normalize as one FFMA, x * rstd + nmr
};
__device__ __forceinline__ void tile_chan_init(TileChan& t) {
  t.lo = 0; t.hi = 0; t.bnd = 0;
  t.mean0 = t.mean1 = t.nmr0 = t.nmr1 = 0.0f;
  t.std0 = t.std1 = t.rstd0 = t.rstd1 = 1.0f;
}
__device__ __forceinline__ void tile_chan_update(TileChan& t, const ChannelConsts& cc, int p0) {
  if (p0 < t.lo || p0 >= t.hi) {
    int c0 = 0;
    for (int c = 1; c < cc.C; ++c) c0 += (p0 >= c * cc.hw) ? 1 : 0;
    const int c1 = min(c0 + 1, kMaxC - 1);
    t.lo = c0 * cc.hw;
    t.hi = t.lo + cc.hw;
    t.mean0 = cc.mean[c0]; t.std0 = cc.stdv[c0]; t.rstd0 = cc.rstd[c0];
    t.mean1 = cc.mean[c1]; t.std1 = cc.stdv[c1]; t.rstd1 = cc.rstd[c1];
    t.nmr0 = -t.mean0 * t.rstd0;
    t.nmr1 = -t.mean1 * t.rstd1;
  }
  t.bnd = t.hi - p0;
}

// synthesis: staged code row (fp32, `crow`) -> hi / lo TF32 terms in tensor memory, 8-atom chunks c = c0, c0 + 4, ...
// (the four warps of a TMEM quadrant share the chunks); atoms >= K and rows of images >= B are written as zeros.
// Eight consecutive code values of one row, straight from global memory (the register path of the synthesis kernel: code
// rows that do not fit the staging buffer, i.e. more than ~100 atoms).  With 16-byte-aligned rows two 128-bit loads
// instead of eight scalar ones: a worker thread issues 8 requests per round instead of 32 (K = 200: the scalar gather
// was a fifth of all stall samples of the kernel, profiles/r02c5_synth_kernel_ncu_summary.txt).
__device__ __forceinline__ void load_codes8(const float* vrow, int k0, int K, bool ok, bool vec4, float (&out)[8]) {
  if (vec4) {
    float4 lo4 = make_float4(0.f, 0.f, 0.f, 0.f), hi4 = lo4;
    if (ok && k0 < K) lo4 = __ldg(reinterpret_cast<const float4*>(vrow + k0));
    if (ok && k0 + 4 < K) hi4 = __ldg(reinterpret_cast<const float4*>(vrow + k0 + 4));
    out[0] = lo4.x; out[1] = lo4.y; out[2] = lo4.z; out[3] = lo4.w;
    out[4] = hi4.x; out[5] = hi4.y; out[6] = hi4.z; out[7] = hi4.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = (ok && k0 + i < K) ? __ldg(vrow + k0 + i) : 0.0f;
  }
}

// (the export of the same values as rows of the contiguous [B, K] block the backward kernel of the step fetches)
__device__ __forceinline__ void store_codes8(float* orow, int k0, int K, bool vec4, const float (&val)[8]) {
  if (vec4) {
    if (k0 < K) *reinterpret_cast<float4*>(orow + k0) = make_float4(val[0], val[1], val[2], val[3]);
    if (k0 + 4 < K) *reinterpret_cast<float4*>(orow + k0 + 4) = make_float4(val[4], val[5], val[6], val[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (k0 + i < K) orow[k0 + i] = val[i];
  }
}

template <int CW>
__device__ __forceinline__ void synth_codes_to_tmem(const float* crow, bool b_ok, int K, int Kp8, int c0, uint32_t lane_base) {
  const int nchunks = Kp8 >> 3;
#pragma unroll 1
  for (int c = c0; c < nchunks; c += 4) {
    float val[8];
    const int k0 = 8 * c;
    if (CW == 4) {
      const float4 v0 = *reinterpret_cast<const float4*>(crow + k0);  // (k0 < K always: c < ceil(K / 8))
      const float4 v1 = (k0 + 4 < K) ? *reinterpret_cast<const float4*>(crow + k0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      val[0] = v0.x; val[1] = v0.y; val[2] = v0.z; val[3] = v0.w;
      val[4] = v1.x; val[5] = v1.y; val[6] = v1.z; val[7] = v1.w;
    } else if (CW == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = (k0 + 2 * i < K) ? *reinterpret_cast<const float2*>(crow + k0 + 2 * i) : make_float2(0.f, 0.f);
        val[2 * i] = t.x; val[2 * i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) val[i] = (k0 + i < K) ? crow[k0 + i] : 0.0f;
    }
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_tf32(b_ok ? val[i] : 0.0f, hi[i], lo[i]);
    tmem_st8(lane_base + (uint32_t)(8 * c), hi);
    tmem_st8(lane_base + (uint32_t)(Kp8 + 8 * c), lo);
  }
}

// =========================================================================================================
// synthesis:  acc[b, p] = sum_k v[b,k] D[p,k]      M = 128 image lanes, N = TP pixels, contraction over atoms
//   A = batch codes (hi/lo) in TENSOR MEMORY: lane = image, column = atom (written once per CTA with tcgen05.st) --
//       a loop-invariant operand read from shared memory would cost 4 KB of shared-memory bandwidth per MMA
//       (measured: 48.7 -> 32.8 cycles per M=128,N=64 MMA, scripts/umma_probe_ts.cu) and 50 KB of capacity
//   B = D tile images        r = pixel, c = atom   K-major      (rebuilt per tile from the TMA-landed raw tile,
//       double-buffered so that the split of tile i+1 overlaps the MMAs of tile i)
// The image-row tile x[b, p0:p0+TP] lands by cp.async in a padded [B][TP+4] buffer; the epilogue thread of image b
// reads its accumulator row, applies +x / clamps / Normalize in place, and the finished tile goes out as coalesced
// 128-bit streaming stores.
// =========================================================================================================
struct SynthArgs {
  float* out;
  float* delta;
  float* codes_out;      // [B, K] (may be NULL): the code rows of the batch, contiguous, for the backward kernel of the step
  const float* x;
  const int64_t* xidx;
  const float* D2;
  const float* v;
  const int64_t* vidx;
  int hx_on, hv_on;      // the batch indices travel in the kernel parameters (host index arrays hx[], hv[] at the end,
                         // B <= 128): no dependent cold miss on an index array at the top of the kernel
  int B, P, K;
  int Kp8;               // contraction length: round_up(K, 8)
  int Sd;                // byte stride between 4-atom groups of a dictionary image
  int dimg;              // floats per dictionary image
  int raw_floats;        // floats per raw stage
  int vk;                // vector width of the dictionary split: 4, 2 or 1 (K % vk == 0)
  unsigned kdiv;         // ceil(2^32 / (K / vk))
  uint32_t tmem_cols;
  float eps;
  int flags;
  int stage_codes;       // the code rows are gathered by cp.async into the second image buffer (they fit), pitch cpitch
  int cpitch;            // floats between staged code rows: K, or K + 4 when K % 8 == 0 (conflict-free 128-bit reads)
  int cw;                // floats per cp.async of the gather / per shared-memory read: 4, 2 or 1
  int rpw;               // rows of the gather per warp: ceil(B / 25)
  ChannelConsts cc;
  int hx[128], hv[128];
};

// TRAIN = the learning-loop configuration (x and out given, no delta output, no clamps): those branches vanish.
// STAGE = the code rows are staged in shared memory by cp.async (they fit the second image buffer: up to ~100 atoms);
// otherwise every worker thread fetches its share straight into registers.  A template parameter, not a run-time branch:
// the path that is not taken still costs its instruction fetch when the kernel starts with cold instruction caches, as it
// does inside a step (measured at config 2: the 128-bit register path compiled into the same kernel took the synthesis
// from 35.3 to 36.8 us in-step -- and nothing in an isolated loop, scripts/gpu_r02_v.sh).
template <int TP, bool TRAIN, bool STAGE>
__global__ void __launch_bounds__(NTHREADS_SYNTH, 1) synth_kernel(const SynthArgs a) {
  constexpr int XP = TP + 4;       // floats per staged image row (pitch 16 bytes off a multiple of 128)
  constexpr int Q4 = TP / 4;       // float4 per image row
  constexpr int NCG = TP / 16;     // 16-column groups of the accumulator
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // (contiguous, in this order: the issuer warp initialises them one per lane)
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(smem_raw);  // [NS]
  uint64_t* empty_raw = full_raw + NS;                         // [NS]
  uint64_t* full_x = empty_raw + NS;                           // [NSX] rows landed / stage free (I/O warps -> workers)
  uint64_t* out_ready = full_x + NSX;                          // [NSX] finished tile staged (workers -> I/O warps)
  uint64_t* mma_done = out_ready + NSX;                        // [2] MMAs of the tiles using buffer 0 / 1 retired
  uint64_t* staged = mma_done + 2;                             // [2] dictionary images of buffer 0 / 1 written
  uint64_t* codes_ready = staged + 2;                          // [1] batch codes written to tensor memory (workers -> issuer)
  uint64_t* codes_landed = codes_ready + 1;                    // [1] code rows of the batch gathered into shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(codes_landed + 1);
  long long* xoff_s = reinterpret_cast<long long*>(smem_raw + 256);  // [128]
  float* raw = reinterpret_cast<float*>(smem_raw + HDR_BYTES);       // [NS][raw_floats]
  float* Dimg = raw + NS * a.raw_floats;                             // [2 buffers][hi, lo][dimg]
  float* xs = Dimg + 4 * a.dimg;                                     // [NSX][B][XP]
  float* cstage = Dimg + 2 * a.dimg;                                 // [B][cpitch] code rows, until they are in TMEM

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, P = a.P, B = a.B;
  const int ntiles = (P + TP - 1) / TP;
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool need_x = (a.x != nullptr) && (a.out != nullptr);
  const int xstage = B * XP;
  const bool ragged = (P % TP) != 0;
  if (tid == 0) SSTAMP(0);
  const int iot = tid - WARP_LOAD * 32;
  auto request_x = [&](int j) {
    if (need_x && j < my_tiles) {
      const int p0 = (blockIdx.x + j * gridDim.x) * TP;
      float* dst = xs + (j % NSX) * xstage;
      for (int e = iot; e < B * Q4; e += NTIO) {
        const int b = e / Q4, col = (e - b * Q4) * 4;
        if (p0 + col < P) cp_async16(dst + b * XP + col, a.x + xoff_s[b] + p0 + col);
      }
    }
  };
  auto load_x = [&](int j) {
    request_x(j);
    cp_async_arrive_noinc(full_x + (j % NSX));  // rows landed (or, without x, simply: stage free)
  };
  // raw dictionary tile `it` -> stage it % NS by one TMA bulk copy (issuer's elected lane)
  auto load_D = [&](int it) {
    const int p0 = (blockIdx.x + it * gridDim.x) * TP;
    const uint32_t bytes = (uint32_t)(min(TP, P - p0) * K * 4);
    mbar_expect_tx(full_raw + it % NS, bytes);
    bulk_g2s(raw + (it % NS) * a.raw_floats, a.D2 + (size_t)p0 * K, bytes, full_raw + it % NS);
  };
  // contraction padding k in [K, Kp8) of a dictionary image buffer must be zero (everything else is rewritten per tile)
  auto zero_padding = [&](float* Dhi, int t0, int nt) {
    const int npad = a.Kp8 - K;
    float* Dlo = Dhi + a.dimg;
    for (int e = t0; e < TP * npad; e += nt) {
      const int p = e / npad, k = K + (e - p * npad);
      const int o = (k >> 2) * (a.Sd >> 2) + (p >> 3) * 32 + (p & 7) * 4 + (k & 3);
      Dhi[o] = 0.0f;
      Dlo[o] = 0.0f;
    }
  };

  // ---- Kernel entry.  Everything tile 0 needs is one cold miss away and depends on nothing but the kernel parameters:
  // every role puts its requests out before any set-up work -- the I/O warps the image rows of tile 0, the issuer the
  // first dictionary tile (TMA), the workers the code rows of the batch (cp.async gather into the second image buffer;
  // when they do not fit there: straight into registers).  The requests of the later tiles follow after the set-up
  // barrier so that they queue behind these. ----
  float vv[4][8];
  if constexpr (STAGE) {
    // code rows: warp w gathers rows [w rpw, (w + 1) rpw) -- a warp keeps only about four cp.async in flight (8 rows per
    // warp took two latencies), so the rows are spread over all the warps; lane l holds the address of row
    // w rpw + (l & 7) (one line of the parameters or of the index array per warp)
    const int r0 = warp * a.rpw;
    if (r0 < B) {
      const int rb = min(r0 + (lane & 7), B - 1);
      const long long row = a.hv_on ? (long long)a.hv[rb] : (a.vidx ? (long long)a.vidx[rb] : (long long)rb);
      gather_rows_cw<8>(cstage + r0 * a.cpitch, a.cpitch, a.v + row * K, min(a.rpw, B - r0), K, a.cw, lane);
    }
  }
  if (warp < NW) {
    if constexpr (!STAGE) {
      // worker thread <-> image b = 32*quad + lane (its TMEM lane); the warps of a quadrant share the 8-atom chunks
      const int b = (warp & 3) * 32 + lane;
      const int bi = min(b, B - 1);
      const long long row = a.hv_on ? (long long)a.hv[bi] : (a.vidx ? (long long)a.vidx[bi] : (long long)bi);
      const float* vrow = a.v + row * K;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) load_codes8(vrow, 8 * ((warp >> 2) + 4 * ci), K, b < B, a.cw == 4, vv[ci]);
      // the last CTA (never one with more tiles than the others) leaves the rows behind as the contiguous [B, K] block the
      // backward kernel of the step fetches with one bulk copy (a launch of its own before: 5-7 us per step at K = 200)
      if (a.codes_out != nullptr && blockIdx.x == gridDim.x - 1 && b < B) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) store_codes8(a.codes_out + (size_t)b * K, 8 * ((warp >> 2) + 4 * ci), K, a.cw == 4, vv[ci]);
      }
    }
    if (tid == 0) SSTAMP(13);
#pragma unroll 1
    for (int buf = 0; buf < (STAGE ? 1 : 2); ++buf) zero_padding(Dimg + buf * 2 * a.dimg, tid, NT);
    if (ragged) {  // stale rows of the raw stages must stay finite; the tiles are requested after the set-up barrier
      float4* z = reinterpret_cast<float4*>(raw);
      for (int e = tid; e < (NS * a.raw_floats) >> 2; e += NT) z[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      fence_proxy_async();  // the zero fill (generic proxy) must be ordered before the TMA writes
    }
  } else if (warp == WARP_MMA) {
    if (lane < 2 * NS + 2 * NSX + 6) {  // one mbarrier per lane
      uint32_t cnt = NW;                                                     // empty_raw, out_ready, staged, codes_ready
      if (lane == 2 * NS + 2 * NSX + 5) cnt = NTHREADS_SYNTH;                // codes_landed
      else if (lane < NS) cnt = 1;                                           // full_raw
      else if (lane >= 2 * NS && lane < 2 * NS + NSX) cnt = NTIO;            // full_x
      else if (lane >= 2 * NS + 2 * NSX && lane < 2 * NS + 2 * NSX + 2) cnt = 1;  // mma_done
      mbar_init(full_raw + lane, cnt);
      fence_mbar_init();
    }
    __syncwarp();
    const bool leader = elect_one();
    if (leader && !ragged && my_tiles > 0) load_D(0);
    __syncwarp();
    tmem_alloc(tmem_slot, a.tmem_cols);
    if (lane == 0) SSTAMP(12);
  } else {
    for (int b = iot; b < B; b += NTIO)
      xoff_s[b] = (a.hx_on ? (long long)a.hx[b] : a.xidx ? (long long)a.xidx[b] : (long long)b) * (long long)P;
    bar_sync(4, NTIO);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (STAGE) cp_async_arrive_noinc(codes_landed);  // this thread's share of the gather (and nothing else: the image rows follow)
  if (tid == 0) SSTAMP(1);
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) SSTAMP(2);
  // From here the roles run free: the issuer and the I/O warps start fetching at once, the workers write the codes to
  // tensor memory (the issuer cannot start before every worker has handed over tile 0).

  if (warp >= WARP_LOAD) {
    // ===== I/O warps: every global access of the image rows, so that the workers only ever touch shared memory and
    // TMEM.  Per tile j: wait until the workers have staged the finished tile, store it with coalesced 128-bit
    // streaming stores, and refill the stage with the rows of tile j+NSX by cp.async (each thread overwrites exactly
    // the elements it has just read). =====
    float* dstg = a.out != nullptr ? a.out : a.delta;
    // (one loop, one call site of the row request: rounds -NSX .. -1 only request the first NSX tiles)
#pragma unroll 1
    for (int j = -NSX; j < my_tiles; ++j) {
      if (j >= 0) {
        const int p0 = (blockIdx.x + j * gridDim.x) * TP;
        const int sx = j % NSX;
        const float* xt = xs + sx * xstage;
        mbar_wait(out_ready + sx, (j / NSX) & 1);
        if (warp == WARP_LOAD && j == 0) SSTAMP(7);
        if (warp == WARP_LOAD && j == my_tiles - 1) SSTAMP(9);
#pragma unroll 4
        for (int e = iot; e < B * Q4; e += NTIO) {
          const int b = e / Q4, col = (e - b * Q4) * 4;
          if (p0 + col < P) st_stream4(dstg + (size_t)b * P + p0 + col, *reinterpret_cast<const float4*>(xt + b * XP + col));
        }
        if (warp == WARP_LOAD && j == 0) SSTAMP(8);
      }
      load_x(j + NSX);
    }
    if (warp == WARP_LOAD) SSTAMP(10);
  } else if (warp == WARP_MMA) {
    // ===== issuer: 3 x ksteps MMAs per tile into the accumulator buffer (it & 1) =====
    const uint32_t idesc = make_idesc_tf32(128, TP, false, false);
    const uint32_t a_hi = tmem_base + (uint32_t)(2 * TP), a_lo = a_hi + (uint32_t)a.Kp8;  // codes: columns after the accumulators
    const uint64_t dhi0 = make_desc(smem_u32(Dimg), a.Sd, 128);
    const uint64_t dlo_off = (uint64_t)((a.dimg * 4) >> 4), buf_off = 2 * dlo_off;
    const uint64_t bstep = (uint64_t)((2 * a.Sd) >> 4);
    const int ksteps = a.Kp8 / 8;
    const bool leader = elect_one();
    // (A raw stage is known to be free when its next tile is requested: the copy for tile it is issued right after the
    // hand-off of tile it-NS, i.e. after every worker has finished splitting it.)
    if (leader)
      for (int it = ragged ? 0 : 1; it < NS && it < my_tiles; ++it) load_D(it);
    for (int it = 0; it < my_tiles; ++it) {
      // Workers staged tile `it`.  An mbarrier per image buffer, not a named barrier: with double-buffered images a
      // fast warp reaches the hand-off of tile it+1 before a slow warp has arrived for tile it, and a second
      // bar.arrive of the same warp would complete the named barrier early.
      mbar_wait(staged + (it & 1), (it >> 1) & 1);
      if (it == 0) mbar_wait(codes_ready, 0);  // the workers have written the batch codes to tensor memory
      tc_fence_after();
      if (it == 0) SSTAMP(4);
      if (leader) {
        if (it + NS < my_tiles) load_D(it + NS);
        const uint32_t acc = tmem_base + (uint32_t)((it & 1) * TP);
        const uint64_t dhi = dhi0 + (uint64_t)(it & 1) * buf_off, dlo = dhi + dlo_off;
        {  // lo*hi
          uint32_t at = a_lo;
          uint64_t bd = dhi;
          for (int ks = 0; ks < ksteps; ++ks, at += 8, bd += bstep) mma_tf32_ts(acc, at, bd, idesc, ks ? 1u : 0u);
        }
        {  // hi*lo
          uint32_t at = a_hi;
          uint64_t bd = dlo;
          for (int ks = 0; ks < ksteps; ++ks, at += 8, bd += bstep) mma_tf32_ts(acc, at, bd, idesc, 1u);
        }
        {  // hi*hi
          uint32_t at = a_hi;
          uint64_t bd = dhi;
          for (int ks = 0; ks < ksteps; ++ks, at += 8, bd += bstep) mma_tf32_ts(acc, at, bd, idesc, 1u);
        }
        mma_commit(mma_done + (it & 1));
      }
      __syncwarp();
    }
  } else {
    // ===== workers =====
    const int quad = warp & 3, cg = warp >> 2;
    const int nitems = TP * (K / a.vk);  // vector items of the raw tile
    const int kv = K / a.vk;

    TileChan tc;
    tile_chan_init(tc);
    TIM_DECL;
    // batch codes (hi / lo) -> tensor memory, once per CTA: thread <-> image b = 32 quad + lane (its TMEM lane), the four
    // warps of a quadrant share the 8-atom chunks
    {
      const int nchunks = a.Kp8 / 8;
      const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(2 * TP);
      if constexpr (STAGE) {
        mbar_wait(codes_landed, 0);  // every warp's share of the gather has landed
        if (a.codes_out != nullptr && blockIdx.x == gridDim.x - 1) {
          // the last CTA (never one with more tiles than the others) leaves the gathered rows behind as a contiguous
          // [B, K] block: the backward kernel of the step fetches it with one bulk copy instead of a gather of its own
          for (int r = warp; r < B; r += NW)
            for (int k = lane; k < K; k += 32) a.codes_out[r * K + k] = cstage[r * a.cpitch + k];
        }
        const int b = quad * 32 + lane;
        const float* crow = cstage + min(b, B - 1) * a.cpitch;
        if (a.cw == 4) synth_codes_to_tmem<4>(crow, b < B, K, a.Kp8, cg, lane_base);
        else if (a.cw == 2) synth_codes_to_tmem<2>(crow, b < B, K, a.Kp8, cg, lane_base);
        else synth_codes_to_tmem<1>(crow, b < B, K, a.Kp8, cg, lane_base);
        tmem_st_wait();
        bar_sync(2, NT);  // every worker has read its rows: the second image buffer is free
        zero_padding(Dimg + 2 * a.dimg, tid, NT);
      } else {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = cg + 4 * ci;
          if (c < nchunks) {  // warp-uniform
            float hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_tf32(vv[ci][i], hi[i], lo[i]);
            tmem_st8(lane_base + (uint32_t)(8 * c), hi);
            tmem_st8(lane_base + (uint32_t)(a.Kp8 + 8 * c), lo);
          }
        }
        if (nchunks > 16) {
          // more than 128 atoms (up to 224: the codes take 2 Kp8 of the 512 TMEM columns): the chunks beyond the 16 that
          // were loaded at kernel entry follow in a second round -- one more cold miss, at large K only
          const int b = quad * 32 + lane;
          const int bi = min(b, B - 1);
          const long long row = a.hv_on ? (long long)a.hv[bi] : (a.vidx ? (long long)a.vidx[bi] : (long long)bi);
          const float* vrow = a.v + row * K;
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) load_codes8(vrow, 8 * (16 + cg + 4 * ci), K, b < B, a.cw == 4, vv[ci]);
          if (a.codes_out != nullptr && blockIdx.x == gridDim.x - 1 && b < B) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) store_codes8(a.codes_out + (size_t)b * K, 8 * (16 + cg + 4 * ci), K, a.cw == 4, vv[ci]);
          }
#pragma unroll
          for (int ci = 0; ci < 4; ++ci) {
            const int c = 16 + cg + 4 * ci;
            if (c < nchunks) {  // warp-uniform
              float hi[8], lo[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) split_tf32(vv[ci][i], hi[i], lo[i]);
              tmem_st8(lane_base + (uint32_t)(8 * c), hi);
              tmem_st8(lane_base + (uint32_t)(a.Kp8 + 8 * c), lo);
            }
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(codes_ready);
    }
    if (warp == 0) SSTAMP(3);

    auto epilogue = [&](int j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int p0 = tile * TP;
      const int sx = j % NSX;
      float* xt = xs + sx * xstage;
      mbar_wait(mma_done + (j & 1), (j >> 1) & 1);  // accumulator of tile j complete
      tc_fence_after();
      TIM(9);
      if (warp == 0 && j == 0) SSTAMP(5);
      mbar_wait(full_x + sx, (j / NSX) & 1);        // rows landed / the I/O warps are done with this stage
      TIM(3);
      if (warp == 0 && j == 0) SSTAMP(6);
      // phase 1: thread <-> image row b; 16 accumulator columns per warp
      if (cg < NCG) {
        const int b = quad * 32 + lane;
        float r[16];
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((j & 1) * TP + cg * 16), r);
        TIM(4);
        if (b < B) {
          if (a.cc.use) tile_chan_update(tc, a.cc, p0);
          float* xrow = xt + b * XP + cg * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float d[4] = {r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]};
            if (!TRAIN && (a.flags & ADIL_SYNTH_CLAMP_DELTA)) {
#pragma unroll
              for (int i = 0; i < 4; ++i) d[i] = fminf(fmaxf(d[i], -a.eps), a.eps);
            }
            float o[4] = {d[0], d[1], d[2], d[3]};
            if (TRAIN || a.out != nullptr) {
              if (!TRAIN && a.delta != nullptr) {  // secondary output: written straight from the accumulator row
                const int p = p0 + cg * 16 + 4 * c;
                if (p < P) st_stream4(a.delta + (size_t)b * P + p, make_float4(d[0], d[1], d[2], d[3]));
              }
              if (TRAIN || need_x) {
                const float4 xv = *reinterpret_cast<const float4*>(xrow + 4 * c);
                o[0] = __fadd_rn(xv.x, d[0]); o[1] = __fadd_rn(xv.y, d[1]);
                o[2] = __fadd_rn(xv.z, d[2]); o[3] = __fadd_rn(xv.w, d[3]);
              }
              if (!TRAIN && (a.flags & ADIL_SYNTH_CLAMP01)) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = fminf(fmaxf(o[i], 0.0f), 1.0f);
              }
              if (a.cc.use) {
                // Normalize (demo_dL_attack.py:22-25) as one FFMA: (o - mean) / std == o * rstd - mean * rstd up to
                // 1.5 ulp of the result (|result| < 3: < 4e-7)
                const bool hi_c = (cg * 16 + 4 * c) >= tc.bnd;
                const float rstd = hi_c ? tc.rstd1 : tc.rstd0, nmr = hi_c ? tc.nmr1 : tc.nmr0;
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = __fmaf_rn(o[i], rstd, nmr);
              }
            }
            *reinterpret_cast<float4*>(xrow + 4 * c) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        tc_fence_before();
      }
      TIM(5);
      __syncwarp();
      if (lane == 0) mbar_arrive(out_ready + sx);  // this warp's part of the finished tile is staged
      TIM(8);
    };

    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % NS;
      if (it >= 2) {
        mbar_wait(mma_done + (it & 1), ((it - 2) >> 1) & 1);  // MMAs(it-2) retired: this image buffer is free again
        tc_fence_after();
      }
      TIM(0);
      mbar_wait(full_raw + s, (it / NS) & 1);
      TIM(1);
      const float* rt = raw + s * a.raw_floats;
      float* Dhi = Dimg + (it & 1) * 2 * a.dimg;
      float* Dlo = Dhi + a.dimg;
      if (a.vk == 4) {
        for (int e0 = tid; e0 < nitems; e0 += 4 * NT) {  // four items per thread in flight
          float4 rawv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * NT;
            rawv[u] = e < nitems ? *reinterpret_cast<const float4*>(rt + 4 * e) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * NT;
            if (e < nitems) {
              const int p = div_magic_dev(e, a.kdiv), k = (e - p * kv) * 4;
              float4 hi, lo;
              split_tf32(rawv[u].x, hi.x, lo.x); split_tf32(rawv[u].y, hi.y, lo.y);
              split_tf32(rawv[u].z, hi.z, lo.z); split_tf32(rawv[u].w, hi.w, lo.w);
              const int o = (k >> 2) * (a.Sd >> 2) + (p >> 3) * 32 + (p & 7) * 4;
              *reinterpret_cast<float4*>(Dhi + o) = hi;
              *reinterpret_cast<float4*>(Dlo + o) = lo;
            }
          }
        }
      } else if (a.vk == 2) {
        for (int e0 = tid; e0 < nitems; e0 += 4 * NT) {
          float2 rawv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * NT;
            rawv[u] = e < nitems ? *reinterpret_cast<const float2*>(rt + 2 * e) : make_float2(0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * NT;
            if (e < nitems) {
              const int p = div_magic_dev(e, a.kdiv), k = (e - p * kv) * 2;
              float2 hi, lo;
              split_tf32(rawv[u].x, hi.x, lo.x); split_tf32(rawv[u].y, hi.y, lo.y);
              const int o = (k >> 2) * (a.Sd >> 2) + (p >> 3) * 32 + (p & 7) * 4 + (k & 3);
              *reinterpret_cast<float2*>(Dhi + o) = hi;
              *reinterpret_cast<float2*>(Dlo + o) = lo;
            }
          }
        }
      } else {
        for (int e = tid; e < nitems; e += NT) {
          const int p = div_magic_dev(e, a.kdiv), k = e - p * kv;
          float hi, lo;
          split_tf32(rt[e], hi, lo);
          const int o = (k >> 2) * (a.Sd >> 2) + (p >> 3) * 32 + (p & 7) * 4 + (k & 3);
          Dhi[o] = hi;
          Dlo[o] = lo;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();                      // the ragged last split iteration of every lane is over
      if (lane == 0) mbar_arrive(staged + (it & 1));  // hand the tile (and the consumed raw stage) to the issuer
      TIM(2);
      if (it > 0) epilogue(it - 1);      // overlaps the MMAs being issued
    }
    if (my_tiles > 0) epilogue(my_tiles - 1);
    TIM_FLUSH(my_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, a.tmem_cols);
  if (tid == 0) SSTAMP(11);
}

// =========================================================================================================
// backward contractions (+ optional fused AdamW / clamp), bf16x3:
//   dD^T[k, p] = sum_b v[b,k] gx[b,p]   M = 128 atom lanes, N = TP pixels, contraction over images
//                A = batch codes (three bf16 terms) in TENSOR MEMORY: lane = atom, column = image pair (loop
//                    invariant: written once per CTA with tcgen05.st; saves 2 KB of shared-memory reads per MMA)
//                B = gradient images  r = image, c = pixel  MN-major
//   dv[b, k]   = sum_p gx[b,p] D[p,k]   M = 128 image lanes, N = atoms, contraction over pixels (accumulated in TMEM
//                over all tiles of the CTA)
//                A = gradient images  (the same bytes, read K-major)    B = D tile images  r = pixel, c = atom  MN-major
// gx = g / std[c] is folded into the operands' consumers: dD accumulators are multiplied by 1/std on their way out and
// the dictionary tile of the dv contraction is divided by std while it is split.
// The raw D (and m, s) tiles land by TMA; the dD tile is transposed through shared memory into the flat [p][k]
// layout of the dictionary, where AdamW + clamp run as one 128-bit vectorised pass that stores D, m, s coalesced.
// The gradient and dictionary images are double-buffered: the workers stage tile i+1 while the MMAs of tile i run.
// =========================================================================================================
struct GradArgs {
  float* dD2;
  float* D2w;
  float* m;
  float* s;
  float* partial;
  const float* g;
  const float* D2;
  const float* v;
  const int64_t* vidx;
  int hv_on;          // the batch indices travel in the kernel parameters (host index array hv[], at the end)
  int B, P, K;
  int Bp;             // contraction length of dD: round_up(B, 16)
  int Kp;             // N of the dv MMA: round_up(K, 16)
  int Sg, Sd;         // byte strides between column groups of the gradient (= code) / dictionary images
  int dimg, gimg;     // bf16 elements per image term
  int raw_floats;     // floats per raw stage
  int nraw;           // 1: the stages receive the D rows of each tile by TMA; 0: no loads (dD only: staging buffers)
  int vk;             // vector width of the dictionary split
  unsigned kdiv;      // ceil(2^32 / (K / vk))
  unsigned gdiv;      // ceil(2^32 / ceil(K / 8)): 8-atom groups of the G_SCALED dictionary split
  uint32_t tmem_cols;
  int want_dD, want_dv, atoms_mode;
  int ldk;            // row pitch (floats) of D2 / m / s / dD2 / v in global memory: K, or the full atom count when this
                      // launch handles a window of K columns of a wider dictionary (STRIDED: more than 128 atoms)
  unsigned k4div;     // ceil(2^32 / (K / 4)) (STRIDED: float4 index inside a dense tile -> row)
  int wsh;            // nwin >> 1
  int dv2;            // two dv accumulators, tiles alternate (a CTA of a window pair runs twice as many tiles: the fp32
                      // accumulation chains in tensor memory keep the length of the single-window kernels)
  int fullrow;        // STRIDED: the raw stage holds the FULL dictionary rows of a tile (ldk floats apart: one bulk copy per
                      // tile instead of one per row; this CTA's window starts `wof` floats into each row)
  int rpitch;         // floats between the rows of a tile in the raw stage: ldk (fullrow) or K
  unsigned rmask;     // all ones (fullrow) or 0: masks `wof` in stage offsets
  int rsv_[2];        // (keeps `hp` 16-byte aligned in the constant bank: the AdamW pass fetches it with one LDCU.128)
  int nwin;           // STRIDED: 1, or 2 = BOTH column windows in this launch: CTA c takes window c & 1 (columns
                      // [K (c & 1), K (c & 1) + K) of every array) and the pixel tiles (c >> 1) + j (gridDim.x >> 1)
  int accumulate;     // plain dD output: dD2 += tile (TMA reduce-add store) instead of dD2 = tile
  int dreg;           // bytes of the dictionary-image region; at kernel entry it stages the code rows of the batch
  int cw;             // floats per cp.async of the code-row gather: 4, 2 or 1 (alignment of v, ldk and K)
  int rpw;            // rows of the code gather per warp: ceil(B / 25) (warps 0..24 take part)
  int pfast;          // dictionary split: consecutive lanes take consecutive pixels (K % 8 != 0) instead of atom groups
  int codes_contig;   // the code rows are rows 0..B-1 of v, K floats apart, 16-byte aligned as a block: one bulk copy
  ChannelConsts cc;
  AdamwDev hp;
  int hv[128];
};

// FUSED (AdamW + clamp in the epilogue warps) is a template parameter: the two variants get their own register
// allocation -- at the 72-register cap of 896 threads the epilogue code of the one perturbed the other's spills.
// STRIDED: this launch handles a window of K columns of a dictionary whose rows are `ldk` floats apart (more than 128
// atoms: two column windows, two launches).  Tiles are dense [TP][K] in shared memory as always; only the global
// addresses change: the D rows of a tile arrive as one bulk copy PER ROW (issued by the 32 lanes of the loader warp),
// the moments / outputs are addressed by (row, column) instead of a flat offset.
template <int TP, bool FUSED, bool STRIDED, int NPF>
__global__ void __launch_bounds__(NTHREADS_GRAD, 1) grad_kernel(const GradArgs a) {
  constexpr int Q4 = TP / 4;                           // float4 per gradient row
  constexpr int GJ = (128 * Q4 + NT - 1) / NT;         // float4 per worker thread (B <= 128)
  constexpr int HALF = TP / 2;                         // accumulator columns per epilogue warp
  // NPF: float4 items per thread of the AdamW pass whose moments are prefetched into registers: 3 when the tile has at
  // most 3 * 320 of them (K <= 60 at TP = 64) -- ten warps then share the pass; else 4 items on the eight epilogue warps.
  // (Ten warps with 4 items, or any spill in this kernel, cost more than they save: a reload from local memory queues
  // behind the stores of the pass.)
  constexpr int NEP = (FUSED && NPF == 3) ? NEP_MAX : NE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // (contiguous, in this order: the loader initialises them one per lane)
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(smem_raw);  // [NS] D (, m, s) tile landed
  uint64_t* empty_raw = full_raw + NS;                         // [NS] stage consumed (workers: split, epilogue warps: AdamW)
  uint64_t* mma_done = empty_raw + NS;                         // [2] MMAs of the tiles using image buffer 0 / 1 retired
  uint64_t* staged = mma_done + 2;                             // [2] images of buffer 0 / 1 written (workers -> issuer)
  uint64_t* acc_empty = staged + 2;                            // [2] dD accumulator 0 / 1 read out (epilogue warps -> issuer)
  uint64_t* epi_done = acc_empty + 2;                          // [NS] output tile written into the stage (epilogue warps -> loader)
  uint64_t* dD_done = epi_done + NS;                           // [2] dD MMAs of the tiles using accumulator 0 / 1 retired
  uint64_t* codes_ready = dD_done + 2;                         // [1] batch codes written to tensor memory (epilogue warps)
  uint64_t* codes_landed = codes_ready + 1;                    // [1] code rows of the batch gathered into shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(codes_landed + 1);
  typedef unsigned short bf16_t;
  const int dbuf = 3 * a.dimg, gbuf = 3 * a.gimg + 1024;          // bf16 elements per buffer (gradient: + 2 KB pad)
  bf16_t* Di = reinterpret_cast<bf16_t*>(smem_raw + HDR_BYTES);   // [2 buffers][3 terms] dictionary images (a.dreg bytes)
  bf16_t* Gi = reinterpret_cast<bf16_t*>(smem_raw + HDR_BYTES + a.dreg);  // [2 buffers][3 terms (+ over-read pad)] gradient images
  float* dDs = reinterpret_cast<float*>(Gi + 2 * gbuf);           // [2][TP*K] dD tiles, flat like the dictionary (fused step)
  float* raw = dDs + (a.D2w != nullptr ? 2 * TP * a.K : 0);       // [NS][raw_floats]
  // [B][K] code rows of the batch, until they are in TMEM: in the dD tiles when they fit there (fused step, B <= 2 TP: first
  // written after the MMAs of tile 0, which need the codes), else in the dictionary-image region (whose first use then
  // waits for the codes)
  const bool codes_in_dDs = FUSED && a.B <= 2 * TP;
  float* cstage = codes_in_dDs ? dDs : reinterpret_cast<float*>(Di);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, P = a.P, B = a.B;
  const int ntiles = (P + TP - 1) / TP;
  // Two column windows in one launch (STRIDED, a.nwin == 2): neighbouring CTAs 2i, 2i+1 work on the SAME pixel tiles at
  // the same time, one window each -- the gradient rows are fetched from HBM once (the second CTA hits L2), both halves
  // of every dictionary / moment row are touched together (full DRAM pages, the shared sectors at the window boundary
  // merge in L2), and the step costs one launch instead of two.
  // (cta_x, cta_n, wof are recomputed from the special registers / the constant bank where they are used: held in
  // registers across the kernel they made the 4-item AdamW pass spill)
#define cta_x ((int)blockIdx.x >> a.wsh)
#define cta_n ((int)gridDim.x >> a.wsh)
#define wof ((unsigned)(((int)blockIdx.x & a.wsh) * a.K))
  // float offset of float4 `e4` of the dense [rows][K] tile inside a raw stage
  auto soff = [&](int e4) -> int {
    if constexpr (STRIDED) {
      const int r = div_magic_dev(e4, a.k4div);  // (shared with goff)
      return r * a.rpitch + (int)(wof & a.rmask) + 4 * (e4 - r * (K >> 2));
    } else {
      return 4 * e4;
    }
  };
  auto W = [&](auto* ptr) { if constexpr (STRIDED) return ptr + wof; else return ptr; };
  auto tile_p0 = [&](int it) -> int {
    if constexpr (STRIDED) return (cta_x + it * cta_n) * TP;
    else return (blockIdx.x + it * gridDim.x) * TP;
  };
  const int my_tiles = [&]() -> int {
    if constexpr (STRIDED) return (ntiles - cta_x + cta_n - 1) / cta_n;
    else return (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  }();
  constexpr bool fused = FUSED;
  // G_SCALED: 1/std is folded into the gradient images (gx = g * (1/std): one multiply per gradient element; both
  // contractions want gx) and the dictionary tile is split as it is, eight atoms per item with 16-byte stores; otherwise
  // the dictionary split divides by std and the epilogue warps scale the dD accumulator.  Measured with the two variants
  // compiled separately: plain contractions 59.4 -> 56.3 us, fused step 67.6 -> 63.5 us.  (The old path stays for A/B runs.)
#ifndef ADIL_G_SCALED_FUSED
#define ADIL_G_SCALED_FUSED 1
#endif
#ifndef ADIL_G_SCALED_PLAIN
#define ADIL_G_SCALED_PLAIN 1
#endif
  constexpr bool G_SCALED = FUSED ? (ADIL_G_SCALED_FUSED != 0) : (ADIL_G_SCALED_PLAIN != 0);
  const int tile_elems = TP * K;
  const uint32_t ne_tmem = 2u * (uint32_t)((K + 31) / 32);  // epilogue warps whose TMEM quadrant holds atoms
  STAMP(0);
  // global element offset of float4 `e4` of the dense tile that starts at pixel p0
  auto goff = [&](int p0, int e4) -> size_t {
    if constexpr (STRIDED) {
      const int r = div_magic_dev(e4, a.k4div);
      return (size_t)(p0 + r) * (size_t)a.ldk + (size_t)(wof + 4u * (unsigned)(e4 - r * (K >> 2)));  // (+ this CTA's window)
    } else {
      return (size_t)p0 * K + 4 * (size_t)e4;
    }
  };
  // D tile `it` -> raw stage it % NS: a contiguous run, one TMA bulk copy (the stages are never zero-filled: a ragged
  // last tile is handled by the consumers).
  auto load_raw = [&](int it, bool leader) {  // (called by the whole loader warp)
    const int p0 = tile_p0(it);
    const int rows = min(TP, P - p0);
    const int s = it % NS;
    const uint32_t bytes = (uint32_t)(rows * K * 4);
    float* dst = raw + s * a.raw_floats;
    if constexpr (STRIDED) {
      if (a.fullrow) {  // whole rows, both windows: a contiguous run
        if (leader) {
          mbar_expect_tx(full_raw + s, (uint32_t)(rows * a.ldk * 4));
          bulk_g2s(dst, a.D2 + (size_t)p0 * (size_t)a.ldk, (uint32_t)(rows * a.ldk * 4), full_raw + s);
        }
      } else {
        if (leader) mbar_expect_tx(full_raw + s, bytes);
        __syncwarp();
        for (int r = lane; r < rows; r += 32)
          bulk_g2s(dst + r * K, W(a.D2) + (size_t)(p0 + r) * (size_t)a.ldk, (uint32_t)(K * 4), full_raw + s);
      }
    } else if (leader) {
      mbar_expect_tx(full_raw + s, bytes);
      bulk_g2s(dst, a.D2 + (size_t)p0 * K, bytes, full_raw + s);
    }
    __syncwarp();
  };
  // workers: fixed per-thread share of the gradient tile: float4 e = tid + j*NT -> image b = e / Q4, 4-pixel column q.
  int gsrc[GJ], gdst[GJ];  // gsrc: 4q, the pixel column (-1: none); gdst: bf16 offset inside an image
  const float* grow[GJ];   // &g[b, 4q]
  float4 greg[GJ];
  auto prefetch = [&](int it) {
    const int p0 = tile_p0(it);
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      greg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gsrc[j] >= 0 && p0 + gsrc[j] < P) greg[j] = ld_stream4(grow[j] + p0);
    }
  };

  // ---- Kernel entry.  Everything the first tile needs is ONE cold miss away and depends on nothing but the kernel
  // parameters, so every role puts its requests out before any set-up work: the workers the gradient rows of tile 0
  // (registers), the epilogue warps the code rows of the batch (cp.async gather into shared memory: B rows of K floats
  // in ceil(K / 32 cw) requests per row -- the former register gather was 512 requests of 128 bytes per CTA and took
  // 3 us just to issue), the loader the first NS dictionary tiles (TMA).  Barrier init, the TMEM allocation and the zero
  // fill of the operand images run underneath. ----
  // Code rows of the batch -> shared memory.  Contiguous rows (no index, 16-byte-aligned block: the caller passes the
  // [B, K] block that adil_synth left behind) arrive as ONE bulk copy issued by the loader; else every warp but the
  // loader's gathers rows [w rpw, (w + 1) rpw) by cp.async: lane l holds the address of row w rpw + (l & 7) (one line
  // of the parameters or of the index array per warp).  The load/store unit accepts such a sub-line cp.async only every
  // ~20 ns, SM-wide (measured: B = 100 rows take 2 us however they are spread over the warps), so the gather is the
  // slow path.
  auto gather_share = [&]() {
    const int r0 = warp * a.rpw;
    if (a.want_dD && !a.codes_contig && r0 < B) {
      const int rb = min(r0 + (lane & 7), B - 1);
      const long long row = a.hv_on ? (long long)a.hv[rb] : (a.vidx ? (long long)a.vidx[rb] : (long long)rb);
      gather_rows_cw<8>(cstage + r0 * K, K, W(a.v) + row * (long long)a.ldk, min(a.rpw, B - r0), K, a.cw, lane);
    }
  };
  if (warp < NW) {
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      const int e = tid + j * NT;
      const int b = e / Q4, q = e - b * Q4;
      gsrc[j] = (b < B) ? 4 * q : -1;
      grow[j] = a.g + (size_t)min(b, B - 1) * P + 4 * q;
      gdst[j] = (q >> 1) * (a.Sg >> 1) + (b >> 3) * 64 + (b & 7) * 8 + (q & 1) * 4;
    }
    if (my_tiles > 0) prefetch(0);
  }
  if (warp != WARP_LOAD_G) gather_share();  // (one call site: the workers' share queues behind their gradient rows)
  if (warp < NW) {
    STAMP(8);
    // zero the operand images once: contraction padding (images B..Bp, atoms K..Kp) must be zero, over-read regions
    // finite.  (No proxy fence here: it is a MEMBAR.ALL.CTA and would wait for the loads in flight; every thread fences
    // before its first hand-off to the tensor core, which covers these stores too.)  The dictionary-image region is
    // left alone while it stages the code rows.
    uint4* z = reinterpret_cast<uint4*>(codes_in_dDs || !a.want_dD ? reinterpret_cast<bf16_t*>(Di) : Gi);
    const int nz = ((codes_in_dDs || !a.want_dD ? a.dreg : 0) >> 4) + ((2 * gbuf) >> 3);
    for (int e = tid; e < nz; e += NT) z[e] = make_uint4(0u, 0u, 0u, 0u);
    STAMP(12);
  } else if (warp == WARP_EPI + NE) {
    tmem_alloc(tmem_slot, a.tmem_cols);
    WSTAMP(1);
  } else if (warp == WARP_LOAD_G) {
    // the loader initialises the mbarriers, one per lane (they are contiguous from full_raw on), and requests the first
    // NS dictionary tiles
    if (lane < 3 * NS + 10) {
      const uint32_t c_empty = (uint32_t)((a.want_dv ? NW : 0) + (fused ? NEP : 0) + ((a.want_dv || fused) ? 0 : 1));
      uint32_t cnt = 1;                                            // full_raw, mma_done, dD_done
      if (lane >= NS && lane < 2 * NS) cnt = c_empty;              // empty_raw
      else if (lane >= 2 * NS + 2 && lane < 2 * NS + 4) cnt = NW;  // staged
      else if (lane >= 2 * NS + 4 && lane < 3 * NS + 6) cnt = ne_tmem;  // acc_empty, epi_done
      else if (lane == 3 * NS + 8) cnt = NE;                       // codes_ready
      else if (lane == 3 * NS + 9) cnt = a.codes_contig ? 1 : NTHREADS_GRAD - 32;  // codes_landed: the bulk copy / every gathering thread
      mbar_init(full_raw + lane, cnt);
      fence_mbar_init();
    }
    __syncwarp();
    const bool leader = elect_one();
    if (leader && a.want_dD && a.codes_contig) {
      mbar_expect_tx(codes_landed, (uint32_t)(B * K * 4));
      bulk_g2s(cstage, a.v, (uint32_t)(B * K * 4), codes_landed);
    }
    // (only the first dictionary tile now: the others would queue megabytes ahead of what tile 0 is waiting for)
    if (a.nraw > 0 && my_tiles > 0) load_raw(0, leader);
    WSTAMP(2);
  }
  // set-up barrier: the operand images are zeroed, the mbarriers initialised, tensor memory allocated
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.want_dD && !a.codes_contig && warp != WARP_LOAD_G) cp_async_arrive_noinc(codes_landed);  // this thread's share of the gather
  STAMP(10);
  const uint32_t tmem_base = *tmem_slot;
  // tensor memory: [0, 2 TP) two dD^T accumulators | [2 TP, 2 TP + Kp) dv accumulator | then the codes, 3 x Bp/2 columns
  const uint32_t acc_dv = tmem_base + (uint32_t)(2 * TP);
  const uint32_t codes = [&]() -> uint32_t {
    if constexpr (STRIDED) return acc_dv + (uint32_t)(a.dv2 ? 2 * a.Kp : a.Kp);
    else return acc_dv + (uint32_t)a.Kp;
  }();
  STAMP(1);
  // (the setmaxnreg instructions open the role branches below)

  if (warp == WARP_LOAD_G) {
    reg_dealloc<24>();
    // ===== loader: the remaining D tiles, each as soon as its stage has been recycled (by the workers' dictionary split
    // and, in the fused step, by the AdamW pass of the epilogue warps).  Plain dD output: the epilogue warps write
    // the dD tile into the stage and it leaves here as one TMA bulk store (full-line writes whatever the row pitch);
    // the stage is reused when the copy engine has read it. =====
    const bool leader = elect_one();
    const bool dD_out = a.want_dD && !fused;
    if (a.nraw > 0)
      for (int it = 1; it < NS && it < my_tiles; ++it) load_raw(it, leader);  // (tile 0 was requested at kernel entry)
    for (int it = NS; it < my_tiles + NS; ++it) {
      const int jt = it - NS;  // the tile whose stage is recycled now
      {
        const int sj = jt % NS;
        if (dD_out) {
          mbar_wait(epi_done + sj, (jt / NS) & 1);
          if constexpr (STRIDED) {
            // (column window: the epilogue warps stored the tile themselves; the stage is free)
            if (leader && it < my_tiles && a.nraw == 0) mbar_arrive(full_raw + sj);
          } else if (leader) {
            const int p0 = (blockIdx.x + jt * gridDim.x) * TP;
            if (a.accumulate) bulk_red_add_s2g(a.dD2 + (size_t)p0 * K, raw + sj * a.raw_floats, (uint32_t)(min(TP, P - p0) * K * 4));
            else bulk_s2g(a.dD2 + (size_t)p0 * K, raw + sj * a.raw_floats, (uint32_t)(min(TP, P - p0) * K * 4));
            bulk_commit();
            if (it < my_tiles) {  // the stage is about to be refilled
              bulk_wait_read0();
              if (a.nraw == 0) mbar_arrive(full_raw + sj);  // (no loads: hand the staging buffer back directly)
            }
          }
          __syncwarp();
        }
        if ((a.want_dv || fused) && it < my_tiles) mbar_wait(empty_raw + sj, (jt / NS) & 1);  // D rows consumed
      }
      if (it < my_tiles && a.nraw > 0) load_raw(it, leader);
    }
    if (!STRIDED && leader && dD_out) bulk_wait0();  // every output tile has been written before the CTA retires
    __syncwarp();
  } else if (warp == WARP_MMA_G) {
    // ===== issuer =====
    reg_dealloc<24>();
    const uint32_t idesc_dD = make_idesc_bf16(128, TP, false, true);
    const uint32_t idesc_dv = make_idesc_bf16(128, a.Kp, false, true);
    const uint32_t gb = smem_u32(Gi), db = smem_u32(Di);
    // term pairs (gradient term, code / dictionary term), smallest contributions first: (2,0)(0,2)(1,1)(1,0)(0,1)(0,0)
    // dD^T: A = codes in TMEM (term tv at column offset tv * Bp/2, 8 columns per 16-image k-step)
    //       B = gradient image [N = pixel, K = image] MN-major: LBO = 128 (8-image groups), SBO = Sg (8-pixel groups)
    // dv:   A = gradient image [M = image, K = pixel] K-major: LBO = Sg (8-pixel chunks), SBO = 128 (8-image groups)
    //       B = D image [N = atom, K = pixel] MN-major: LBO = 128 (8-pixel groups), SBO = Sd (8-atom groups)
    const uint64_t dD_b0 = make_desc(gb, 128, a.Sg);
    const uint64_t dv_a0 = make_desc(gb, a.Sg, 128), dv_b0 = make_desc(db, 128, a.Sd);
    const uint64_t gs16 = (uint64_t)((2u * a.gimg) >> 4), ds16 = (uint64_t)((2u * a.dimg) >> 4);  // term strides
    const uint64_t gb16 = (uint64_t)((2u * gbuf) >> 4), db16 = (uint64_t)((2u * dbuf) >> 4);      // buffer strides
    const uint64_t astep = (uint64_t)((2 * a.Sg) >> 4);
    const int ksteps_dD = a.Bp / 16;
    const uint32_t cterm = (uint32_t)(a.Bp / 2);
    const bool leader = elect_one();
    for (int it = 0; it < my_tiles; ++it) {
      const int buf = it & 1;
      if (a.want_dD && it >= 2) mbar_wait(acc_empty + buf, ((it - 2) >> 1) & 1);  // accumulator of tile it-2 read out
      CHAIN(0, it);
      mbar_wait(staged + buf, (it >> 1) & 1);  // workers staged tile `it` (mbarrier: fast warps run one tile ahead)
      tc_fence_after();
      if (it == 0) WSTAMP(6);
      CHAIN(1, it);
      auto issue_dD = [&]() {
        const uint32_t acc = tmem_base + (uint32_t)(buf * TP);
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          constexpr int tg[6] = {2, 0, 1, 1, 0, 0}, tv[6] = {0, 2, 1, 0, 1, 0};
          const uint32_t at = codes + (uint32_t)tv[t] * cterm;
          const uint64_t bd = dD_b0 + (uint64_t)buf * gb16 + (uint64_t)tg[t] * gs16;
          // unrolled over the (at most 8) 16-image steps: as a rolled loop every MMA cost ~11 dependent uniform-datapath
          // instructions (descriptor arithmetic, loop control) on a warp that shares its scheduler with six busy ones,
          // and the 42 dD MMAs of a tile were issue-bound (tensor pipe 36 % active)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            if (ks < ksteps_dD) mma_bf16_ts(acc, at + 8u * ks, bd + 16u * ks, idesc_dD, (t | ks) ? 1u : 0u);
        }
      };
      auto issue_dv = [&]() {
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          constexpr int tg[6] = {2, 0, 1, 1, 0, 0}, td[6] = {0, 2, 1, 0, 1, 0};
          uint64_t ad = dv_a0 + (uint64_t)buf * gb16 + (uint64_t)tg[t] * gs16;
          uint64_t bd = dv_b0 + (uint64_t)buf * db16 + (uint64_t)td[t] * ds16;
#pragma unroll
          for (int ks = 0; ks < TP / 16; ++ks, ad += astep, bd += 16)
            if constexpr (STRIDED) {
              const bool two = a.dv2 != 0;
              mma_bf16(acc_dv + (uint32_t)((two && (it & 1)) ? a.Kp : 0), ad, bd, idesc_dv, ((two ? (it >> 1) : it) | t | ks) ? 1u : 0u);
            } else {
              mma_bf16(acc_dv, ad, bd, idesc_dv, (it | t | ks) ? 1u : 0u);
            }
        }
      };
      // Tile 0: the dv MMAs need no codes, so they go first and the code conversion of the epilogue warps hides behind
      // them; from tile 1 on the dD MMAs go first and are committed on their own, so that the epilogue warps start on
      // the accumulator while the dv MMAs of the tile run.
      if (it == 0) {
        if (leader && a.want_dv) issue_dv();
        __syncwarp();
        if (a.want_dD) {
          mbar_wait(codes_ready, 0);  // the epilogue warps have written the batch codes to TMEM
          tc_fence_after();
          if (leader) { issue_dD(); mma_commit(dD_done + buf); }
        }
        if (leader) mma_commit(mma_done + buf);
        WSTAMP(7);
      } else if (leader) {
        if (a.want_dD) { issue_dD(); mma_commit(dD_done + buf); }
        if (a.want_dv) issue_dv();
        mma_commit(mma_done + buf);
      }
      __syncwarp();
      if (it == 0) WSTAMP(8);
      CHAIN(2, it);
    }
  } else if (warp >= WARP_EPI + NEP || (warp >= WARP_EPI + NE && !a.want_dD)) {
    // (warps 24, 25 outside the fused step: idle filler so that the issuer and the loader land on schedulers 2 and 3)
    reg_dealloc<24>();
  } else if (warp >= WARP_EPI) {
    reg_alloc<96>();
    const bool helper = warp >= WARP_EPI + NE;  // warps 24, 25: the AdamW pass only (no TMEM quadrant of their own)
    // ===== epilogue warps (8, two per scheduler).  Phase A: the warps whose TMEM quadrant holds atoms (quadrant q,
    // pixel half h: atoms [32q, 32q+32) of pixels [h TP/2, (h+1) TP/2)) read the dD^T accumulator (lane = atom,
    // column = pixel), release it to the tensor core, multiply by 1/std and write it flat -- [pixel][atom], like the
    // dictionary -- into shared memory: per pixel a warp writes 32 consecutive floats, conflict-free.  Fused step:
    // into a double-buffered gradient tile; then (phase B) all eight warps run AdamW + clamp as one 128-bit pass over
    // the TMA-landed D / m / s rows IN PLACE.  Plain dD output: straight into the stage.  Either way the finished
    // stage leaves through the loader warp as TMA bulk stores. =====
    const int quad = warp & 3, half = (warp - WARP_EPI) >> 2;
    const bool has_atoms = quad * 32 < K && !helper;
    if (a.want_dD && !helper) {
      // batch codes (three bf16 terms) -> tensor memory, once per CTA: lane = atom, a column holds the image pair
      // (2c, 2c+1).  The rows were gathered into shared memory at kernel entry; thread <-> atom m = 32 quad + lane reads
      // its column (consecutive lanes, consecutive words: conflict-free); the two warps of a quadrant share the
      // 16-image chunks.  A quadrant whose 32 atoms are all padding writes zeros.
      mbar_wait(codes_landed, 0);  // every warp's share of the gather has landed
      if (warp == WARP_EPI) { WSTAMP(3); WSTAMP(4); }
      const int m = quad * 32 + lane;
      const bool m_ok = m < K;
      const int nchunks = a.Bp / 16;  // 8-column chunks per term
      const uint32_t lane_base = codes + ((uint32_t)(quad * 32) << 16);
      const float* ccol = cstage + min(m, K - 1);
#pragma unroll 1
      for (int c = half; c < nchunks; c += 2) {  // (rolled: unrolled, the cold instruction fetch of 16 KB cost 5 us)
        uint32_t t0[8], t1[8], t2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int b0 = 16 * c + 2 * i;
          const float x0 = (m_ok && b0 < B) ? ccol[b0 * K] : 0.0f;
          const float x1 = (m_ok && b0 + 1 < B) ? ccol[(b0 + 1) * K] : 0.0f;
          uint32_t a0, a1, a2, b0w, b1w, b2w;
          split_bf16x3(x0, a0, a1, a2);
          split_bf16x3(x1, b0w, b1w, b2w);
          t0[i] = pack_hi16(a0, b0w);
          t1[i] = pack_hi16(a1, b1w);
          t2[i] = pack_hi16(a2, b2w);
        }
        tmem_st8u(lane_base + (uint32_t)(8 * c), t0);
        tmem_st8u(lane_base + (uint32_t)(a.Bp / 2 + 8 * c), t1);
        tmem_st8u(lane_base + (uint32_t)(a.Bp + 8 * c), t2);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(codes_ready);
      if (warp == WARP_EPI) WSTAMP(5);
    }
    if (a.want_dD && (fused || has_atoms)) {  // (plain dD output: a quadrant without atoms has nothing to do)
      const int k = quad * 32 + lane;
      const bool k_ok = k < K;
      const int etid = tid - WARP_EPI * 32;
      TileChan tc;
      tile_chan_init(tc);
      // AdamW moments: straight from global memory into registers, one tile ahead (they never need shared memory)
      float4 Mp[NPF], Sp[NPF];
      auto prefetch_ms = [&](int p0) {  // p0: first pixel of the tile
        const int n4 = (min(TP, P - p0) * K) >> 2;
#pragma unroll
        for (int u = 0; u < NPF; ++u) {
          const int e4 = etid + u * (NEP * 32);
          if (e4 < n4) {
            const size_t go = goff(p0, e4);
            Mp[u] = ld_global4(a.m + go);
            Sp[u] = ld_global4(a.s + go);
          }
        }
      };
      if constexpr (STRIDED) {
        if (fused && my_tiles > 0) prefetch_ms(cta_x * TP);
      } else {
        if (fused && my_tiles > 0) prefetch_ms((int)blockIdx.x * TP);
      }
      // (p0 advances by gridDim.x * TP with gridDim.x read as a constant-bank operand: written as j * gridDim.x ptxas hoisted
      // the factor into a register, SPILLED it, and reloaded it from local memory at the top of every round -- a load
      // that queues behind the twelve stores of the round before: 8 % of all stall samples, +6 us per launch.)
      int p0;
      if constexpr (STRIDED) p0 = (cta_x - cta_n) * TP;
      else p0 = (int)blockIdx.x * TP - (int)gridDim.x * TP;
      for (int j = 0; j < my_tiles; ++j) {
        if constexpr (STRIDED) p0 += cta_n * TP;
        else p0 += (int)gridDim.x * TP;
        const int rows = min(TP, P - p0);
        const int sj = j % NS, buf = j & 1;
        float* stage = raw + sj * a.raw_floats;
        float* gtile = fused ? dDs + buf * tile_elems : stage;
        if (has_atoms) {
          if (a.cc.use) tile_chan_update(tc, a.cc, p0);
          mbar_wait(dD_done + buf, (j >> 1) & 1);  // dD MMAs of tile j retired: its accumulator is complete
          tc_fence_after();
          if (warp == WARP_EPI) CHAIN(3, j);
          if (warp == WARP_EPI && j == 0) WSTAMP(9);
          // (plain dD output with dv: the stage held the D rows of the dictionary split, which every worker finished
          // before the MMAs of this tile were issued; without dv it is a staging buffer handed back by the loader)
          if (!fused && a.nraw == 0 && j >= NS) mbar_wait(full_raw + sj, ((j / NS) - 1) & 1);
          const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * TP + half * HALF);
          if constexpr (!FUSED) {
            // plain dD output: no prefetched moments in registers, so the whole share of the accumulator is read with
            // one wait and the tensor core gets it back before the shared-memory writes (57 -> 54 us)
            float r[HALF];
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 8) tmem_ld8_nowait(tcol + (uint32_t)c0, &r[c0]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);  // the accumulator is in registers: the tensor core may reuse it
            if (k_ok) {
              float* col = gtile + (half * HALF) * K + k;
#pragma unroll
              for (int i = 0; i < HALF; ++i) {
                const float sc = (a.cc.use && !G_SCALED) ? ((half * HALF + i) >= tc.bnd ? tc.rstd1 : tc.rstd0) : 1.0f;
                col[i * K] = __fmul_rn(r[i], sc);
              }
            }
          } else {
#pragma unroll
          for (int c0 = 0; c0 < HALF; c0 += 16) {  // sixteen columns per wait (all at once spills next to the moments: dearer)
            constexpr int NC = (HALF % 16 == 0) ? 16 : 8;  // (HALF is 8, 16, 24 or 32)
            const int nc = (HALF - c0 >= 16) ? 16 : 8;
            float r[16];
            tmem_ld8_nowait(tcol + (uint32_t)c0, &r[0]);
            if (NC == 16 || nc == 16) tmem_ld8_nowait(tcol + (uint32_t)(c0 + 8), &r[8]);
            tmem_ld_wait();
            if (k_ok) {
              float* col = gtile + (half * HALF + c0) * K + k;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (i < nc) {
                  const float sc = (a.cc.use && !G_SCALED) ? ((half * HALF + c0 + i) >= tc.bnd ? tc.rstd1 : tc.rstd0) : 1.0f;
                  col[i * K] = __fmul_rn(r[i], sc);  // (rows past a ragged end are written too: never stored)
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + buf);  // the accumulator has been read: the tensor core may reuse it
          }
          if (warp == WARP_EPI) CHAIN(4, j);
        }
        if (fused) {
          bar_sync(3, NEP * 32);  // gradient tile complete (and every warp is done with the tile two steps back)
          if (warp == WARP_EPI) CHAIN(11, j);
          mbar_wait(full_raw + sj, (j / NS) & 1);  // D rows landed
          if (warp == WARP_EPI) CHAIN(12, j);
          const int n4 = (rows * K) >> 2;
#ifndef ADIL_EXP_NO_EPI
#pragma unroll
          for (int u = 0; u < NPF; ++u) {
            const int e4 = etid + u * (NEP * 32);
            if (e4 < n4) {
              float4 Dv;
              if constexpr (STRIDED) Dv = *reinterpret_cast<const float4*>(stage + soff(e4));
              else Dv = *reinterpret_cast<const float4*>(stage + 4 * e4);
              const float4 gd = *reinterpret_cast<const float4*>(gtile + 4 * e4);
              adamw_update_fast(Dv.x, Mp[u].x, Sp[u].x, gd.x, a.hp);
              adamw_update_fast(Dv.y, Mp[u].y, Sp[u].y, gd.y, a.hp);
              adamw_update_fast(Dv.z, Mp[u].z, Sp[u].z, gd.z, a.hp);
              adamw_update_fast(Dv.w, Mp[u].w, Sp[u].w, gd.w, a.hp);
              if (a.atoms_mode == ADIL_ATOMS_CLAMP1) {
                Dv.x = clamp1(Dv.x); Dv.y = clamp1(Dv.y); Dv.z = clamp1(Dv.z); Dv.w = clamp1(Dv.w);
              }
              const size_t go = goff(p0, e4);
              *reinterpret_cast<float4*>(a.D2w + go) = Dv;
              *reinterpret_cast<float4*>(a.m + go) = Mp[u];
              *reinterpret_cast<float4*>(a.s + go) = Sp[u];
            }
          }
          for (int e4 = etid + NPF * (NEP * 32); e4 < n4; e4 += NEP * 32) {  // (large K: beyond the prefetched part)
            float4 Dv;
            if constexpr (STRIDED) Dv = *reinterpret_cast<const float4*>(stage + soff(e4));
            else Dv = *reinterpret_cast<const float4*>(stage + 4 * e4);
            const size_t go = goff(p0, e4);
            float4 Mv = ld_global4(a.m + go);
            float4 Sv = ld_global4(a.s + go);
            const float4 gd = *reinterpret_cast<const float4*>(gtile + 4 * e4);
            adamw_update_fast(Dv.x, Mv.x, Sv.x, gd.x, a.hp);
            adamw_update_fast(Dv.y, Mv.y, Sv.y, gd.y, a.hp);
            adamw_update_fast(Dv.z, Mv.z, Sv.z, gd.z, a.hp);
            adamw_update_fast(Dv.w, Mv.w, Sv.w, gd.w, a.hp);
            if (a.atoms_mode == ADIL_ATOMS_CLAMP1) {
              Dv.x = clamp1(Dv.x); Dv.y = clamp1(Dv.y); Dv.z = clamp1(Dv.z); Dv.w = clamp1(Dv.w);
            }
            *reinterpret_cast<float4*>(a.D2w + go) = Dv;
            *reinterpret_cast<float4*>(a.m + go) = Mv;
            *reinterpret_cast<float4*>(a.s + go) = Sv;
          }
#endif
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_raw + sj);  // this warp is done with the D rows of the stage
          if constexpr (STRIDED) {
            if (j + 1 < my_tiles) prefetch_ms(p0 + cta_n * TP);
          } else {
            if (j + 1 < my_tiles) prefetch_ms(p0 + (int)gridDim.x * TP);
          }  // the moments of the next tile fly while its MMAs run
        } else {
          if constexpr (STRIDED) {
            // Column window, plain dD output: the rows of the tile are K floats at a pitch of ldk in global memory.  As
            // one TMA bulk store per row (issued and waited for by the loader, which then requested the next dictionary
            // tile late) they were the slowest stage of the kernel -- K = 200: 192 us against 178 us for the FUSED step,
            // which moves four times the bytes.  The warps that hold atoms store the tile themselves, 128 bits per lane.
            const int nq = (K + 31) >> 5;  // TMEM quadrants that hold atoms: 2 nq warps take part
            bar_sync(3, 2 * nq * 32);      // the tile is complete in the stage
            const int n4 = (rows * K) >> 2;
            for (int e4 = (half * nq + quad) * 32 + lane; e4 < n4; e4 += 2 * nq * 32) {
              float4 t = *reinterpret_cast<const float4*>(stage + 4 * e4);
              float* gp = a.dD2 + goff(p0, e4);
              if (a.accumulate) {  // (chunks of a large batch run in stream order: plain read-modify-write)
                const float4 o = ld_global4(gp);
                t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
              }
              *reinterpret_cast<float4*>(gp) = t;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(epi_done + sj);  // this warp is done with the stage
          } else {
            fence_proxy_async();  // the dD tile in the stage is read by the copy engine
            __syncwarp();
            if (lane == 0) mbar_arrive(epi_done + sj);
          }
        }
        if (warp == WARP_EPI) CHAIN(5, j);
        if (warp == WARP_EPI && j == 0) WSTAMP(10);
        if (warp == WARP_EPI && j == my_tiles - 1) WSTAMP(11);
      }
    }
  } else {
    // ===== workers: stage the operand images of tile `it` while the MMAs of tile it-1 run =====
    const int quad = warp & 3, cg = warp >> 2;
    const int kv = K / a.vk;
    const int nitems = TP * kv;

    TileChan tc;
    tile_chan_init(tc);
    STAMP(3);
    for (int it = 0; it < my_tiles; ++it) {
      const int p0 = tile_p0(it);
      const int s = it % NS;
      const int buf = it & 1;
      bf16_t* Gb = Gi + buf * gbuf;
      bf16_t* Db = Di + buf * dbuf;
      if (it >= 2) {
        mbar_wait(mma_done + buf, ((it - 2) >> 1) & 1);  // MMAs(it-2) retired: this image buffer is free again
        tc_fence_after();
      }
      // gradient tile: registers -> three bf16 images (G_SCALED: of gx = g * (1/std); a group of 4 pixels never straddles
      // a channel)
      if (G_SCALED && a.cc.use) tile_chan_update(tc, a.cc, p0);
#pragma unroll
      for (int j = 0; j < GJ; ++j) {
        if (gsrc[j] >= 0) {
          const float sc = (G_SCALED && a.cc.use) ? (gsrc[j] >= tc.bnd ? tc.rstd1 : tc.rstd0) : 1.0f;
          const float val[4] = {__fmul_rn(greg[j].x, sc), __fmul_rn(greg[j].y, sc), __fmul_rn(greg[j].z, sc),
                                __fmul_rn(greg[j].w, sc)};
          uint32_t w0[4], w1[4], w2[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split_bf16x3(val[i], w0[i], w1[i], w2[i]);
          bf16_t* dst = Gb + gdst[j];
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_hi16(w0[0], w0[1]), pack_hi16(w0[2], w0[3]));
          *reinterpret_cast<uint2*>(dst + a.gimg) = make_uint2(pack_hi16(w1[0], w1[1]), pack_hi16(w1[2], w1[3]));
          *reinterpret_cast<uint2*>(dst + 2 * a.gimg) = make_uint2(pack_hi16(w2[0], w2[1]), pack_hi16(w2[2], w2[3]));
        }
      }
      if (it + 1 < my_tiles) prefetch(it + 1);      // the registers are free again: next tile's rows fly from here on
      if (a.want_dv) {
        // dictionary tile (pre-update values): TMA-landed raw rows -> three bf16 images of D / std.  Rows past the end
        // of the array (ragged last tile) were never written: they are staged as zeros.
        if (warp == 0) { CHAIN(7, it - NS); CHAIN(10, it); }
        if (it == 0 && a.want_dD && !codes_in_dDs) {
          // the dictionary-image region staged the code rows until now: once they are in tensor memory it is zeroed
          // (contraction padding must be zero, rows past a ragged end finite)
          mbar_wait(codes_ready, 0);
          uint4* zd = reinterpret_cast<uint4*>(Di);
          for (int e = tid; e < (a.dreg >> 4); e += NT) zd[e] = make_uint4(0u, 0u, 0u, 0u);
          bar_sync(2, NT);
        }
        if (it == 0) STAMP(7);
        mbar_wait(full_raw + s, (it / NS) & 1);
        if (warp == 0) CHAIN(8, it - NS);
        if constexpr (G_SCALED) {
          // item = (pixel p, group of 8 atoms): 8 floats of one raw row -> one 16-byte store per bf16 term.  Consecutive
          // lanes take consecutive groups of a pixel: contiguous reads, conflict-free 128-bit writes.
          const float* rt = raw + s * a.raw_floats;
          const int rows = min(TP, P - p0);
          const int ngroups = (K + 7) / 8;
          // item order: consecutive lanes take consecutive PIXELS of one atom group unless K is a multiple of 8 -- the
          // row pitch K then maps the 8-float reads of 16 (8) lanes onto distinct banks and the 16-byte stores of 8 lanes
          // onto one 128-byte line of an image; with consecutive groups of one pixel both were 2-way conflicts (ncu: 0.66 M
          // of the kernel's 3.07 M shared-memory wavefronts).  K % 8 == 0: consecutive groups (contiguous reads).
          for (int e = tid; e < TP * ngroups; e += NT) {
            int p, gq;
            if (a.pfast) { gq = e / TP; p = e - gq * TP; }
            else { p = div_magic_dev(e, a.gdiv); gq = e - p * ngroups; }
            const int k0 = 8 * gq;
            float val[8];
            const float* src = rt + p * K + k0;
            if constexpr (STRIDED) src = rt + p * a.rpitch + (int)(wof & a.rmask) + k0;
            if (p < rows) {
              if (a.vk == 4) {
                const float4 lo4 = *reinterpret_cast<const float4*>(src);
                const float4 hi4 = (k0 + 4 < K) ? *reinterpret_cast<const float4*>(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                val[0] = lo4.x; val[1] = lo4.y; val[2] = lo4.z; val[3] = lo4.w;
                val[4] = hi4.x; val[5] = hi4.y; val[6] = hi4.z; val[7] = hi4.w;
              } else if (a.vk == 2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 t = (k0 + 2 * i < K) ? *reinterpret_cast<const float2*>(src + 2 * i) : make_float2(0.f, 0.f);
                  val[2 * i] = t.x; val[2 * i + 1] = t.y;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) val[i] = (k0 + i < K) ? src[i] : 0.0f;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) val[i] = 0.0f;
            }
            uint32_t w0[8], w1[8], w2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_bf16x3(val[i], w0[i], w1[i], w2[i]);
            bf16_t* dst = Db + gq * (a.Sd >> 1) + (p >> 3) * 64 + (p & 7) * 8;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_hi16(w0[0], w0[1]), pack_hi16(w0[2], w0[3]),
                                                        pack_hi16(w0[4], w0[5]), pack_hi16(w0[6], w0[7]));
            *reinterpret_cast<uint4*>(dst + a.dimg) = make_uint4(pack_hi16(w1[0], w1[1]), pack_hi16(w1[2], w1[3]),
                                                                 pack_hi16(w1[4], w1[5]), pack_hi16(w1[6], w1[7]));
            *reinterpret_cast<uint4*>(dst + 2 * a.dimg) = make_uint4(pack_hi16(w2[0], w2[1]), pack_hi16(w2[2], w2[3]),
                                                                     pack_hi16(w2[4], w2[5]), pack_hi16(w2[6], w2[7]));
          }
        } else {
        const float* rt = raw + s * a.raw_floats;
        const int nvalid = min(TP, P - p0) * kv;
        if (a.cc.use) tile_chan_update(tc, a.cc, p0);
        if (a.vk == 4) {
          // four items per thread in flight: the loads first, then four independent split chains
          for (int e0 = tid; e0 < nitems; e0 += 4 * NT) {
            float4 rawv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * NT;
              rawv[u] = e < nvalid ? *reinterpret_cast<const float4*>(rt + 4 * e) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * NT;
              if (e < nitems) {
                const int p = div_magic_dev(e, a.kdiv), k = (e - p * kv) * 4;
                float val[4] = {rawv[u].x, rawv[u].y, rawv[u].z, rawv[u].w};
                if (a.cc.use) {
                  const float sd = p >= tc.bnd ? tc.std1 : tc.std0, rs = p >= tc.bnd ? tc.rstd1 : tc.rstd0;
#pragma unroll
                  for (int i = 0; i < 4; ++i) val[i] = div_by_const(val[i], sd, rs);
                }
                uint32_t w0[4], w1[4], w2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_bf16x3(val[i], w0[i], w1[i], w2[i]);
                bf16_t* dst = Db + (k >> 3) * (a.Sd >> 1) + (p >> 3) * 64 + (p & 7) * 8 + (k & 7);
                *reinterpret_cast<uint2*>(dst) = make_uint2(pack_hi16(w0[0], w0[1]), pack_hi16(w0[2], w0[3]));
                *reinterpret_cast<uint2*>(dst + a.dimg) = make_uint2(pack_hi16(w1[0], w1[1]), pack_hi16(w1[2], w1[3]));
                *reinterpret_cast<uint2*>(dst + 2 * a.dimg) = make_uint2(pack_hi16(w2[0], w2[1]), pack_hi16(w2[2], w2[3]));
              }
            }
          }
        } else if (a.vk == 2) {
          for (int e0 = tid; e0 < nitems; e0 += 4 * NT) {
            float2 rawv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * NT;
              rawv[u] = e < nvalid ? *reinterpret_cast<const float2*>(rt + 2 * e) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = e0 + u * NT;
              if (e < nitems) {
                const int p = div_magic_dev(e, a.kdiv), k = (e - p * kv) * 2;
                float val[2] = {rawv[u].x, rawv[u].y};
                if (a.cc.use) {
                  const float sd = p >= tc.bnd ? tc.std1 : tc.std0, rs = p >= tc.bnd ? tc.rstd1 : tc.rstd0;
                  val[0] = div_by_const(val[0], sd, rs);
                  val[1] = div_by_const(val[1], sd, rs);
                }
                uint32_t w0[2], w1[2], w2[2];
                split_bf16x3(val[0], w0[0], w1[0], w2[0]);
                split_bf16x3(val[1], w0[1], w1[1], w2[1]);
                bf16_t* dst = Db + (k >> 3) * (a.Sd >> 1) + (p >> 3) * 64 + (p & 7) * 8 + (k & 7);
                *reinterpret_cast<uint32_t*>(dst) = pack_hi16(w0[0], w0[1]);
                *reinterpret_cast<uint32_t*>(dst + a.dimg) = pack_hi16(w1[0], w1[1]);
                *reinterpret_cast<uint32_t*>(dst + 2 * a.dimg) = pack_hi16(w2[0], w2[1]);
              }
            }
          }
        } else {
          for (int e = tid; e < nitems; e += NT) {
            const int p = div_magic_dev(e, a.kdiv), k = e - p * kv;
            float val = e < nvalid ? rt[e] : 0.0f;
            if (a.cc.use) val = div_by_const(val, p >= tc.bnd ? tc.std1 : tc.std0, p >= tc.bnd ? tc.rstd1 : tc.rstd0);
            uint32_t w0, w1, w2;
            split_bf16x3(val, w0, w1, w2);
            bf16_t* dst = Db + (k >> 3) * (a.Sd >> 1) + (p >> 3) * 64 + (p & 7) * 8 + (k & 7);
            dst[0] = (bf16_t)(w0 >> 16);
            dst[a.dimg] = (bf16_t)(w1 >> 16);
            dst[2 * a.dimg] = (bf16_t)(w2 >> 16);
          }
        }
              }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();                                 // the ragged loops of every lane are over
      if (lane == 0) {
        mbar_arrive(staged + buf);                  // hand the tile to the issuer warp, keep going
        if (a.want_dv) mbar_arrive(empty_raw + s);  // this warp is done with the raw dictionary rows
      }
      if (warp == 0) CHAIN(9, it);
    }
    STAMP(4);
    if (my_tiles > 0) {
      mbar_wait(mma_done + ((my_tiles - 1) & 1), ((my_tiles - 1) >> 1) & 1);  // every MMA of this CTA has retired
      tc_fence_after();
      STAMP(5);
      if (a.want_dv) {
        // dv accumulator (lane = image, column = atom) -> this CTA's slab of the partial buffer.  The slab is first
        // laid out flat in shared memory (the operand images are dead by now) so that it leaves as coalesced stores:
        // 148 CTAs writing 4-byte pieces at a 4K-byte stride cost several microseconds at the end of the kernel.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the reduction kernel may start launching
        float* stg = reinterpret_cast<float*>(Di);
        const int b = quad * 32 + lane;
        for (int c0 = cg * 16; c0 < a.Kp; c0 += 64) {  // warp-uniform
          float r[16];
          tmem_ld16(acc_dv + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
          if constexpr (STRIDED) {
            if (a.dv2 && my_tiles > 1) {  // (odd tiles accumulated in the second accumulator)
              float r2[16];
              tmem_ld16(acc_dv + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a.Kp + c0), r2);
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] += r2[i];
            }
          }
          if (b < B) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < K) stg[b * K + c0 + i] = r[i];
          }
        }
        tc_fence_before();
        bar_sync(2, NT);
        float* dst = a.partial + (size_t)blockIdx.x * B * K;
        const int n = B * K;
        if ((n & 3) == 0) {
          for (int e = tid; e < (n >> 2); e += NT)
            reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(stg)[e];
        } else {
          for (int e = tid; e < n; e += NT) dst[e] = stg[e];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_EPI + NE) tmem_dealloc(tmem_base, a.tmem_cols);
  STAMP(6);
#undef cta_x
#undef cta_n
#undef wof
}

// ---------------------------------------------------------------------------------------------------------
// host side: tile-size selection by shared-memory fit, launches
// ---------------------------------------------------------------------------------------------------------
uint32_t pow2_cols(int need) {
  uint32_t c = 32;
  while ((int)c < need) c <<= 1;
  return c;
}

int vec_width(int K) { return (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1); }
unsigned div_magic(int d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned long long)d - 1) / (unsigned long long)d); }

struct SynthPlan {
  int TP, Kp8, Sd, dimg, raw_floats;
  size_t smem;
  uint32_t tmem_cols;
  bool ok;
};

SynthPlan plan_synth(int B, int P, int K, int hw) {
  SynthPlan pl{};
  pl.ok = false;
  if (B < 1 || B > 128 || K < 1 || K > 224 || P % 4 != 0 || hw % 4 != 0) return pl;  // (codes: 2 Kp8 <= 448 TMEM columns)
  const int tps[4] = {64, 48, 32, 16};
  static const int max_tp = getenv("ADIL_SYNTH_MAX_TP") ? atoi(getenv("ADIL_SYNTH_MAX_TP")) : 64;  // tuning knob
  for (int i = 0; i < 4; ++i) {
    const int TP = tps[i];
    if (hw < TP || TP > max_tp) continue;  // a tile may span at most two channels
    pl.TP = TP;
    pl.Kp8 = rup(K, 8);
    pl.Sd = img_stride(TP);
    pl.dimg = (pl.Kp8 / 4) * (pl.Sd / 4);
    pl.raw_floats = rup(TP * K, 32);
    pl.smem = HDR_BYTES + sizeof(float) * ((size_t)NS * pl.raw_floats + 4 * (size_t)pl.dimg + (size_t)NSX * B * (TP + 4));
    pl.tmem_cols = pow2_cols(2 * TP + 2 * pl.Kp8);  // two accumulators + the codes (hi, lo)
    if (pl.smem <= (size_t)SMEM_LIMIT && pl.tmem_cols <= 512) { pl.ok = true; return pl; }
  }
  return pl;
}

struct GradPlan {
  int TP, Bp, Kp, Sg, Sd, dimg, gimg, raw_floats, nraw, dreg;
  size_t smem;
  uint32_t tmem_cols;
  bool ok;
};

// Tile cap of a column window of the fused step (two windows in one launch): measured at K = 136 / 176 / 200 / 256
// (windows of 68 / 88 / 100 / 128 atoms): TP = 32 beats 64 and 48 (160 vs 180 us, 168 vs 252, -, -), and a window whose
// 32-pixel tile has more than 3 x 320 float4 items (the ten-warp AdamW pass) is faster at TP = 16 (K = 256: 249 vs 294 us).
// The plain contractions of a window do not care (K = 200: 122 / 124 / 124 us at TP = 64 / 48 / 32; TP = 16 is slower): up to
// 120 atoms per window they take the fused step's tile, so that the code gradient of the multi-GPU path (plain
// contractions + sharded dictionary step) accumulates in the same order as the single-GPU fused step -- bit-identical dv.
int window_tp_cap(int Kwin, bool fused) {
  if ((32 * Kwin) / 4 <= 3 * NEP_MAX * 32) return 32;
  return fused ? 16 : 64;
}

GradPlan plan_grad(int B, int P, int K, int hw, bool want_dD, bool want_dv, bool fused, int tp_cap = 64) {
  GradPlan pl{};
  pl.ok = false;
  if (B < 1 || B > 128 || K < 1 || K > 128 || P % 4 != 0 || hw % 4 != 0) return pl;
  if (!want_dD && !want_dv) return pl;
  const int tps[4] = {64, 48, 32, 16};
  static const int max_tp = getenv("ADIL_GRAD_MAX_TP") ? atoi(getenv("ADIL_GRAD_MAX_TP")) : 64;  // tuning knob
  for (int i = 0; i < 4; ++i) {
    const int TP = tps[i];
    if (hw < TP || TP > max_tp || TP > tp_cap) continue;
    pl.TP = TP;
    pl.Bp = rup(B, 16);
    pl.Kp = rup(K, 16);
    pl.Sg = img_stride(pl.Bp);
    pl.Sd = img_stride(TP);
    pl.dimg = want_dv ? (pl.Kp / 8) * (pl.Sd / 2) : 0;
    pl.gimg = (TP / 8) * (pl.Sg / 2);
    pl.nraw = (fused || want_dv) ? 1 : 0;
    pl.raw_floats = (pl.nraw > 0 ? pl.nraw : 1) * TP * K;  // (dD only: the stage is the staging buffer of the output tile)
    // dictionary-image region: two buffers of three bf16 images; at kernel entry it stages the B code rows (fp32), and at
    // the end the [B][K] slab of the code gradient
    pl.dreg = rup(2 * 2 * 3 * pl.dimg, 128);
    if (pl.dreg < rup(4 * B * K, 128)) pl.dreg = rup(4 * B * K, 128);
    pl.smem = HDR_BYTES + (size_t)pl.dreg + 2 * 2 * (3 * (size_t)pl.gimg + 1024) +  // + two buffers of gradient images
              sizeof(float) * ((fused ? 2 * (size_t)TP * K : 0) + (size_t)NS * pl.raw_floats);
    pl.tmem_cols = pow2_cols(2 * TP + pl.Kp + 3 * (pl.Bp / 2));  // accumulators + the codes
    if (pl.smem <= (size_t)SMEM_LIMIT && pl.tmem_cols <= 512) { pl.ok = true; return pl; }
  }
  return pl;
}

template <typename KernelT>
int set_smem(KernelT kern, size_t smem, const char* what) {
  return check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), what);
}

template <int TP, bool TRAIN>
int launch_synth_tp(const SynthArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (a.stage_codes) {
    int rc = set_smem(synth_kernel<TP, TRAIN, true>, smem, "cudaFuncSetAttribute(synth_kernel)");
    if (rc) return rc;
    synth_kernel<TP, TRAIN, true><<<grid, NTHREADS_SYNTH, smem, st>>>(a);
    return 0;
  }
  int rc = set_smem(synth_kernel<TP, TRAIN, false>, smem, "cudaFuncSetAttribute(synth_kernel)");
  if (rc) return rc;
  synth_kernel<TP, TRAIN, false><<<grid, NTHREADS_SYNTH, smem, st>>>(a);
  return 0;
}

}  // namespace

bool tc_synth_ok(int B, int P, int K, int hw) { return plan_synth(B > 128 ? 128 : B, P, K, hw).ok; }

int launch_synth_tc(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2,
                    const float* v, const int64_t* v_index, float* codes_out, int B, int P, int K,
                    const ChannelConsts& cc, float eps, int flags, cudaStream_t st) {
  const bool norm = (flags & ADIL_SYNTH_NORMALIZE) != 0;
  const int hw = norm ? cc.hw : P;
  const bool x_index_on_host = x_index && is_host_pointer(x_index), v_index_on_host = v_index && is_host_pointer(v_index);
  // images beyond 128 go through further launches (M = 128 image lanes per pass)
  for (int b0 = 0; b0 < B; b0 += 128) {
    const int nb = B - b0 < 128 ? B - b0 : 128;
    const SynthPlan pl = plan_synth(nb, P, K, hw);
    if (!pl.ok) return set_error(-4, "adil_synth: shape B=%d P=%d K=%d does not qualify for the tcgen05 path", B, P, K);
    SynthArgs a;
    a.out = out ? out + (size_t)b0 * P : nullptr;
    a.delta = delta_out ? delta_out + (size_t)b0 * P : nullptr;
    a.codes_out = codes_out ? codes_out + (size_t)b0 * K : nullptr;
    a.x = x ? (x_index ? x : x + (size_t)b0 * P) : nullptr;
    a.xidx = x_index ? x_index + b0 : nullptr;
    a.hx_on = 0; a.hv_on = 0;
    if (x_index && x_index_on_host) {
      for (int i = 0; i < nb; ++i) a.hx[i] = (int)x_index[b0 + i];
      a.hx_on = 1;
    }
    if (v_index && v_index_on_host) {
      for (int i = 0; i < nb; ++i) a.hv[i] = (int)v_index[b0 + i];
      a.hv_on = 1;
    }
    a.D2 = D2;
    a.v = v_index ? v : v + (size_t)b0 * K;
    a.vidx = v_index ? v_index + b0 : nullptr;
    if (a.hx_on) a.xidx = nullptr;
    if (a.hv_on) a.vidx = nullptr;
    a.B = nb; a.P = P; a.K = K; a.Kp8 = pl.Kp8; a.Sd = pl.Sd; a.dimg = pl.dimg;
    a.raw_floats = pl.raw_floats; a.vk = vec_width(K); a.kdiv = div_magic(K / a.vk); a.tmem_cols = pl.tmem_cols;
    a.eps = eps; a.flags = flags; a.cc = cc;
    a.cc.use = norm ? 1 : 0;
    {  // code-row gather: widest cp.async / shared-memory read the alignment of v and K allows; rows staged K (+4) apart
      const uintptr_t va = reinterpret_cast<uintptr_t>(a.v);
      a.cw = (va % 16 == 0 && K % 4 == 0) ? 4 : ((va % 8 == 0 && K % 2 == 0) ? 2 : 1);
      a.cpitch = K + ((K % 8 == 0) ? 4 : 0);
      static const int stage_knob = getenv("ADIL_SYNTH_STAGE_CODES") ? atoi(getenv("ADIL_SYNTH_STAGE_CODES")) : 1;  // A/B knob
      a.stage_codes = (stage_knob != 0 && (size_t)nb * a.cpitch <= 2 * (size_t)pl.dimg) ? 1 : 0;
      a.rpw = (nb + NTHREADS_SYNTH / 32 - 1) / (NTHREADS_SYNTH / 32);
    }
    const int ntiles = (P + pl.TP - 1) / pl.TP;
    int grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    const bool train = a.out != nullptr && a.x != nullptr && a.delta == nullptr &&
                       (flags & (ADIL_SYNTH_CLAMP_DELTA | ADIL_SYNTH_CLAMP01)) == 0;
    int rc = 0;
    switch (pl.TP) {
      case 64: rc = train ? launch_synth_tp<64, true>(a, pl.smem, grid, st) : launch_synth_tp<64, false>(a, pl.smem, grid, st); break;
      case 48: rc = train ? launch_synth_tp<48, true>(a, pl.smem, grid, st) : launch_synth_tp<48, false>(a, pl.smem, grid, st); break;
      case 32: rc = train ? launch_synth_tp<32, true>(a, pl.smem, grid, st) : launch_synth_tp<32, false>(a, pl.smem, grid, st); break;
      default: rc = train ? launch_synth_tp<16, true>(a, pl.smem, grid, st) : launch_synth_tp<16, false>(a, pl.smem, grid, st); break;
    }
    if (rc) return rc;
    rc = check_cuda(cudaGetLastError(), "synth_kernel launch");
    if (rc) return rc;
#if defined(ADIL_CHAIN) && !defined(ADIL_CHAIN_QUIET)
    {
      cudaDeviceSynchronize();
      long long h[16];
      cudaMemcpyFromSymbol(h, g_sstamp, sizeof(h));
      fprintf(stderr, "synth CTA0 stamps (ns from entry): sync1=%lld sync2=%lld codes_in_tmem=%lld staged0=%lld mma0_done=%lld x0_landed=%lld out0_ready=%lld "
              "out0_stored=%lld last_out_ready=%lld last_stored=%lld exit=%lld | init+tma=%lld tmem_alloc=%lld\n", h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0],
              h[6] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0], h[10] - h[0], h[11] - h[0], h[13] - h[0], h[12] - h[0]);
    }
#endif
#ifdef ADIL_TIMING
    {
      cudaDeviceSynchronize();
      long long h[16];
      cudaMemcpyFromSymbol(h, g_tim, sizeof(h));
      long long z[16] = {0};
      cudaMemcpyToSymbol(g_tim, z, sizeof(z));
      fprintf(stderr, "synth TP=%d tiles=%lld cycles/tile:", pl.TP, h[12]);
      const char* nm[10] = {"wait_mma", "wait_raw", "split+arrive", "wait_x", "tmem_ld", "phase1", "bar", "load_x", "store", "tail_wait"};
      long long tot = 0;
      for (int i = 0; i < 10; ++i) { fprintf(stderr, " %s=%.0f", nm[i], (double)h[i] / (double)h[12]); tot += h[i]; }
      fprintf(stderr, " | total=%.0f\n", (double)tot / (double)h[12]);
    }
#endif
  }
  return 0;
}

namespace {
template <int TP, bool FUSED, bool STRIDED, int NPF>
int launch_grad_tpfn(const GradArgs& a, size_t smem, int grid, cudaStream_t st) {
  int rc = set_smem(grad_kernel<TP, FUSED, STRIDED, NPF>, smem, "cudaFuncSetAttribute(grad_kernel)");
  if (rc) return rc;
  grad_kernel<TP, FUSED, STRIDED, NPF><<<grid, NTHREADS_GRAD, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "grad_kernel launch");
}
template <int TP, bool FUSED, bool STRIDED>
int launch_grad_tpf(const GradArgs& a, size_t smem, int grid, cudaStream_t st) {
  static const int npf_knob = getenv("ADIL_GRAD_NPF") ? atoi(getenv("ADIL_GRAD_NPF")) : 0;  // tuning knob: 4 forces the 8-warp pass
  // (knob 3: the ten-warp pass also for tiles of up to 4 * 320 items -- the items beyond 3 * 320 run unprefetched)
  if (FUSED && npf_knob != 4 && ((TP * a.K) / 4 <= 3 * NEP_MAX * 32 || (npf_knob == 3 && (TP * a.K) / 4 <= 4 * NEP_MAX * 32))) return launch_grad_tpfn<TP, FUSED, STRIDED, FUSED ? 3 : 4>(a, smem, grid, st);
  return launch_grad_tpfn<TP, FUSED, STRIDED, 4>(a, smem, grid, st);
}
template <int TP>
int launch_grad_tp(const GradArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (a.ldk != a.K)
    return a.D2w != nullptr ? launch_grad_tpf<TP, true, true>(a, smem, grid, st) : launch_grad_tpf<TP, false, true>(a, smem, grid, st);
  return a.D2w != nullptr ? launch_grad_tpf<TP, true, false>(a, smem, grid, st) : launch_grad_tpf<TP, false, false>(a, smem, grid, st);
}

// One launch over a window of K columns (K <= 128) of arrays whose rows are ldk floats apart; dvb rows are dv_ld apart.
int launch_grad_window(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                       const float* v, const int64_t* v_index, int B, int P, int K, int ldk, int dv_ld, int nwin,
                       const ChannelConsts& cc, const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes,
                       const GradOpts& opt, cudaStream_t st);
}  // namespace

bool tc_grad_ok(int B, int P, int K, int hw, bool want_dD, bool want_dv, bool fused) {
  if (K <= 128) return plan_grad(B, P, K, hw, want_dD, want_dv, fused).ok;
  // more than 128 atoms: two column windows of K/2 atoms (row pitch and window offsets must keep 16-byte alignment)
  if (K > 256 || K % 8 != 0) return false;
  return plan_grad(B, P, K / 2, hw, want_dD, want_dv, fused, window_tp_cap(K / 2, fused)).ok;
}

int launch_grad_tc(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                   const float* v, const int64_t* v_index, int B, int P, int K, const ChannelConsts& cc,
                   const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes, const GradOpts& opt,
                   cudaStream_t st) {
  if (K <= 128)
    return launch_grad_window(dD2, D2_rw, m, s, dvb, g, D2, v, v_index, B, P, K, K, K, 1, cc, hp, atoms_mode, scratch,
                              scratch_bytes, opt, st);
  if (opt.keep_partials)
    return set_error(-5, "adil_grad: ADIL_GRAD_KEEP_PARTIALS is limited to K <= 128 on the tcgen05 path (K=%d runs as two "
                     "column windows): pass dvb", K);
  const int Kh = K / 2;
  // Both windows in ONE launch (CTA parity = window; see grad_kernel) unless ADIL_GRAD_WINDOWS=serial asks for the former
  // two launches (A/B runs; each window then reads g from HBM once more).
  static const bool serial = getenv("ADIL_GRAD_WINDOWS") && !strcmp(getenv("ADIL_GRAD_WINDOWS"), "serial");
  if (!serial && sm_count() >= 2)
    return launch_grad_window(dD2, D2_rw, m, s, dvb, g, D2, v, v_index, B, P, Kh, K, K, 2, cc, hp, atoms_mode, scratch,
                              scratch_bytes, opt, st);
  for (int h = 0; h < 2; ++h) {
    const int k0 = h * Kh;
    int rc = launch_grad_window(dD2 ? dD2 + k0 : nullptr, D2_rw ? D2_rw + k0 : nullptr, m ? m + k0 : nullptr,
                                s ? s + k0 : nullptr, dvb ? dvb + k0 : nullptr, g, D2 + k0, v + k0, v_index, B, P, Kh, K,
                                K, 1, cc, hp, atoms_mode, scratch, scratch_bytes, opt, st);
    if (rc) return rc;
  }
  return 0;
}

namespace {
int launch_grad_window(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                       const float* v, const int64_t* v_index, int B, int P, int K, int ldk, int dv_ld, int nwin,
                       const ChannelConsts& cc, const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes,
                       const GradOpts& opt, cudaStream_t st) {
  const bool want_dD = dD2 != nullptr || D2_rw != nullptr, want_dv = dvb != nullptr || opt.keep_partials != 0,
             fused = D2_rw != nullptr;
  if (!want_dD && !want_dv) return 0;
  const int hw = cc.use ? cc.hw : P;
  const GradPlan pl = plan_grad(B, P, K, hw, want_dD, want_dv, fused, nwin == 2 ? window_tp_cap(K, fused) : 64);
  if (!pl.ok) return set_error(-4, "adil_grad: shape B=%d P=%d K=%d does not qualify for the tcgen05 path", B, P, K);
  GradArgs a;
  a.dD2 = dD2; a.D2w = D2_rw; a.m = m; a.s = s; a.partial = scratch; a.g = g; a.D2 = D2; a.v = v; a.vidx = v_index;
  a.hv_on = 0;
  if (v_index && is_host_pointer(v_index)) {
    for (int i = 0; i < B; ++i) a.hv[i] = (int)v_index[i];
    a.hv_on = 1;
    a.vidx = nullptr;
  }
  a.B = B; a.P = P; a.K = K; a.Bp = pl.Bp; a.Kp = pl.Kp; a.Sg = pl.Sg; a.Sd = pl.Sd;
  a.dimg = pl.dimg; a.gimg = pl.gimg; a.raw_floats = pl.raw_floats; a.nraw = pl.nraw;
  size_t smem_bytes = pl.smem;
  a.fullrow = 0; a.rpitch = K; a.rmask = 0u;
  {  // column windows: full dictionary rows in the raw stages when the wider stages fit (A/B knob: ADIL_GRAD_FULLROW=0)
    static const bool fullrow_knob = !(getenv("ADIL_GRAD_FULLROW") && atoi(getenv("ADIL_GRAD_FULLROW")) == 0);
    const size_t wider = pl.smem + sizeof(float) * (size_t)NS * pl.TP * (size_t)(ldk - K);
    if (ldk != K && pl.nraw > 0 && fullrow_knob && ADIL_G_SCALED_FUSED && ADIL_G_SCALED_PLAIN && wider <= (size_t)SMEM_LIMIT) {
      a.fullrow = 1; a.rpitch = ldk; a.rmask = 0xffffffffu;
      a.raw_floats = pl.TP * ldk;
      smem_bytes = wider;
    }
  }
  a.vk = vec_width(K); a.kdiv = div_magic(K / a.vk); a.gdiv = div_magic((K + 7) / 8); a.tmem_cols = pl.tmem_cols;
  a.ldk = ldk; a.k4div = div_magic(K >> 2 > 0 ? K >> 2 : 1); a.nwin = nwin; a.wsh = nwin >> 1;
  {  // column windows in one launch: two dv accumulators when tensor memory has the columns (A/B knob: ADIL_GRAD_DV2=0).
     // (Measured on the single-window kernels too, K = 50 ... 128: no change -- they keep one accumulator.)
    static const bool dv2_knob = !(getenv("ADIL_GRAD_DV2") && atoi(getenv("ADIL_GRAD_DV2")) == 0);
    a.dv2 = (dv2_knob && nwin == 2 && want_dv && 2 * pl.TP + 2 * pl.Kp + 3 * (pl.Bp / 2) <= 512) ? 1 : 0;
  }
  if (a.dv2) a.tmem_cols = pow2_cols(2 * pl.TP + 2 * pl.Kp + 3 * (pl.Bp / 2));
  a.want_dD = want_dD ? 1 : 0; a.want_dv = want_dv ? 1 : 0; a.atoms_mode = atoms_mode; a.cc = cc;
  a.accumulate = (opt.accumulate && !fused) ? 1 : 0;
  a.dreg = pl.dreg;
  {  // widest cp.async the code-row gather may use: source rows start at v + row * ldk, destination rows at b * K floats
    const uintptr_t va = reinterpret_cast<uintptr_t>(v);
    a.cw = (va % 16 == 0 && ldk % 4 == 0 && K % 4 == 0) ? 4 : ((va % 8 == 0 && ldk % 2 == 0 && K % 2 == 0) ? 2 : 1);
    a.rpw = (B + 24) / 25;
    a.pfast = (K % 8 != 0) ? 1 : 0;
    a.codes_contig = (v_index == nullptr && ldk == K && va % 16 == 0 && (B * K) % 4 == 0) ? 1 : 0;
  }
  if (hp) a.hp = *hp;
  const int ntiles = (P + pl.TP - 1) / pl.TP;
  int grid = sm_count();
  if (grid > kMaxGradCtas) grid = kMaxGradCtas;
  if (nwin == 2) {  // CTA pairs: (window 0, window 1) of the same pixel tiles
    grid &= ~1;
    if (grid > 2 * ntiles) grid = 2 * ntiles;
  } else if (grid > ntiles) {
    grid = ntiles;
  }
  if (want_dv) {
    const size_t need = (size_t)grid * B * K * sizeof(float);
    if (scratch == nullptr || scratch_bytes < need)
      return set_error(-2, "adil_grad: scratch too small (%zu < %zu bytes)", scratch_bytes, need);
  }
  int rc;
  switch (pl.TP) {
    case 64: rc = launch_grad_tp<64>(a, smem_bytes, grid, st); break;
    case 48: rc = launch_grad_tp<48>(a, smem_bytes, grid, st); break;
    case 32: rc = launch_grad_tp<32>(a, smem_bytes, grid, st); break;
    default: rc = launch_grad_tp<16>(a, smem_bytes, grid, st); break;
  }
  if (rc) return rc;
#if defined(ADIL_CHAIN) && !defined(ADIL_CHAIN_QUIET)
  {
    cudaDeviceSynchronize();
    long long ch[16];
    cudaMemcpyFromSymbol(ch, g_chain, sizeof(ch));
    fprintf(stderr, "grad CTA5 tile 6 chain (ns after worker D-split start of tile 6): issuer_waits=%lld staged=%lld mma_issued=%lld E_sees_done=%lld E_ld_done=%lld E_bar=%lld E_raw_landed=%lld E_adamw_done=%lld | worker_staged6=%lld worker_at_raw9=%lld raw9_landed=%lld\n",
            ch[0] - ch[10], ch[1] - ch[10], ch[2] - ch[10], ch[3] - ch[10], ch[4] - ch[10], ch[11] - ch[10], ch[12] - ch[10], ch[5] - ch[10], ch[9] - ch[10], ch[7] - ch[10], ch[8] - ch[10]);
  }
#endif
#if defined(ADIL_CHAIN) && !defined(ADIL_CHAIN_QUIET)
  {
    long long st8[16];
    cudaMemcpyFromSymbol(st8, g_stamp, sizeof(st8));
    long long w[16];
    cudaMemcpyFromSymbol(w, g_wstamp, sizeof(w));
    fprintf(stderr, "grad CTA0 stamps (ns from entry): W: g0_requested+mbar_init=%lld g_zeroed=%lld sync1=%lld D_region_zeroed=%lld loop_end=%lld last_mma_seen=%lld exit=%lld | "
            "E: codes_requested=%lld codes_landed=%lld bar=%lld codes_in_tmem=%lld dD0_seen=%lld epi0_done=%lld epi_last_done=%lld | alloc=%lld | L: tma_issued=%lld | "
            "I: staged0=%lld codes_ready=%lld mma0_issued=%lld | E gather: start=%lld row0=%lld row3=%lld\n",
            st8[8] - st8[0], st8[12] - st8[0], st8[10] - st8[0], st8[7] - st8[0], st8[4] - st8[0], st8[5] - st8[0], st8[6] - st8[0],
            w[0] - st8[0], w[3] - st8[0], w[4] - st8[0], w[5] - st8[0], w[9] - st8[0], w[10] - st8[0], w[11] - st8[0], w[1] - st8[0], w[2] - st8[0],
            w[6] - st8[0], w[7] - st8[0], w[8] - st8[0], w[12] - st8[0], w[13] - st8[0], w[14] - st8[0]);
  }
#endif
  if (want_dv) {
    if (opt.keep_partials) {
      if (opt.nslabs_out) *opt.nslabs_out = grid;
      return 0;
    }
    // slabs of window h: CTAs h, h + 2, ... -> columns [h K, h K + K) of dvb (one launch, blockIdx.y = window)
    if (nwin == 2) return launch_reduce_partials(dvb, scratch, B * K, grid / 2, K, dv_ld, st, 2);
    return launch_reduce_partials(dvb, scratch, B * K, grid, K, dv_ld, st);
  }
  return 0;
}
}  // namespace

}  // namespace adil
