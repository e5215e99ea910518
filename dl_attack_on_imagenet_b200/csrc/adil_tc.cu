// tcgen05 (5th-gen tensor core) kernels of the ADiL hot path -- split-TF32 ("3xTF32") contractions with fp32
// accumulators in TMEM, used when the atom count makes the CUDA-core FFMA pipe the limiter.
//
// Every fp32 operand x is split exactly into hi = x & 0xffffe000 (representable in TF32) and lo = x - hi, and each
// contraction is issued as three tcgen05.mma.kind::tf32 passes  lo*hi + hi*lo + hi*hi  into the same TMEM
// accumulator: relative error ~2^-21, i.e. fp32-grade (the 1e-5 parity bound of the north star needs it: plain TF32
// is 2^-11).
//
// Operand staging.  All three matrices (dictionary tile D[p][k], gradient tile g[b][p], batch codes v[b][k]) are laid
// out in shared memory by the CTA's threads in the UMMA no-swizzle canonical form: 128-byte core matrices of
// 8 "rows" x 16 bytes.  One image serves a matrix in both of its roles because the core matrix of a K-major operand
// (8 M/N-rows x 4 contiguous K-elements) and of an MN-major operand (8 K-rows x 4 contiguous MN-elements) is the
// same memory pattern:
//     img(r, c) = (c/4)*S + (r/8)*128 + (r%8)*16 + (c%4)*4        S = 128*ceil(R/8) + 16  (the +16 de-phases banks)
//     D tile : r = pixel, c = atom   -> A of synth (K-major, M=pixel)      / B of dv   (MN-major, N=atom)
//     g tile : r = image, c = pixel  -> A of dD    (MN-major, M=pixel)     / A of dv   (K-major,  M=image)
//     codes  : r = image, c = atom   -> B of synth (K-major, N=image)      / B of dD   (MN-major, N=atom)
// The synthesis kernel stages the raw dictionary tile with a 1-D TMA bulk copy (cp.async.bulk + mbarrier,
// double-buffered) before the split; the gradient kernel prefetches the next tile through registers because its
// shared memory is taken by the hi/lo images.
#include <cstdio>

#include "adil_common.cuh"

namespace adil {

int launch_reduce_partials(float* dvb, const float* partial, int n, int nslabs, cudaStream_t st);

namespace {

constexpr int TC_WARPS = 15;                  // worker warps: staging, prefetch, epilogues
constexpr int TC_THREADS = TC_WARPS * 32;
constexpr int TC_BLOCK = TC_THREADS + 32;     // + warp 15, which only issues TMA / tcgen05.mma (512 threads, 128 regs)
constexpr int TC_TP = 128;                    // pixels per tile = UMMA M
constexpr int SMEM_LIMIT = 227 * 1024;

__host__ __device__ inline int rup(int a, int b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
      printf("adil_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// named barriers: id 1 = "tile staged" hand-off workers -> issuer warp, id 2 = worker-only barrier
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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
// instruction descriptor for kind::tf32, fp32 accumulate (cute/arch/mma_sm100_desc.hpp: InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 8 consecutive fp32 columns -> 8 registers per thread (thread t <-> TMEM lane base+t)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = __fsub_rn(x, hi);  // exact
}

// canonical image offset in floats: r = "8-row" index, c = contiguous index, S = byte stride between c-groups
__device__ __forceinline__ int img_off(int r, int c, int S_bytes) {
  return (c >> 2) * (S_bytes >> 2) + (r >> 3) * 32 + (r & 7) * 4 + (c & 3);
}
__host__ __device__ inline int img_stride(int R) { return 128 * ((R + 7) / 8) + 16; }

// codes image (hi/lo): rows b < Rz are written (zero beyond B / K), layout img(b, k)
__device__ void build_code_images(float* Vhi, float* Vlo, const float* v, const int64_t* vidx, int B, int K, int Rz,
                                  int Kz, int Sv) {
  for (int e = threadIdx.x; e < Rz * Kz; e += blockDim.x) {
    const int b = e / Kz, k = e - b * Kz;
    float val = 0.0f;
    if (b < B && k < K) {
      const int64_t row = vidx ? vidx[b] : (int64_t)b;
      val = v[row * K + k];
    }
    float hi, lo;
    split_tf32(val, hi, lo);
    const int o = img_off(b, k, Sv);
    Vhi[o] = hi;
    Vlo[o] = lo;
  }
}

// =========================================================================================================
// synthesis:  acc[p, b] = sum_k D[p,k] v[b,k]   (M = 128 pixels, N = images, K = atoms)
// =========================================================================================================
struct SynthTcArgs {
  float* out;
  float* delta;
  const float* x;
  const int64_t* xidx;
  const float* D2;
  const float* v;
  const int64_t* vidx;
  int B, P, K;
  int Np;      // N of the MMA: round_up(B, 16)
  int Kp8;     // contraction length: round_up(K, 8)
  int Sd, Sv;  // image strides (bytes)
  int raw_floats, dimg_floats, vimg_floats;
  uint32_t tmem_cols;
  float eps;
  int flags;
  ChannelConsts cc;
};

constexpr int S_OS = TC_TP + 4;  // row stride (floats) of the epilogue staging tile [image][pixel]
constexpr int S_EB = 7;          // images per thread per epilogue batch (7 x 16 warps >= 100)

__global__ void __launch_bounds__(TC_BLOCK, 1) synth_tc_kernel(const SynthTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_raw);  // [1] raw tile landed
  uint64_t* bar_mma = bar_full + 1;                            // [1] MMAs of a tile retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
  long long* xoff_s = reinterpret_cast<long long*>(smem_raw + 128);  // [256] element offset of each image's x row
  float* raw = reinterpret_cast<float*>(smem_raw + 128 + 2048);     // TMA landing buffer of the raw D tile
  float* Vhi = raw + a.raw_floats;
  float* Vlo = Vhi + a.vimg_floats;
  float* Dhi = Vlo + a.vimg_floats;
  float* Dlo = Dhi + a.dimg_floats;
  float* outs = Dlo + a.dimg_floats;                           // [Np][S_OS] accumulator tile, image-major

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, P = a.P, B = a.B;
  const int ntiles = (P + TC_TP - 1) / TC_TP;

  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, a.tmem_cols);
  // zero the dictionary images once: contraction padding (k in [K, Kp8)) must stay zero
  for (int e = tid; e < 2 * a.dimg_floats; e += TC_BLOCK) Dhi[e] = 0.0f;
  build_code_images(Vhi, Vlo, a.v, a.vidx, B, K, a.Np, a.Kp8, a.Sv);
  // x row offsets in shared memory: a dependent global load per image inside the epilogue would serialise it
  for (int b = tid; b < B; b += TC_BLOCK) xoff_s[b] = (a.xidx ? (long long)a.xidx[b] : (long long)b) * (long long)P;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_first = blockIdx.x;
  if (tid == 0 && my_first < ntiles) {
    const int rows = min(TC_TP, P - my_first * TC_TP);
    mbar_expect_tx(bar_full, (uint32_t)(rows * K * 4));
    bulk_g2s(raw, a.D2 + (size_t)my_first * TC_TP * K, (uint32_t)(rows * K * 4), bar_full);
  }

  const uint32_t idesc = make_idesc(128, a.Np, false, false);
  const int ksteps = a.Kp8 / 8;
  const int quad = warp & 3, cgrp = warp >> 2;          // TMEM lane quadrant / column group of this warp
  const int ncg = (quad == 3) ? 3 : 4;                  // quadrant 3 lost warp 15 to the issuer role
  const int nchunks = a.Np / 8;                         // 8-column chunks of the accumulator
  const bool need_x = (a.x != nullptr) && (a.out != nullptr);

  auto epilogue = [&](int tile, int it) {
    bar_sync(2, TC_THREADS);  // staging tile free (previous epilogue fully drained)
    // Phase A: accumulator (row = pixel = TMEM lane, column = image) -> shared memory, image-major
    {
      const uint32_t acc = tmem_base + (uint32_t)((it & 1) * a.Np) + ((uint32_t)(quad * 32) << 16);
      float* col = outs + quad * 32 + lane;
      for (int ch = cgrp; ch < nchunks; ch += ncg) {    // warp-uniform
        float d[8];
        tmem_ld8(acc + (uint32_t)(ch * 8), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) col[(ch * 8 + j) * S_OS] = d[j];
      }
      tc_fence_before();
    }
    bar_sync(2, TC_THREADS);
    // Phase B: coalesced 128-bit pass over image rows of 128 pixels: x + delta, clamps, (.-mean)/std, store.
    // A thread keeps one 4-pixel column (channel constants hoisted) and walks images b = warp, warp+16, ...
    const int pq = lane, p = tile * TC_TP + 4 * pq;
    if (p < P) {
      float mean[4], stdv[4], rstd[4];
      if (a.cc.use) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = (p + j) / a.cc.hw;
          mean[j] = a.cc.mean[c]; stdv[j] = a.cc.stdv[c]; rstd[j] = a.cc.rstd[c];
        }
      }
      for (int b0 = warp; b0 < B; b0 += S_EB * TC_WARPS) {
        float4 xv[S_EB];
        if (need_x) {
#pragma unroll
          for (int i = 0; i < S_EB; ++i) {
            const int b = b0 + i * TC_WARPS;
            if (b < B) xv[i] = ld_stream4(a.x + xoff_s[b] + p);
          }
        }
#pragma unroll
        for (int i = 0; i < S_EB; ++i) {
          const int b = b0 + i * TC_WARPS;
          if (b < B) {
            const float4 d4 = *reinterpret_cast<const float4*>(outs + b * S_OS + 4 * pq);
            float d[4] = {d4.x, d4.y, d4.z, d4.w};
            if (a.flags & ADIL_SYNTH_CLAMP_DELTA) {
#pragma unroll
              for (int j = 0; j < 4; ++j) d[j] = fminf(fmaxf(d[j], -a.eps), a.eps);
            }
            if (a.delta) st_stream4(a.delta + (size_t)b * P + p, make_float4(d[0], d[1], d[2], d[3]));
            if (a.out) {
              float o[4] = {d[0], d[1], d[2], d[3]};
              if (need_x) {
                o[0] = __fadd_rn(xv[i].x, d[0]);
                o[1] = __fadd_rn(xv[i].y, d[1]);
                o[2] = __fadd_rn(xv[i].z, d[2]);
                o[3] = __fadd_rn(xv[i].w, d[3]);
              }
              if (a.flags & ADIL_SYNTH_CLAMP01) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = fminf(fmaxf(o[j], 0.0f), 1.0f);
              }
              if (a.cc.use) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = div_by_const(__fsub_rn(o[j], mean[j]), stdv[j], rstd[j]);
              }
              st_stream4(a.out + (size_t)b * P + p, make_float4(o[0], o[1], o[2], o[3]));
            }
          }
        }
      }
    }
  };

  if (warp == TC_WARPS) {
    // ===== issuer warp: TMA prefetch of the next raw tile + the 3 x ksteps MMAs of the current one =====
    const uint64_t dhi = make_desc(smem_u32(Dhi), a.Sd, 128), dlo = make_desc(smem_u32(Dlo), a.Sd, 128);
    const uint64_t vhi = make_desc(smem_u32(Vhi), a.Sv, 128), vlo = make_desc(smem_u32(Vlo), a.Sv, 128);
    const uint64_t astep = (uint64_t)((2 * a.Sd) >> 4), bstep = (uint64_t)((2 * a.Sv) >> 4);
    int it = 0;
    for (int tile = my_first; tile < ntiles; tile += gridDim.x, ++it) {
      bar_sync(1, TC_BLOCK);                                       // workers staged tile `it`
      tc_fence_after();
      if (lane == 0) {
        const int nxt = tile + gridDim.x;
        if (nxt < ntiles) {                                         // raw buffer is consumed: prefetch the next tile
          const int nrows = min(TC_TP, P - nxt * TC_TP);
          mbar_expect_tx(bar_full, (uint32_t)(nrows * K * 4));
          bulk_g2s(raw, a.D2 + (size_t)nxt * TC_TP * K, (uint32_t)(nrows * K * 4), bar_full);
        }
        const uint32_t acc = tmem_base + (uint32_t)((it & 1) * a.Np);
        // A (D image, K-major): LBO = Sd between the two 16-byte K chunks, SBO = 128 between 8-pixel groups
        // B (code image, K-major): LBO = Sv, SBO = 128 between 8-image groups
        for (int pass = 0; pass < 3; ++pass) {
          uint64_t ad = (pass == 0) ? dlo : dhi;                    // lo*hi, hi*lo, hi*hi
          uint64_t bd = (pass == 1) ? vlo : vhi;
          for (int ks = 0; ks < ksteps; ++ks, ad += astep, bd += bstep) mma_tf32(acc, ad, bd, idesc, (pass | ks) ? 1u : 0u);
        }
        mma_commit(bar_mma);
      }
      __syncwarp();
    }
  } else {
    // ===== worker warps: split/scatter of the raw tile, epilogue of the previous tile =====
    int it = 0, prev_tile = -1;
    for (int tile = my_first; tile < ntiles; tile += gridDim.x, ++it) {
      const int rows = min(TC_TP, P - tile * TC_TP);
      mbar_wait(bar_full, it & 1);                                 // raw tile `it` landed
      if (it > 0) mbar_wait(bar_mma, (it - 1) & 1);                // MMAs(it-1) retired: images reusable, acc ready
      tc_fence_after();
      for (int p = warp; p < TC_TP; p += TC_WARPS) {                // a warp takes one pixel row, lanes run over atoms
        const float* rrow = raw + p * K;
        const int obase = (p >> 3) * 32 + (p & 7) * 4;
        for (int k = lane; k < K; k += 32) {
          const float val = (p < rows) ? rrow[k] : 0.0f;
          float hi, lo;
          split_tf32(val, hi, lo);
          const int o = (k >> 2) * (a.Sd >> 2) + obase + (k & 3);
          Dhi[o] = hi;
          Dlo[o] = lo;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      bar_arrive(1, TC_BLOCK);                                     // hand the tile to the issuer warp, keep going
      if (prev_tile >= 0) epilogue(prev_tile, it - 1);             // overlaps the MMAs being issued
      prev_tile = tile;
    }
    if (prev_tile >= 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      epilogue(prev_tile, it - 1);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// =========================================================================================================
// backward contractions (+ optional fused AdamW/clamp), bf16x3 on tcgen05.mma.kind::f16:
//   dD[p, k] = sum_b gx[b,p] v[b,k]   M = 128 pixels, N = atoms, K = images  (A = g image MN-major, B = code image MN-major)
//   dv[b, k] = sum_p gx[b,p] D[p,k]   M = images (<=128), N = atoms, K = pixels (A = g image K-major, B = D image MN-major)
// The gradient tile is needed with the contraction along its contiguous dimension (dv) AND across it (dD).  TF32
// operands cannot do that from one image (MN-major TF32 exists only in the 128B_BASE32B swizzle, which has no
// K-major twin -- measured: scripts/umma_probe*.cu), bf16 operands can: an 8x16-byte core matrix is the same bytes
// for a K-major and an MN-major operand.  So every fp32 value is split exactly into three bf16 terms
// x = b0 + b1 + b2 (+ O(2^-23 x)) and each contraction is six MMAs: (2,0) (0,2) (1,1) (1,0) (0,1) (0,0); the two
// dropped cross terms are O(2^-21).  bf16 MMAs run at twice the TF32 rate, so this costs the same tensor time as
// 3xTF32 while the images take 6 instead of 8 bytes per element.
// dD accumulators are double-buffered in TMEM and drained by the AdamW epilogue while the next tile's MMAs run;
// the dv accumulator stays in TMEM across all of the CTA's tiles.
// =========================================================================================================
typedef unsigned short bf16_t;

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

// canonical image offset in bf16 elements: r = "8-row" index, c = contiguous index, S = byte stride between 8-c groups
__device__ __forceinline__ int img16_off(int r, int c, int S_bytes) {
  return (c >> 3) * (S_bytes >> 1) + (r >> 3) * 64 + (r & 7) * 8 + (c & 7);
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct GradTcArgs {
  float* dD2;
  float* D2w;
  float* m;
  float* s;
  float* partial;
  const float* g;
  const float* D2;
  const float* v;
  const int64_t* vidx;
  int B, P, K;
  int Bp16;     // contraction length of dD: round_up(B, 16)
  int Kp16;     // N of both MMAs: round_up(K, 16)
  int Sg, Sd, Sv;
  int gimg, dimg, vimg;  // image sizes in bf16 elements (each matrix has three images)
  int Ks;                // row stride (floats) of the dD staging tile
  unsigned kdiv_mul;     // ceil(2^32 / K): e / K == umulhi(e, kdiv_mul) for e < 2^20
  uint32_t tmem_cols;
  int want_dD, want_dv, atoms_mode;
  ChannelConsts cc;
  AdamwDev hp;
};

constexpr int G_MAXQ = 9;    // image rows per worker warp per tile: 9 x 15 >= 128
constexpr int D_ROWS = (TC_TP + TC_WARPS - 1) / TC_WARPS;  // 9 pixel rows per worker warp per tile
constexpr int EP_BATCH = 2;  // float4 per thread per epilogue batch
constexpr int EP_HALVES = 3; // 3 x 2 x 480 float4 >= 128 x 90 elements

template <int D_KJ>  // atoms per lane when staging the dictionary tile: K <= 32 * D_KJ
__global__ void __launch_bounds__(TC_BLOCK, 1) grad_tc_kernel(const GradTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem_raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
  bf16_t* Vi = reinterpret_cast<bf16_t*>(smem_raw + 128);  // three code images
  bf16_t* Di = Vi + 3 * a.vimg;                             // three dictionary images
  bf16_t* Gi = Di + 3 * a.dimg;                             // three gradient images
  float* dDs = reinterpret_cast<float*>(Gi + 3 * a.gimg);   // [128][Ks] dD tile in global layout (epilogue)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, P = a.P, B = a.B;
  const int ntiles = (P + TC_TP - 1) / TC_TP;

  if (tid == 0) {
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, a.tmem_cols);
  // zero all images once: contraction padding (images in [B, Bp16)) must stay zero, over-read regions stay finite
  {
    uint32_t* z = reinterpret_cast<uint32_t*>(Vi);
    const int nz = (3 * (a.vimg + a.dimg + a.gimg)) >> 1;
    for (int e = tid; e < nz; e += TC_BLOCK) z[e] = 0u;
  }
  __syncthreads();
  if (a.want_dD) {
    for (int e = tid; e < B * K; e += TC_BLOCK) {
      const int b = e / K, k = e - b * K;
      const int64_t row = a.vidx ? a.vidx[b] : (int64_t)b;
      uint32_t w0, w1, w2;
      split_bf16x3(a.v[row * K + k], w0, w1, w2);
      const int o = img16_off(b, k, a.Sv);
      Vi[o] = (bf16_t)(w0 >> 16);
      Vi[a.vimg + o] = (bf16_t)(w1 >> 16);
      Vi[2 * a.vimg + o] = (bf16_t)(w2 >> 16);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_dv = tmem_base + (uint32_t)(2 * a.Kp16);

  const uint32_t idesc_dD = make_idesc_bf16(128, a.Kp16, true, true);
  const uint32_t idesc_dv = make_idesc_bf16(128, a.Kp16, false, true);
  const int quad = warp & 3, cgrp = warp >> 2;
  const int ncg = (quad == 3) ? 3 : 4;  // quadrant 3 lost warp 15 to the issuer role
  const int nchunks = a.Kp16 / 8;

  float4 greg[G_MAXQ];
  float dreg[D_ROWS][D_KJ];
  // fixed per-thread roles: gradient tile -> image rows b = warp + 16*i, 4-pixel column pq = lane;
  //                         dictionary tile -> pixel rows p = warp + 16*i, atoms k = lane + 32*j
  auto prefetch = [&](int tile) {
    const int p0 = tile * TC_TP;
    const int p = p0 + 4 * lane;
#pragma unroll
    for (int i = 0; i < G_MAXQ; ++i) {
      const int b = warp + i * TC_WARPS;
      greg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B && p < P) greg[i] = ld_stream4(a.g + (size_t)b * P + p);
    }
    if (a.want_dv) {
#pragma unroll
      for (int i = 0; i < D_ROWS; ++i) {
        const int pl = warp + i * TC_WARPS, pr = p0 + pl;
#pragma unroll
        for (int j = 0; j < D_KJ; ++j) {
          const int k = lane + 32 * j;
          dreg[i][j] = (pl < TC_TP && pr < P && k < K) ? __ldg(a.D2 + (size_t)pr * K + k) : 0.0f;
        }
      }
    }
  };

  auto stage = [&](int tile) {
    const int p = tile * TC_TP + 4 * lane;
    float stdv[4], rstd[4];
    if (a.cc.use) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = min((p + j) / a.cc.hw, kMaxC - 1);
        stdv[j] = a.cc.stdv[c]; rstd[j] = a.cc.rstd[c];
      }
    }
    const int gcol = (lane >> 1) * (a.Sg >> 1) + (lane & 1) * 4;   // img16_off(b, 4*lane) without the row part
#pragma unroll
    for (int i = 0; i < G_MAXQ; ++i) {
      const int b = warp + i * TC_WARPS;
      if (b < B) {
        float val[4] = {greg[i].x, greg[i].y, greg[i].z, greg[i].w};
        if (a.cc.use) {
#pragma unroll
          for (int j = 0; j < 4; ++j) val[j] = div_by_const(val[j], stdv[j], rstd[j]);
        }
        uint32_t w0[4], w1[4], w2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_bf16x3(val[j], w0[j], w1[j], w2[j]);
        const int o = gcol + (b >> 3) * 64 + (b & 7) * 8;           // 8-byte slot (b, 4*lane .. 4*lane+3)
        *reinterpret_cast<uint2*>(Gi + o) = make_uint2(pack_hi16(w0[0], w0[1]), pack_hi16(w0[2], w0[3]));
        *reinterpret_cast<uint2*>(Gi + a.gimg + o) = make_uint2(pack_hi16(w1[0], w1[1]), pack_hi16(w1[2], w1[3]));
        *reinterpret_cast<uint2*>(Gi + 2 * a.gimg + o) = make_uint2(pack_hi16(w2[0], w2[1]), pack_hi16(w2[2], w2[3]));
      }
    }
    if (a.want_dv) {
#pragma unroll
      for (int i = 0; i < D_ROWS; ++i) {
        const int pr = warp + i * TC_WARPS;                          // pixel row inside the tile
        const int obase = (pr >> 3) * 64 + (pr & 7) * 8;
#pragma unroll
        for (int j = 0; j < D_KJ; ++j) {
          const int k = lane + 32 * j;
          if (k < K && pr < TC_TP) {
            uint32_t w0, w1, w2;
            split_bf16x3(dreg[i][j], w0, w1, w2);
            const int o = (k >> 3) * (a.Sd >> 1) + obase + (k & 7);
            Di[o] = (bf16_t)(w0 >> 16);
            Di[a.dimg + o] = (bf16_t)(w1 >> 16);
            Di[2 * a.dimg + o] = (bf16_t)(w2 >> 16);
          }
        }
      }
    }
  };

  auto issue = [&](int it) {
    const uint32_t gb = smem_u32(Gi), vb = smem_u32(Vi), db = smem_u32(Di);
    const uint32_t gsz = 2u * a.gimg, vsz = 2u * a.vimg, dsz = 2u * a.dimg;  // image sizes in bytes
    // term order: smallest contributions first
    const int ta[6] = {2, 0, 1, 1, 0, 0};
    const int tb[6] = {0, 2, 1, 0, 1, 0};
    if (a.want_dD) {
      // A = g image as [M=pixel, K=image] MN-major: SBO = Sg between 8-pixel groups, LBO = 128 between 8-image groups
      // B = code image as [N=atom, K=image] MN-major: SBO = Sv between 8-atom groups, LBO = 128
      const uint32_t acc = tmem_base + (uint32_t)((it & 1) * a.Kp16);
      const int ksteps = a.Bp16 / 16;
#pragma unroll 1
      for (int t = 0; t < 6; ++t) {
        uint64_t ad = make_desc(gb + ta[t] * gsz, 128, a.Sg);
        uint64_t bd = make_desc(vb + tb[t] * vsz, 128, a.Sv);
        for (int ks = 0; ks < ksteps; ++ks, ad += 16, bd += 16) mma_bf16(acc, ad, bd, idesc_dD, (t | ks) ? 1u : 0u);
      }
    }
    if (a.want_dv) {
      // A = g image as [M=image, K=pixel] K-major: LBO = Sg between the two 8-pixel chunks, SBO = 128 (8-image groups)
      // B = D image as [N=atom, K=pixel] MN-major: SBO = Sd between 8-atom groups, LBO = 128 between 8-pixel groups
      const uint64_t astep = (uint64_t)((2 * a.Sg) >> 4);
#pragma unroll 1
      for (int t = 0; t < 6; ++t) {
        uint64_t ad = make_desc(gb + ta[t] * gsz, a.Sg, 128);
        uint64_t bd = make_desc(db + tb[t] * dsz, 128, a.Sd);
        for (int ks = 0; ks < TC_TP / 16; ++ks, ad += astep, bd += 16)
          mma_bf16(acc_dv, ad, bd, idesc_dv, (it | t | ks) ? 1u : 0u);
      }
    }
    mma_commit(bar_mma);
  };

  auto epilogue = [&](int tile, int it) {
    // Phase A: accumulator (row = pixel = TMEM lane, column = atom) -> shared memory in the tile's global layout
    // [pixel][atom] (row stride Ks chosen so the 16-byte stores of a quarter-warp hit distinct banks).
    const uint32_t acc = tmem_base + (uint32_t)((it & 1) * a.Kp16) + ((uint32_t)(quad * 32) << 16);
    const int Ks = a.Ks;
    float* row = dDs + (quad * 32 + lane) * Ks;
    bar_sync(2, TC_THREADS);  // every worker is done reading the staging tile of the previous epilogue
    for (int ch = cgrp; ch < nchunks; ch += ncg) {  // warp-uniform
      const int k0 = ch * 8;
      if (k0 >= K) break;
      float r[8];
      tmem_ld8(acc + (uint32_t)k0, r);
      *reinterpret_cast<float4*>(row + k0) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(row + k0 + 4) = make_float4(r[4], r[5], r[6], r[7]);
    }
    tc_fence_before();
    bar_sync(2, TC_THREADS);
    // Phase B: one coalesced 128-bit pass over the tile's contiguous [rows x K] block of D / m / s (or dD)
    const int rows = min(TC_TP, P - tile * TC_TP);
    const int n4 = (rows * K) >> 2;
    const size_t base = (size_t)tile * TC_TP * K;
#pragma unroll 1
    for (int half = 0; half < EP_HALVES; ++half) {
      float4 Dv[EP_BATCH], Mv[EP_BATCH], Sv[EP_BATCH];
      if (a.D2w != nullptr) {
#pragma unroll
        for (int i = 0; i < EP_BATCH; ++i) {
          const int e4 = tid + (half * EP_BATCH + i) * TC_THREADS;
          if (e4 < n4) {
            Dv[i] = *reinterpret_cast<const float4*>(a.D2 + base + 4 * (size_t)e4);
            Mv[i] = *reinterpret_cast<const float4*>(a.m + base + 4 * (size_t)e4);
            Sv[i] = *reinterpret_cast<const float4*>(a.s + base + 4 * (size_t)e4);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < EP_BATCH; ++i) {
        const int e4 = tid + (half * EP_BATCH + i) * TC_THREADS;
        if (e4 < n4) {
          const int e = 4 * e4;
          int pp = a.kdiv_mul ? (int)__umulhi((unsigned)e, a.kdiv_mul) : e;      // e / K (exact for e < 2^20; 0: K == 1)
          int kk = e - pp * K;
          float gd[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            gd[j] = dDs[pp * Ks + kk];
            if (++kk == K) { kk = 0; ++pp; }
          }
          if (a.D2w != nullptr) {
            adamw_update(Dv[i].x, Mv[i].x, Sv[i].x, gd[0], a.hp);
            adamw_update(Dv[i].y, Mv[i].y, Sv[i].y, gd[1], a.hp);
            adamw_update(Dv[i].z, Mv[i].z, Sv[i].z, gd[2], a.hp);
            adamw_update(Dv[i].w, Mv[i].w, Sv[i].w, gd[3], a.hp);
            if (a.atoms_mode == ADIL_ATOMS_CLAMP1) {
              Dv[i].x = clamp1(Dv[i].x); Dv[i].y = clamp1(Dv[i].y); Dv[i].z = clamp1(Dv[i].z); Dv[i].w = clamp1(Dv[i].w);
            }
            *reinterpret_cast<float4*>(a.D2w + base + 4 * (size_t)e4) = Dv[i];
            *reinterpret_cast<float4*>(a.m + base + 4 * (size_t)e4) = Mv[i];
            *reinterpret_cast<float4*>(a.s + base + 4 * (size_t)e4) = Sv[i];
          } else {
            *reinterpret_cast<float4*>(a.dD2 + base + 4 * (size_t)e4) = make_float4(gd[0], gd[1], gd[2], gd[3]);
          }
        }
      }
    }
  };

  if (warp == TC_WARPS) {
    // ===== issuer warp: waits for "tile staged", issues the 6-term MMAs of both contractions, commits =====
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      bar_sync(1, TC_BLOCK);
      tc_fence_after();
      if (lane == 0) issue(it);
      __syncwarp();
    }
  } else {
    int it = 0, prev_tile = -1;
    int tile = blockIdx.x;
    if (tile < ntiles) prefetch(tile);
    for (; tile < ntiles; tile += gridDim.x, ++it) {
      if (it > 0) {
        mbar_wait(bar_mma, (it - 1) & 1);  // MMAs(it-1) retired: images free, dD accumulator (it-1) complete
        tc_fence_after();
      }
      stage(tile);
      fence_proxy_async();
      tc_fence_before();
      bar_arrive(1, TC_BLOCK);                            // hand the tile to the issuer warp, keep going
      const int nxt = tile + gridDim.x;
      if (nxt < ntiles) prefetch(nxt);                    // global loads in flight while the tensor core works
      if (a.want_dD && prev_tile >= 0) epilogue(prev_tile, it - 1);
      prev_tile = tile;
    }
    if (prev_tile >= 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (a.want_dD) epilogue(prev_tile, it - 1);
      if (a.want_dv) {
        // dv accumulator: row = image (TMEM lane), column = atom -> this CTA's slab of the partial buffer
        const int b = quad * 32 + lane;
        float* dst = a.partial + (size_t)blockIdx.x * B * K + (size_t)b * K;
        for (int ch = cgrp; ch < nchunks; ch += ncg) {
          const int k0 = ch * 8;
          if (k0 >= K) break;
          float r[8];
          tmem_ld8(acc_dv + ((uint32_t)(quad * 32) << 16) + (uint32_t)k0, r);
          if (b < B) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (k0 + j < K) dst[k0 + j] = r[j];
          }
        }
        tc_fence_before();
      }
    } else if (a.want_dv) {
      for (int e = tid; e < B * K; e += TC_THREADS) a.partial[(size_t)blockIdx.x * B * K + e] = 0.0f;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

uint32_t pow2_cols(int need) {
  uint32_t c = 32;
  while ((int)c < need) c <<= 1;
  return c;
}

struct SynthPlan {
  int Np, Kp8, Sd, Sv, raw_floats, dimg_floats, vimg_floats;
  size_t smem;
  uint32_t tmem_cols;
  bool ok;
};

SynthPlan plan_synth(int B, int P, int K) {
  SynthPlan pl{};
  pl.ok = false;
  if (B < 1 || B > 256 || K < 1 || P % 4 != 0) return pl;
  pl.Np = rup(B, 16);
  pl.Kp8 = rup(K, 8);
  pl.Sd = img_stride(TC_TP);
  pl.Sv = img_stride(pl.Np);
  pl.raw_floats = rup(TC_TP * K, 32);
  pl.dimg_floats = (pl.Kp8 / 4) * (pl.Sd / 4) + 64;
  pl.vimg_floats = (pl.Kp8 / 4) * (pl.Sv / 4) + 64;
  pl.smem = 128 + 2048 + sizeof(float) * ((size_t)pl.raw_floats + 2 * (size_t)pl.dimg_floats + 2 * (size_t)pl.vimg_floats +
                                          (size_t)pl.Np * S_OS);
  pl.tmem_cols = pow2_cols(2 * pl.Np);
  pl.ok = pl.smem <= SMEM_LIMIT && pl.tmem_cols <= 512 && pl.Np <= 256;
  return pl;
}

struct GradPlan {
  int Bp16, Kp16, Sg, Sd, Sv, gimg, dimg, vimg, Ks;
  size_t smem;
  uint32_t tmem_cols;
  bool ok;
};

GradPlan plan_grad(int B, int P, int K) {
  GradPlan pl{};
  pl.ok = false;
  if (B < 1 || B > 128 || K < 1 || P % 4 != 0) return pl;
  pl.Bp16 = rup(B, 16);
  pl.Kp16 = rup(K, 16);
  pl.Sg = img_stride(pl.Bp16);
  pl.Sd = img_stride(TC_TP);
  pl.Sv = img_stride(pl.Bp16);
  const int kg = (K + 7) / 8;                             // 8-atom groups actually written
  pl.vimg = kg * (pl.Sv / 2) + 64;                        // bf16 elements per image
  pl.dimg = kg * (pl.Sd / 2) + 64;
  pl.gimg = (TC_TP / 8) * (pl.Sg / 2) + 1024 + 64;        // +2 KB: M = 128 image rows are read even when Bp16 < 128
  pl.Ks = rup(K, 8) + 4;                                  // Ks % 8 == 4: conflict-free 16-byte row stores
  pl.smem = 128 + 2 * 3 * ((size_t)pl.vimg + pl.dimg + pl.gimg) + sizeof(float) * TC_TP * (size_t)pl.Ks;
  pl.tmem_cols = pow2_cols(3 * pl.Kp16);
  // the MMAs read N = Kp16 atoms: the code / dictionary images over-read into the buffers that follow them
  // (codes -> dictionary -> gradient images), which must be large enough to absorb it
  const size_t over_v = (size_t)(pl.Kp16 / 8) * pl.Sv, over_d = (size_t)(pl.Kp16 / 8) * pl.Sd;
  const size_t tail_after_v = 2 * (3 * (size_t)pl.dimg + 3 * (size_t)pl.gimg);
  const size_t tail_after_d = 2 * 3 * (size_t)pl.gimg;
  pl.ok = pl.smem <= SMEM_LIMIT && pl.tmem_cols <= 512 && over_v <= 2 * (size_t)pl.vimg + tail_after_v &&
          over_d <= 2 * (size_t)pl.dimg + tail_after_d && B <= G_MAXQ * TC_WARPS && K <= 96 &&
          TC_TP * K <= 4 * EP_BATCH * EP_HALVES * TC_THREADS;
  return pl;
}

}  // namespace

bool tc_synth_ok(int B, int P, int K) { return plan_synth(B, P, K).ok; }
bool tc_grad_ok(int B, int P, int K) { return plan_grad(B, P, K).ok; }
bool tc_shape_ok(int B, int P, int K) { return tc_synth_ok(B, P, K) && tc_grad_ok(B, P, K); }

int launch_synth_tc(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2,
                    const float* v, const int64_t* v_index, int B, int P, int K, const ChannelConsts& cc, float eps,
                    int flags, cudaStream_t st) {
  const SynthPlan pl = plan_synth(B, P, K);
  if (!pl.ok) return set_error(-4, "adil_synth: shape B=%d P=%d K=%d does not qualify for the tcgen05 path", B, P, K);
  SynthTcArgs a;
  a.out = out; a.delta = delta_out; a.x = x; a.xidx = x_index; a.D2 = D2; a.v = v; a.vidx = v_index;
  a.B = B; a.P = P; a.K = K; a.Np = pl.Np; a.Kp8 = pl.Kp8; a.Sd = pl.Sd; a.Sv = pl.Sv;
  a.raw_floats = pl.raw_floats; a.dimg_floats = pl.dimg_floats; a.vimg_floats = pl.vimg_floats;
  a.tmem_cols = pl.tmem_cols; a.eps = eps; a.flags = flags; a.cc = cc;
  a.cc.use = (flags & ADIL_SYNTH_NORMALIZE) ? 1 : 0;
  int rc = check_cuda(cudaFuncSetAttribute(synth_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem),
                      "cudaFuncSetAttribute(synth_tc)");
  if (rc) return rc;
  const int ntiles = (P + TC_TP - 1) / TC_TP;
  int grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  synth_tc_kernel<<<grid, TC_BLOCK, pl.smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "synth_tc_kernel launch");
}

int launch_grad_tc(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                   const float* v, const int64_t* v_index, int B, int P, int K, const ChannelConsts& cc,
                   const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes, cudaStream_t st) {
  const GradPlan pl = plan_grad(B, P, K);
  if (!pl.ok) return set_error(-4, "adil_grad: shape B=%d P=%d K=%d does not qualify for the tcgen05 path", B, P, K);
  GradTcArgs a;
  a.dD2 = dD2; a.D2w = D2_rw; a.m = m; a.s = s; a.partial = scratch; a.g = g; a.D2 = D2; a.v = v; a.vidx = v_index;
  a.B = B; a.P = P; a.K = K; a.Bp16 = pl.Bp16; a.Kp16 = pl.Kp16; a.Sg = pl.Sg; a.Sd = pl.Sd; a.Sv = pl.Sv;
  a.gimg = pl.gimg; a.dimg = pl.dimg; a.vimg = pl.vimg; a.Ks = pl.Ks;
  a.kdiv_mul = K == 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned long long)K - 1) / (unsigned long long)K);
  a.tmem_cols = pl.tmem_cols;
  a.want_dD = (dD2 != nullptr || D2_rw != nullptr) ? 1 : 0;
  a.want_dv = (dvb != nullptr) ? 1 : 0;
  a.atoms_mode = atoms_mode; a.cc = cc;
  if (hp) a.hp = *hp;
  if (!a.want_dD && !a.want_dv) return 0;
  const int ntiles = (P + TC_TP - 1) / TC_TP;
  int grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  if (grid > kMaxGradCtas) grid = kMaxGradCtas;
  if (a.want_dv) {
    const size_t need = (size_t)grid * B * K * sizeof(float);
    if (scratch == nullptr || scratch_bytes < need)
      return set_error(-2, "adil_grad: scratch too small (%zu < %zu bytes)", scratch_bytes, need);
  }
  auto kern = K <= 32 ? grad_tc_kernel<1> : (K <= 64 ? grad_tc_kernel<2> : grad_tc_kernel<3>);
  int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem),
                      "cudaFuncSetAttribute(grad_tc)");
  if (rc) return rc;
  kern<<<grid, TC_BLOCK, pl.smem, st>>>(a);
  rc = check_cuda(cudaGetLastError(), "grad_tc_kernel launch");
  if (rc) return rc;
  if (a.want_dv) return launch_reduce_partials(dvb, scratch, B * K, grid, st);
  return 0;
}

}  // namespace adil
