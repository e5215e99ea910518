// CUDA-core (FFMA) kernels of the ADiL hot path: perturbation synthesis and the fused backward
// contractions.  These serve every shape (any B, K <= 256) and are the path for small atom counts, where the
// contraction is far below the fp32 ridge and the kernels are HBM-bound; adil_tc.cu holds the tcgen05
// split-TF32 kernels used when K is large enough for the FFMA pipe to become the limiter.
//
// Reference semantics: adil.py:24-27 (synthesis), demo_dL_attack.py:22-25 (Normalize), autograd backward of
// both (adil.py:185), torch.optim.AdamW + update_d (adil.py:186,188).
#include "adil_common.cuh"

namespace adil {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------------------
// synthesis:  out[b,p] = f(x[b,p] + sum_k v[b,k] * D2[p,k])
// One CTA owns a 128-pixel tile of D (read from HBM exactly once), transposed into shared memory so that a
// lane's four consecutive pixels are one conflict-free LDS.128; a warp produces 8 images x 128 pixels per
// pass and writes 512-byte coalesced rows with 128-bit stores.
// ------------------------------------------------------------------------------------------------------
constexpr int S_TP = 128;   // pixels per tile
constexpr int S_TPS = S_TP;  // row stride of the transposed tile (16-byte chunks are XOR-swizzled by atom index)
constexpr int S_TB = 8;      // images per warp pass

// Dt[k][p] lives at k*S_TPS + ((p>>2) ^ (k&7))*4 + (p&3): a lane's 4 consecutive pixels stay one 16-byte chunk
// (conflict-free LDS.128 in the main loop) and the transposing fill below, whose warp covers 4 pixels x 8 atoms
// per store, touches 32 distinct banks.
__device__ __forceinline__ int dt_offset(int k, int p) { return k * S_TPS + ((((p >> 2) ^ (k & 7))) << 2) + (p & 3); }

struct SynthArgs {
  float* out;
  float* delta;
  const float* x;
  const int64_t* xidx;
  const float* D2;
  const float* v;
  const int64_t* vidx;
  int B, P, K, Kp, bch;
  float eps;
  int flags;
  ChannelConsts cc;
};

__global__ void __launch_bounds__(kThreads) synth_fma_kernel(const SynthArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Dt = smem;                 // [Kp][S_TPS]  Dt[k][p]
  float* vs = smem + a.Kp * S_TPS;  // [bch][Kp]
  long long* xoff_s = reinterpret_cast<long long*>(vs + a.bch * a.Kp);  // [bch] element offset of each image's x row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.K, Kp = a.Kp, P = a.P, B = a.B;
  const int ntiles = (P + S_TP - 1) / S_TP;

  // padded atom rows of Dt are zero for the whole kernel
  for (int e = tid; e < (Kp - K) * S_TPS; e += kThreads) Dt[K * S_TPS + e] = 0.0f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * S_TP;
    const int rows = min(S_TP, P - p0);
    __syncthreads();  // previous tile fully consumed
    {
      // transposing fill: a warp step loads 4 pixel rows x 8 consecutive atoms (four 32-byte runs) and scatters
      // them bank-conflict-free into the swizzled Dt
      const float* src = a.D2 + (size_t)p0 * K;
      const int pl = lane >> 3, kl = lane & 7;
      const int kblocks = (K + 7) >> 3;
      for (int it = warp; it < (S_TP / 4) * kblocks; it += kWarps) {
        const int pq = it / kblocks, kb = it - pq * kblocks;
        const int p = 4 * pq + pl, k = 8 * kb + kl;
        if (k < K) Dt[dt_offset(k, p)] = (p < rows) ? src[(size_t)p * K + k] : 0.0f;
      }
    }
    for (int b0 = 0; b0 < B; b0 += a.bch) {
      const int nb = min(a.bch, B - b0);
      const int nbp = round_up(nb, S_TB);
      __syncthreads();  // Dt filled / previous chunk consumed
      for (int e = tid; e < nbp * Kp; e += kThreads) {
        int r = e / Kp, k = e - r * Kp;
        float val = 0.0f;
        if (r < nb && k < K) {
          int64_t row = a.vidx ? a.vidx[b0 + r] : (int64_t)(b0 + r);
          val = a.v[row * K + k];
        }
        vs[e] = val;
      }
      // x row offsets via shared memory: a dependent global index load per image would serialise the prefetch below
      for (int r = tid; r < nb; r += kThreads)
        xoff_s[r] = (a.xidx ? (long long)a.xidx[b0 + r] : (long long)(b0 + r)) * (long long)P;
      __syncthreads();
      const int ngroups = nbp / S_TB;
      const int p = p0 + 4 * lane;
      for (int gi = warp; gi < ngroups; gi += kWarps) {
        const int bbase = b0 + gi * S_TB;
        float4 xv[S_TB];
        if (a.x != nullptr && a.out != nullptr && p < P) {
#pragma unroll
          for (int ib = 0; ib < S_TB; ++ib) {
            int b = bbase + ib;
            if (b < B) xv[ib] = ld_stream4(a.x + xoff_s[b - b0] + p);
          }
        }
        float acc[S_TB][4];
#pragma unroll
        for (int ib = 0; ib < S_TB; ++ib)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[ib][j] = 0.0f;
        const float* vrow = vs + gi * S_TB * Kp;
        for (int k4 = 0; k4 < Kp; k4 += 4) {
          const float4 d0 = *reinterpret_cast<const float4*>(Dt + dt_offset(k4 + 0, 4 * lane));
          const float4 d1 = *reinterpret_cast<const float4*>(Dt + dt_offset(k4 + 1, 4 * lane));
          const float4 d2 = *reinterpret_cast<const float4*>(Dt + dt_offset(k4 + 2, 4 * lane));
          const float4 d3 = *reinterpret_cast<const float4*>(Dt + dt_offset(k4 + 3, 4 * lane));
#pragma unroll
          for (int ib = 0; ib < S_TB; ++ib) {
            const float4 vv = *reinterpret_cast<const float4*>(vrow + ib * Kp + k4);
            acc[ib][0] = fmaf(vv.x, d0.x, acc[ib][0]);
            acc[ib][1] = fmaf(vv.x, d0.y, acc[ib][1]);
            acc[ib][2] = fmaf(vv.x, d0.z, acc[ib][2]);
            acc[ib][3] = fmaf(vv.x, d0.w, acc[ib][3]);
            acc[ib][0] = fmaf(vv.y, d1.x, acc[ib][0]);
            acc[ib][1] = fmaf(vv.y, d1.y, acc[ib][1]);
            acc[ib][2] = fmaf(vv.y, d1.z, acc[ib][2]);
            acc[ib][3] = fmaf(vv.y, d1.w, acc[ib][3]);
            acc[ib][0] = fmaf(vv.z, d2.x, acc[ib][0]);
            acc[ib][1] = fmaf(vv.z, d2.y, acc[ib][1]);
            acc[ib][2] = fmaf(vv.z, d2.z, acc[ib][2]);
            acc[ib][3] = fmaf(vv.z, d2.w, acc[ib][3]);
            acc[ib][0] = fmaf(vv.w, d3.x, acc[ib][0]);
            acc[ib][1] = fmaf(vv.w, d3.y, acc[ib][1]);
            acc[ib][2] = fmaf(vv.w, d3.z, acc[ib][2]);
            acc[ib][3] = fmaf(vv.w, d3.w, acc[ib][3]);
          }
        }
        if (p < P) {
          int cidx[4];
          if (a.cc.use) {
#pragma unroll
            for (int j = 0; j < 4; ++j) cidx[j] = (p + j) / a.cc.hw;
          }
#pragma unroll
          for (int ib = 0; ib < S_TB; ++ib) {
            const int b = bbase + ib;
            if (b >= B) break;
            float d[4] = {acc[ib][0], acc[ib][1], acc[ib][2], acc[ib][3]};
            if (a.flags & ADIL_SYNTH_CLAMP_DELTA) {
#pragma unroll
              for (int j = 0; j < 4; ++j) d[j] = fminf(fmaxf(d[j], -a.eps), a.eps);
            }
            if (a.delta) st_stream4(a.delta + (size_t)b * P + p, make_float4(d[0], d[1], d[2], d[3]));
            if (a.out) {
              float o[4] = {d[0], d[1], d[2], d[3]};
              if (a.x) {
                o[0] = __fadd_rn(xv[ib].x, d[0]);
                o[1] = __fadd_rn(xv[ib].y, d[1]);
                o[2] = __fadd_rn(xv[ib].z, d[2]);
                o[3] = __fadd_rn(xv[ib].w, d[3]);
              }
              if (a.flags & ADIL_SYNTH_CLAMP01) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = fminf(fmaxf(o[j], 0.0f), 1.0f);
              }
              if (a.cc.use) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  o[j] = __fdiv_rn(__fsub_rn(o[j], a.cc.mean[cidx[j]]), a.cc.stdv[cidx[j]]);
              }
              st_stream4(a.out + (size_t)b * P + p, make_float4(o[0], o[1], o[2], o[3]));
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// backward contractions, one pass over g:
//   dD2[p,k] = sum_b gx[b,p] v[b,k]   (exact per pixel tile; optionally consumed in registers by AdamW+clamp)
//   dvb[b,k] = sum_p gx[b,p] D2[p,k]  (accumulated in shared memory over the CTA's tiles, then one [B,K] slab
//                                      per CTA goes to scratch and reduce_partials_kernel sums the slabs in a
//                                      fixed order -- no float atomics, bit-reproducible run to run)
// ------------------------------------------------------------------------------------------------------
constexpr int G_TBV = 8;  // images per thread in the dv phase

struct GradArgs {
  float* dD2;   // unfused output or nullptr
  float* D2w;   // fused: updated in place (same buffer as D2) or nullptr
  float* m;
  float* s;
  float* partial;  // [grid][B][K] or nullptr (no dv wanted)
  const float* g;
  const float* D2;
  const float* v;
  const int64_t* vidx;
  int B, P, K, Kp;
  int want_dD, want_dv, atoms_mode;
  int accumulate;  // unfused output: dD2 += instead of dD2 =
  const float* delta;  // l2 penalty (adil_regularized.py:112-114): gx += l2_coef * delta
  float l2_coef;
  ChannelConsts cc;
  AdamwDev hp;
};

template <int VK>
__device__ __forceinline__ void dict_epilogue(const GradArgs& a, const float (&acc)[4][4], int p0, int pg, int kg,
                                               const float* Ds, int Kp) {
  // acc[j][c]: pixel p0+4pg+j, atom 4kg+c
  const int K = a.K;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = p0 + 4 * pg + j;
    if (p >= a.P) break;
    const size_t rowoff = (size_t)p * K;
#pragma unroll
    for (int c0 = 0; c0 < 4; c0 += VK) {
      const int k = 4 * kg + c0;
      if (k >= K) break;  // K % VK == 0, so a vector is entirely valid or entirely out
      const size_t idx = rowoff + k;
      if (a.D2w == nullptr) {
        if (VK == 4) {
          float4 o = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
          if (a.accumulate) {
            const float4 t = *reinterpret_cast<const float4*>(a.dD2 + idx);
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          *reinterpret_cast<float4*>(a.dD2 + idx) = o;
        } else if (VK == 2) {
          float2 o = make_float2(acc[j][c0], acc[j][c0 + 1]);
          if (a.accumulate) {
            const float2 t = *reinterpret_cast<const float2*>(a.dD2 + idx);
            o.x += t.x; o.y += t.y;
          }
          *reinterpret_cast<float2*>(a.dD2 + idx) = o;
        } else {
          a.dD2[idx] = a.accumulate ? a.dD2[idx] + acc[j][c0] : acc[j][c0];
        }
      } else {
        float dv[VK], mv[VK], sv[VK];
        if (VK == 4) {
          float4 t = *reinterpret_cast<const float4*>(a.m + idx);
          mv[0] = t.x; mv[1] = t.y; mv[2] = t.z; mv[3] = t.w;
          t = *reinterpret_cast<const float4*>(a.s + idx);
          sv[0] = t.x; sv[1] = t.y; sv[2] = t.z; sv[3] = t.w;
        } else if (VK == 2) {
          float2 t = *reinterpret_cast<const float2*>(a.m + idx);
          mv[0] = t.x; mv[1] = t.y;
          t = *reinterpret_cast<const float2*>(a.s + idx);
          sv[0] = t.x; sv[1] = t.y;
        } else {
          mv[0] = a.m[idx];
          sv[0] = a.s[idx];
        }
#pragma unroll
        for (int c = 0; c < VK; ++c) {
          dv[c] = Ds ? Ds[(4 * pg + j) * Kp + k + c] : a.D2[idx + c];
          adamw_update_fast(dv[c], mv[c], sv[c], acc[j][c0 + c], a.hp);
          if (a.atoms_mode == ADIL_ATOMS_CLAMP1) dv[c] = clamp1(dv[c]);
        }
        if (VK == 4) {
          *reinterpret_cast<float4*>(a.D2w + idx) = make_float4(dv[0], dv[1], dv[2], dv[3]);
          *reinterpret_cast<float4*>(a.m + idx) = make_float4(mv[0], mv[1], mv[2], mv[3]);
          *reinterpret_cast<float4*>(a.s + idx) = make_float4(sv[0], sv[1], sv[2], sv[3]);
        } else if (VK == 2) {
          *reinterpret_cast<float2*>(a.D2w + idx) = make_float2(dv[0], dv[1]);
          *reinterpret_cast<float2*>(a.m + idx) = make_float2(mv[0], mv[1]);
          *reinterpret_cast<float2*>(a.s + idx) = make_float2(sv[0], sv[1]);
        } else {
          a.D2w[idx] = dv[0];
          a.m[idx] = mv[0];
          a.s[idx] = sv[0];
        }
      }
    }
  }
}

template <int TP>
__global__ void __launch_bounds__(kThreads) grad_fma_kernel(const GradArgs a) {
  constexpr int TPS = TP + 4;
  extern __shared__ __align__(16) float smem[];
  const int B = a.B, K = a.K, Kp = a.Kp, P = a.P;
  float* gs = smem;                                   // [B][TPS]   gx tile
  float* Ds = gs + B * TPS;                           // [TP][Kp]   (want_dv)
  float* vs = Ds + (a.want_dv ? TP * Kp : 0);         // [B][Kp]    (want_dD)
  float* dvs = vs + (a.want_dD ? B * Kp : 0);         // [B][Kp]    (want_dv)
  const int tid = threadIdx.x;
  const int ntiles = (P + TP - 1) / TP;
  const int KG = Kp / 4;

  if (a.want_dD) {
    for (int e = tid; e < B * Kp; e += kThreads) {
      int r = e / Kp, k = e - r * Kp;
      float val = 0.0f;
      if (k < K) {
        int64_t row = a.vidx ? a.vidx[r] : (int64_t)r;
        val = a.v[row * K + k];
      }
      vs[e] = val;
    }
  }
  if (a.want_dv) {
    for (int e = tid; e < B * Kp; e += kThreads) dvs[e] = 0.0f;
  }

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int p0 = tile * TP;
    const int rows = min(TP, P - p0);
    __syncthreads();
    // gx tile: g / std[c], zero beyond the image end
    for (int e4 = tid; e4 < B * (TP / 4); e4 += kThreads) {
      const int b = e4 / (TP / 4), q = e4 - b * (TP / 4);
      const int p = p0 + 4 * q;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < P) {
        val = ld_stream4(a.g + (size_t)b * P + p);
        if (a.cc.use) {
          val.x = __fdiv_rn(val.x, a.cc.stdv[(p + 0) / a.cc.hw]);
          val.y = __fdiv_rn(val.y, a.cc.stdv[(p + 1) / a.cc.hw]);
          val.z = __fdiv_rn(val.z, a.cc.stdv[(p + 2) / a.cc.hw]);
          val.w = __fdiv_rn(val.w, a.cc.stdv[(p + 3) / a.cc.hw]);
        }
        if (a.delta != nullptr) {  // gradient of 0.5 * l2_coef * ||D v||^2 w.r.t. the perturbation
          const float4 dl = ld_stream4(a.delta + (size_t)b * P + p);
          val.x = fmaf(a.l2_coef, dl.x, val.x);
          val.y = fmaf(a.l2_coef, dl.y, val.y);
          val.z = fmaf(a.l2_coef, dl.z, val.z);
          val.w = fmaf(a.l2_coef, dl.w, val.w);
        }
      }
      *reinterpret_cast<float4*>(gs + b * TPS + 4 * q) = val;
    }
    if (a.want_dv) {
      const float* src = a.D2 + (size_t)p0 * K;
      for (int e = tid; e < TP * Kp; e += kThreads) {
        const int p = e / Kp, k = e - p * Kp;
        Ds[e] = (p < rows && k < K) ? src[p * K + k] : 0.0f;
      }
    }
    __syncthreads();

    if (a.want_dv) {
      // thread tile: 8 images (strided by BGn so neighbouring lanes hit different banks) x 4 atoms
      const int BGn = (B + G_TBV - 1) / G_TBV;
      for (int t = tid; t < BGn * KG; t += kThreads) {
        const int bg = t / KG, kg = t - bg * KG;
        int brow[G_TBV];
#pragma unroll
        for (int i = 0; i < G_TBV; ++i) brow[i] = min(bg + i * BGn, B - 1);
        float acc[G_TBV][4];
#pragma unroll
        for (int i = 0; i < G_TBV; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[i][c] = 0.0f;
        for (int p = 0; p < TP; p += 4) {
          const float4 d0 = *reinterpret_cast<const float4*>(Ds + (p + 0) * Kp + 4 * kg);
          const float4 d1 = *reinterpret_cast<const float4*>(Ds + (p + 1) * Kp + 4 * kg);
          const float4 d2 = *reinterpret_cast<const float4*>(Ds + (p + 2) * Kp + 4 * kg);
          const float4 d3 = *reinterpret_cast<const float4*>(Ds + (p + 3) * Kp + 4 * kg);
#pragma unroll
          for (int i = 0; i < G_TBV; ++i) {
            const float4 gv = *reinterpret_cast<const float4*>(gs + brow[i] * TPS + p);
            acc[i][0] = fmaf(gv.x, d0.x, acc[i][0]);
            acc[i][1] = fmaf(gv.x, d0.y, acc[i][1]);
            acc[i][2] = fmaf(gv.x, d0.z, acc[i][2]);
            acc[i][3] = fmaf(gv.x, d0.w, acc[i][3]);
            acc[i][0] = fmaf(gv.y, d1.x, acc[i][0]);
            acc[i][1] = fmaf(gv.y, d1.y, acc[i][1]);
            acc[i][2] = fmaf(gv.y, d1.z, acc[i][2]);
            acc[i][3] = fmaf(gv.y, d1.w, acc[i][3]);
            acc[i][0] = fmaf(gv.z, d2.x, acc[i][0]);
            acc[i][1] = fmaf(gv.z, d2.y, acc[i][1]);
            acc[i][2] = fmaf(gv.z, d2.z, acc[i][2]);
            acc[i][3] = fmaf(gv.z, d2.w, acc[i][3]);
            acc[i][0] = fmaf(gv.w, d3.x, acc[i][0]);
            acc[i][1] = fmaf(gv.w, d3.y, acc[i][1]);
            acc[i][2] = fmaf(gv.w, d3.z, acc[i][2]);
            acc[i][3] = fmaf(gv.w, d3.w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < G_TBV; ++i) {
          const int b = bg + i * BGn;
          if (b < B) {
            float4* dst = reinterpret_cast<float4*>(dvs + b * Kp + 4 * kg);  // this (b, kg) belongs to one thread
            float4 cur = *dst;
            cur.x += acc[i][0];
            cur.y += acc[i][1];
            cur.z += acc[i][2];
            cur.w += acc[i][3];
            *dst = cur;
          }
        }
      }
    }

    if (a.want_dD) {
      // thread tile: 4 pixels x 4 atoms, atoms fastest across lanes (coalesced D/m/s rows)
      constexpr int PG = TP / 4;
      for (int t = tid; t < PG * KG; t += kThreads) {
        const int pg = t / KG, kg = t - pg * KG;
        if (p0 + 4 * pg >= P) continue;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[j][c] = 0.0f;
        const float* gp = gs + 4 * pg;
        const float* vp = vs + 4 * kg;
#pragma unroll 4
        for (int b = 0; b < B; ++b) {
          const float4 gv = *reinterpret_cast<const float4*>(gp + b * TPS);
          const float4 vv = *reinterpret_cast<const float4*>(vp + b * Kp);
          acc[0][0] = fmaf(gv.x, vv.x, acc[0][0]);
          acc[0][1] = fmaf(gv.x, vv.y, acc[0][1]);
          acc[0][2] = fmaf(gv.x, vv.z, acc[0][2]);
          acc[0][3] = fmaf(gv.x, vv.w, acc[0][3]);
          acc[1][0] = fmaf(gv.y, vv.x, acc[1][0]);
          acc[1][1] = fmaf(gv.y, vv.y, acc[1][1]);
          acc[1][2] = fmaf(gv.y, vv.z, acc[1][2]);
          acc[1][3] = fmaf(gv.y, vv.w, acc[1][3]);
          acc[2][0] = fmaf(gv.z, vv.x, acc[2][0]);
          acc[2][1] = fmaf(gv.z, vv.y, acc[2][1]);
          acc[2][2] = fmaf(gv.z, vv.z, acc[2][2]);
          acc[2][3] = fmaf(gv.z, vv.w, acc[2][3]);
          acc[3][0] = fmaf(gv.w, vv.x, acc[3][0]);
          acc[3][1] = fmaf(gv.w, vv.y, acc[3][1]);
          acc[3][2] = fmaf(gv.w, vv.z, acc[3][2]);
          acc[3][3] = fmaf(gv.w, vv.w, acc[3][3]);
        }
        const float* Dsrc = a.want_dv ? Ds : nullptr;
        if ((K & 3) == 0) dict_epilogue<4>(a, acc, p0, pg, kg, Dsrc, Kp);
        else if ((K & 1) == 0) dict_epilogue<2>(a, acc, p0, pg, kg, Dsrc, Kp);
        else dict_epilogue<1>(a, acc, p0, pg, kg, Dsrc, Kp);
      }
    }
  }

  if (a.want_dv) {
    __syncthreads();
    float* dst = a.partial + (size_t)blockIdx.x * B * K;
    for (int e = tid; e < B * K; e += kThreads) {
      const int b = e / K, k = e - b * K;
      dst[e] = dvs[b * Kp + k];
    }
  }
}

// Second stage of the deterministic dv reduction: dvb[e] = sum over the per-CTA slabs, in a fixed order.  A block owns
// 32 consecutive outputs; warp w of 32 adds slabs w, w+32, ... (at most 19 independent coalesced 128-byte loads, all
// in flight together), then the 32 partial sums are combined in warp order -- the summation tree depends only on
// nslabs, so results are bit-reproducible.
__global__ void __launch_bounds__(1024) reduce_partials_kernel(float* __restrict__ dvb, const float* __restrict__ partial,
                                                               int n, int nslabs, int K, int ld_out, int slab_step) {
  __shared__ float part[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  // (slab_step > 1: the slabs of slab_step column windows are interleaved -- CTA c of the contraction kernel wrote window
  // c % slab_step; blockIdx.y is the window: its slabs start at slab blockIdx.y, its outputs at column blockIdx.y * K)
  partial += (size_t)blockIdx.y * n;
  dvb += (size_t)blockIdx.y * K;
  // launched with programmatic stream serialisation: the grid may be resident before the contraction kernel has
  // finished; this waits for its completion and for the visibility of its slabs
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float acc = 0.0f;
  if (e < n) {
#pragma unroll 4
    for (int c = warp; c < nslabs; c += 32) acc += partial[(size_t)c * slab_step * n + e];
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && e < n) {
    float t = part[0][lane];
#pragma unroll
    for (int w = 1; w < 32; ++w) t += part[w][lane];
    dvb[ld_out == K ? (size_t)e : (size_t)(e / K) * ld_out + (e % K)] = t;  // (rows of a wider [B, ld_out] array)
  }
}

size_t grad_smem_bytes(int TP, int B, int Kp, bool want_dD, bool want_dv) {
  size_t f = (size_t)B * (TP + 4);
  if (want_dv) f += (size_t)TP * Kp + (size_t)B * Kp;
  if (want_dD) f += (size_t)B * Kp;
  return f * sizeof(float);
}

template <int TP>
int launch_grad_tp(const GradArgs& a, size_t smem, int grid, cudaStream_t st) {
  int rc = check_cuda(cudaFuncSetAttribute(grad_fma_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(grad_fma)");
  if (rc) return rc;
  grad_fma_kernel<TP><<<grid, kThreads, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "grad_fma_kernel launch");
}

}  // namespace

// (slab_step: the slabs to add are slab_step slabs apart -- slab_step column windows contracted by one launch, all of
// them reduced by this one launch: nslabs slabs each)
int launch_reduce_partials(float* dvb, const float* partial, int n, int nslabs, int K, int ld_out, cudaStream_t st,
                           int slab_step = 1) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)((n + 31) / 32), (unsigned)slab_step);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, reduce_partials_kernel, dvb, partial, n, nslabs, K, ld_out, slab_step),
                    "reduce_partials_kernel launch");
}

namespace {
// codes_out[b, :] = v[v_index ? v_index[b] : b, :]  (the contiguous block adil_synth leaves for the backward kernel)
__global__ void gather_codes_kernel(float* codes_out, const float* v, const int64_t* vidx, int K) {
  const int b = blockIdx.x;
  const long long row = vidx ? (long long)vidx[b] : (long long)b;
  for (int k = threadIdx.x; k < K; k += blockDim.x) codes_out[(size_t)b * K + k] = v[row * K + k];
}
}  // namespace

int launch_synth_fma(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2,
                     const float* v, const int64_t* v_index, float* codes_out, int B, int P, int K,
                     const ChannelConsts& cc, float eps, int flags, cudaStream_t st) {
  if (codes_out != nullptr) {
    gather_codes_kernel<<<B, 64, 0, st>>>(codes_out, v, v_index, K);
    int rcg = check_cuda(cudaGetLastError(), "gather_codes_kernel launch");
    if (rcg) return rcg;
  }
  SynthArgs a;
  a.out = out; a.delta = delta_out; a.x = x; a.xidx = x_index; a.D2 = D2; a.v = v; a.vidx = v_index;
  a.B = B; a.P = P; a.K = K; a.Kp = round_up(K, 4);
  a.eps = eps; a.flags = flags; a.cc = cc;
  a.cc.use = (flags & ADIL_SYNTH_NORMALIZE) ? 1 : 0;
  const size_t dt_bytes = (size_t)a.Kp * S_TPS * sizeof(float);
  const size_t budget = 218 * 1024;
  int bch = (int)((budget - dt_bytes) / (a.Kp * sizeof(float)));
  bch = bch / S_TB * S_TB;
  if (bch > 128) bch = 128;
  if (bch > round_up(B, S_TB)) bch = round_up(B, S_TB);
  if (bch < S_TB) return set_error(-3, "adil_synth: K=%d too large for the FMA path", K);
  a.bch = bch;
  const size_t smem = dt_bytes + (size_t)bch * a.Kp * sizeof(float) + (size_t)bch * sizeof(long long);
  int rc = check_cuda(cudaFuncSetAttribute(synth_fma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "cudaFuncSetAttribute(synth_fma)");
  if (rc) return rc;
  const int ntiles = (P + S_TP - 1) / S_TP;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  synth_fma_kernel<<<grid, kThreads, smem, st>>>(a);
  return check_cuda(cudaGetLastError(), "synth_fma_kernel launch");
}

int launch_grad_fma(float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g, const float* D2,
                    const float* v, const int64_t* v_index, int B, int P, int K, const ChannelConsts& cc,
                    const AdamwDev* hp, int atoms_mode, float* scratch, size_t scratch_bytes, const GradOpts& opt,
                    cudaStream_t st) {
  GradArgs a;
  a.dD2 = dD2; a.D2w = D2_rw; a.m = m; a.s = s; a.partial = scratch;
  a.g = g; a.D2 = D2; a.v = v; a.vidx = v_index;
  a.B = B; a.P = P; a.K = K; a.Kp = round_up(K, 4);
  a.want_dD = (dD2 != nullptr || D2_rw != nullptr) ? 1 : 0;
  a.want_dv = (dvb != nullptr || opt.keep_partials) ? 1 : 0;
  a.accumulate = (opt.accumulate && D2_rw == nullptr) ? 1 : 0;
  a.delta = opt.delta;
  a.l2_coef = opt.l2_coef;
  a.atoms_mode = atoms_mode;
  a.cc = cc;
  if (hp) a.hp = *hp;
  if (!a.want_dD && !a.want_dv) return 0;
  // tile size: largest that still allows 2 CTAs per SM, else largest that fits at all
  const int tps[4] = {128, 64, 32, 16};
  int TP = 0;
  size_t smem = 0;
  for (int i = 0; i < 4 && !TP; ++i) {
    size_t sm = grad_smem_bytes(tps[i], B, a.Kp, a.want_dD, a.want_dv);
    if (sm <= 110 * 1024) { TP = tps[i]; smem = sm; }
  }
  for (int i = 0; i < 4 && !TP; ++i) {
    size_t sm = grad_smem_bytes(tps[i], B, a.Kp, a.want_dD, a.want_dv);
    if (sm <= 227 * 1024) { TP = tps[i]; smem = sm; }
  }
  if (!TP) return set_error(-3, "adil_grad: B=%d x K=%d does not fit the FMA path's shared memory; split the batch", B, K);
  const int ntiles = (P + TP - 1) / TP;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;
  int grid = sm_count() * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid > kMaxGradCtas) grid = kMaxGradCtas;
  if (a.want_dv) {
    const size_t need = (size_t)grid * B * K * sizeof(float);
    if (scratch == nullptr || scratch_bytes < need)
      return set_error(-2, "adil_grad: scratch too small (%zu < %zu bytes)", scratch_bytes, need);
  }
  int rc;
  switch (TP) {
    case 128: rc = launch_grad_tp<128>(a, smem, grid, st); break;
    case 64: rc = launch_grad_tp<64>(a, smem, grid, st); break;
    case 32: rc = launch_grad_tp<32>(a, smem, grid, st); break;
    default: rc = launch_grad_tp<16>(a, smem, grid, st); break;
  }
  if (rc) return rc;
  if (a.want_dv) {
    if (opt.keep_partials) {
      if (opt.nslabs_out) *opt.nslabs_out = grid;
      return 0;
    }
    return launch_reduce_partials(dvb, scratch, B * K, grid, K, K, st);
  }
  return 0;
}

// Largest batch one call of the FMA backward kernel can take at this atom count (smallest tile, 227 KB of shared memory).
int grad_fma_max_batch(int K, bool want_dD, bool want_dv) {
  const int Kp = round_up(K, 4);
  int lo = 0, hi = 1 << 16;
  while (lo < hi) {  // grad_smem_bytes is increasing in B
    const int mid = (lo + hi + 1) / 2;
    if (grad_smem_bytes(16, mid, Kp, want_dD, want_dv) <= (size_t)227 * 1024) lo = mid; else hi = mid - 1;
  }
  return lo;
}

}  // namespace adil
