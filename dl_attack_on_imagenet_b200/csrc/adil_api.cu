// C ABI entry points + host-side plumbing of libadil_b200.so (see include/adil_b200.h).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "adil_common.cuh"

namespace adil {

namespace {
thread_local char g_err[512] = "";
int g_impl = ADIL_IMPL_AUTO;
}  // namespace

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool is_host_pointer(const void* p) {
  if (p == nullptr) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();  // (pre-11.0 behaviour for unregistered host memory)
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered || at.type == cudaMemoryTypeHost;
}

AdamwDev make_adamw(const adil_adamw_t* hp) {
  AdamwDev d;
  const double t = (double)hp->step;
  const double bc1 = 1.0 - std::pow(hp->beta1, t);
  const double bc2 = 1.0 - std::pow(hp->beta2, t);
  d.decay = (float)(1.0 - hp->lr * hp->weight_decay);
  d.w1 = (float)(1.0 - hp->beta1);
  d.beta2 = (float)hp->beta2;
  d.w2 = (float)(1.0 - hp->beta2);
  d.bc2_sqrt = (float)std::sqrt(bc2);
  d.eps = (float)hp->eps;
  d.neg_step = (float)(-(hp->lr / bc1));
  d.rbc2_sqrt = (float)(1.0 / (double)d.bc2_sqrt);
  d.lerp_hi = std::fabs(1.0 - hp->beta1) >= 0.5 ? 1 : 0;
  return d;
}

ChannelConsts make_consts(int C, int hw, const float* mean_host, const float* std_host, bool use) {
  ChannelConsts cc;
  memset(&cc, 0, sizeof(cc));
  cc.C = C;
  cc.hw = hw > 0 ? hw : 1;
  cc.use = use ? 1 : 0;
  for (int c = 0; c < kMaxC; ++c) {
    cc.mean[c] = (use && mean_host && c < C) ? mean_host[c] : 0.0f;
    cc.stdv[c] = (use && std_host && c < C) ? std_host[c] : 1.0f;
    cc.rstd[c] = (float)(1.0 / (double)cc.stdv[c]);
  }
  return cc;
}

namespace {

int check_shape(const char* fn, int B, int P, int K, int C, int hw, bool need_channels) {
  if (B < 0 || P <= 0 || K < 1) return set_error(-1, "%s: bad shape B=%d P=%d K=%d", fn, B, P, K);
  if (K > ADIL_MAX_ATOMS) return set_error(-1, "%s: K=%d exceeds ADIL_MAX_ATOMS=%d", fn, K, ADIL_MAX_ATOMS);
  if (P % 4 != 0) return set_error(-1, "%s: P=%d must be a multiple of 4", fn, P);
  if (need_channels) {
    if (C < 1 || C > kMaxC) return set_error(-1, "%s: C=%d out of range [1,%d]", fn, C, kMaxC);
    if (hw <= 0 || (long long)C * hw != P) return set_error(-1, "%s: C*hw = %d*%d != P = %d", fn, C, hw, P);
  }
  return 0;
}

bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

bool use_tc(bool ok, int B, int P, int K, int* rc, const char* fn) {
  *rc = 0;
  if (g_impl == ADIL_IMPL_FMA) return false;
  if (g_impl == ADIL_IMPL_TC && !ok) {
    *rc = set_error(-4, "%s: ADIL_IMPL_TC requested but shape B=%d P=%d K=%d does not qualify", fn, B, P, K);
    return false;
  }
  return ok;
}

}  // namespace
}  // namespace adil

using namespace adil;

extern "C" int adil_version(void) { return ADIL_VERSION; }

extern "C" const char* adil_last_error(void) { return g_err; }

extern "C" int adil_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  cudaDeviceProp prop;
  rc = check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
  if (rc) return rc;
  if (sm) *sm = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}

extern "C" int adil_l2_persist(const void* base, size_t bytes, void* stream) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  cudaStreamAttrValue val;
  memset(&val, 0, sizeof(val));
  if (base == nullptr || bytes == 0) {
    val.accessPolicyWindow.num_bytes = 0;
    rc = check_cuda(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &val),
                    "cudaStreamSetAttribute(access policy window, off)");
    if (rc) return rc;
    return check_cuda(cudaCtxResetPersistingL2Cache(), "cudaCtxResetPersistingL2Cache");
  }
  int max_persist = 0, max_window = 0;
  rc = check_cuda(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev), "cudaDeviceGetAttribute");
  if (rc) return rc;
  rc = check_cuda(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev), "cudaDeviceGetAttribute");
  if (rc) return rc;
  if (max_persist <= 0 || max_window <= 0) return set_error(-3, "adil_l2_persist: the device has no persisting L2 set-aside");
  const size_t limit = bytes < (size_t)max_persist ? bytes : (size_t)max_persist;
  const size_t window = bytes < (size_t)max_window ? bytes : (size_t)max_window;
  rc = check_cuda(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, limit), "cudaDeviceSetLimit(persisting L2)");
  if (rc) return rc;
  val.accessPolicyWindow.base_ptr = const_cast<void*>(base);
  val.accessPolicyWindow.num_bytes = window;
  val.accessPolicyWindow.hitRatio = limit >= window ? 1.0f : (float)((double)limit / (double)window);
  val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  return check_cuda(cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &val),
                    "cudaStreamSetAttribute(access policy window)");
}

extern "C" int adil_set_impl(int impl) {
  if (impl < ADIL_IMPL_AUTO || impl > ADIL_IMPL_TC) return set_error(-1, "adil_set_impl: bad impl %d", impl);
  g_impl = impl;
  return 0;
}

extern "C" int adil_get_impl(void) { return g_impl; }

extern "C" int adil_tc_supported(int B, int P, int K) {
  const int hw = (P % 3 == 0) ? P / 3 : P;
  return (tc_synth_ok(B, P, K, hw) ? 1 : 0) | (tc_grad_ok(B, P, K, hw, true, true, true) ? 2 : 0);
}

extern "C" int adil_synth(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2,
                          const float* v, const int64_t* v_index, float* codes_out, int B, int P, int K, int C, int hw,
                          const float* mean_host, const float* std_host, float eps, int flags, void* stream) {
  const bool norm = (flags & ADIL_SYNTH_NORMALIZE) != 0;
  int rc = check_shape("adil_synth", B, P, K, C, hw, norm);
  if (rc) return rc;
  if (!D2 || !v || (!out && !delta_out)) return set_error(-1, "adil_synth: null pointer");
  if (norm && (!mean_host || !std_host)) return set_error(-1, "adil_synth: NORMALIZE needs mean/std");
  if (!aligned16(out) || !aligned16(delta_out) || !aligned16(x) || !aligned16(D2) || !aligned16(codes_out))
    return set_error(-1, "adil_synth: out/delta/x/D2/codes_out must be 16-byte aligned");
  if (B == 0) return 0;
  ChannelConsts cc = make_consts(C, hw, mean_host, std_host, norm);
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(tc_synth_ok(B, P, K, norm ? hw : P), B, P, K, &rc, "adil_synth"))
    return launch_synth_tc(out, delta_out, x, x_index, D2, v, v_index, codes_out, B, P, K, cc, eps, flags, st);
  if (rc) return rc;
  if (is_host_pointer(x_index) || is_host_pointer(v_index))
    return set_error(-4, "adil_synth: host index arrays travel as kernel parameters of the tcgen05 path; shape B=%d P=%d K=%d "
                     "(or ADIL_IMPL_FMA) needs device index arrays", B, P, K);
  return launch_synth_fma(out, delta_out, x, x_index, D2, v, v_index, codes_out, B, P, K, cc, eps, flags, st);
}

extern "C" size_t adil_grad_scratch_bytes(int B, int K) {
  if (B <= 0 || K <= 0) return 0;
  return (size_t)kMaxGradCtas * (size_t)B * (size_t)K * sizeof(float);
}

namespace {
int grad_common(const char* fn, float* dD2, float* D2_rw, float* m, float* s, float* dvb, const float* g,
                const float* D2, const float* v, const int64_t* v_index, int B, int P, int K, int C, int hw,
                const float* std_host, const adil_adamw_t* hp, int atoms_mode, int flags, int* nslabs_out,
                void* scratch, size_t scratch_bytes, void* stream, const float* delta = nullptr, float l2_coef = 0.0f) {
  const bool scale = std_host != nullptr;
  int rc = check_shape(fn, B, P, K, C, hw, scale);
  if (rc) return rc;
  if (!g || !D2 || !v) return set_error(-1, "%s: null pointer", fn);
  if (!aligned16(g) || !aligned16(D2) || !aligned16(dD2) || !aligned16(m) || !aligned16(s))
    return set_error(-1, "%s: g/D2/dD2/m/s must be 16-byte aligned", fn);
  if (B == 0) return set_error(-1, "%s: empty batch", fn);
  if (flags & ~(ADIL_GRAD_ACCUMULATE_DD | ADIL_GRAD_KEEP_PARTIALS)) return set_error(-1, "%s: unknown flags 0x%x", fn, flags);
  GradOpts opt;
  opt.accumulate = (flags & ADIL_GRAD_ACCUMULATE_DD) ? 1 : 0;
  opt.keep_partials = (flags & ADIL_GRAD_KEEP_PARTIALS) ? 1 : 0;
  opt.nslabs_out = nslabs_out;
  const bool penalised = delta != nullptr && l2_coef != 0.0f;
  opt.delta = penalised ? delta : nullptr;
  opt.l2_coef = penalised ? l2_coef : 0.0f;
  if (penalised && !aligned16(delta)) return set_error(-1, "%s: delta must be 16-byte aligned", fn);
  if (opt.keep_partials && !nslabs_out) return set_error(-1, "%s: ADIL_GRAD_KEEP_PARTIALS needs nslabs_out", fn);
  if (opt.accumulate && (D2_rw != nullptr || dD2 == nullptr))
    return set_error(-1, "%s: ADIL_GRAD_ACCUMULATE_DD applies to the plain dD2 output only", fn);
  const bool want_dv = dvb != nullptr || opt.keep_partials;
  ChannelConsts cc = make_consts(C, hw, nullptr, std_host, scale);
  AdamwDev dev;
  memset(&dev, 0, sizeof(dev));
  if (hp) dev = make_adamw(hp);
  cudaStream_t st = (cudaStream_t)stream;
  if (!penalised &&  // (the l2-penalised contractions run on the CUDA-core kernels)
      use_tc(tc_grad_ok(B, P, K, scale ? hw : P, dD2 != nullptr || D2_rw != nullptr, want_dv, D2_rw != nullptr), B, P,
             K, &rc, fn))
    return launch_grad_tc(dD2, D2_rw, m, s, dvb, g, D2, v, v_index, B, P, K, cc, &dev, atoms_mode, (float*)scratch,
                          scratch_bytes, opt, st);
  if (rc) return rc;
  if (is_host_pointer(v_index))
    return set_error(-4, "%s: a host index array travels as kernel parameters of the tcgen05 path; shape B=%d P=%d K=%d (or "
                     "ADIL_IMPL_FMA) needs a device index array", fn, B, P, K);
  return launch_grad_fma(dD2, D2_rw, m, s, dvb, g, D2, v, v_index, B, P, K, cc, &dev, atoms_mode, (float*)scratch,
                         scratch_bytes, opt, st);
}
}  // namespace

extern "C" int adil_grad(float* dD2, float* dvb, const float* g, const float* D2, const float* v,
                         const int64_t* v_index, int B, int P, int K, int C, int hw, const float* std_host,
                         const float* delta, float l2_coef, int flags, int* nslabs_out, void* scratch,
                         size_t scratch_bytes, void* stream) {
  if (!dD2 && !dvb && !(flags & ADIL_GRAD_KEEP_PARTIALS))
    return set_error(-1, "adil_grad: nothing to compute (dD2 and dvb both NULL)");
  return grad_common("adil_grad", dD2, nullptr, nullptr, nullptr, dvb, g, D2, v, v_index, B, P, K, C, hw, std_host,
                     nullptr, ADIL_ATOMS_NONE, flags, nslabs_out, scratch, scratch_bytes, stream, delta, l2_coef);
}

extern "C" int adil_grad_max_batch(int P, int K, int hw, int fused) {
  if (P <= 0 || K < 1 || K > ADIL_MAX_ATOMS) return 0;
  if (hw <= 0) hw = P;
  if (fused < 0) return grad_fma_max_batch(K, true, true);  // the CUDA-core kernels' limit (l2-penalised contractions)
  if (g_impl != ADIL_IMPL_FMA && tc_grad_ok(128, P, K, hw, true, true, fused != 0)) return 128;
  if (g_impl == ADIL_IMPL_TC) return 0;
  return grad_fma_max_batch(K, true, true);
}

extern "C" int adil_grad_dict_step(float* D2, float* m, float* s, float* dvb, const float* g, const float* v,
                                   const int64_t* v_index, int B, int P, int K, int C, int hw, const float* std_host,
                                   const adil_adamw_t* hp, int atoms_mode, int flags, int* nslabs_out, void* scratch,
                                   size_t scratch_bytes, void* stream) {
  if (!D2 || !m || !s || !hp) return set_error(-1, "adil_grad_dict_step: null pointer");
  if (atoms_mode != ADIL_ATOMS_NONE && atoms_mode != ADIL_ATOMS_CLAMP1)
    return set_error(-1, "adil_grad_dict_step: atoms_mode %d cannot be fused (use adil_project_atoms)", atoms_mode);
  return grad_common("adil_grad_dict_step", nullptr, D2, m, s, dvb, g, D2, v, v_index, B, P, K, C, hw, std_host, hp,
                     atoms_mode, flags, nslabs_out, scratch, scratch_bytes, stream);
}
