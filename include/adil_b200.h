/*
 * adil_b200.h -- C ABI of libadil_b200.so: the B200 (sm_100a) kernels of ADiL's attack-learning hot path.
 *
 * The reference (flavie-yuan-liu/DL_attack_on_ImageNet) has no FFI: its boundary is the Python class
 * `ADIL` (attacks/attacks_classes/adil.py:38).  Each entry point below replaces the PyTorch op sequence of
 * one piece of that class; the reference lines replaced are cited per function.  The Python mirror
 * (dl_attack_on_imagenet_b200/adil.py) binds these through ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous fp32 unless stated ("host"); int64 index arrays are
 *    device pointers too.  The library never allocates persistent device memory: the caller (PyTorch) owns
 *    every buffer, including the scratch areas whose size is returned by *_scratch_bytes().
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *  - return value: 0 = ok; >0 = cudaError_t; <0 = argument / capability error.  adil_last_error() returns
 *    a thread-local description of the last non-zero return.
 *  - the int64 index arrays of adil_synth / adil_grad / adil_grad_dict_step (x_index, v_index) may ALSO be HOST
 *    pointers (pageable or pinned; the reference's DataLoader hands out CPU index tensors, adil.py:168): the
 *    indices are then read at call time and travel as kernel parameters, which removes the dependent cold
 *    miss on the index array at the top of the kernel (~1.4 us).  Needs the tcgen05 path (B <= 128 per pass,
 *    K <= 128, adil_tc_supported); otherwise -4 is returned and the caller passes device arrays.
 *  - P = C*hw pixels per image (hw = H*W), must be a multiple of 4.  K = atoms (1..256).  D2 is the
 *    dictionary viewed as [P, K] row-major (atoms innermost, adil.py:148 creates [C,H,W,K]).
 */
#ifndef ADIL_B200_H_
#define ADIL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADIL_VERSION 211

#define ADIL_MAX_CHANNELS 8
#define ADIL_MAX_ATOMS 256

/* adil_synth flags */
#define ADIL_SYNTH_NORMALIZE   1 /* out = (out - mean[c]) / std[c]      demo_dL_attack.py:22-25 */
#define ADIL_SYNTH_CLAMP_DELTA 2 /* delta = clamp(delta, -eps, eps)     adil.py:482             */
#define ADIL_SYNTH_CLAMP01     4 /* out = clamp(x + delta, 0, 1)        adil.py:484,567,623     */

/* row (coding-vector) projection modes */
#define ADIL_ROWS_NONE       0
#define ADIL_ROWS_L1BALL     1 /* utils.py:21-41  project_onto_l1_ball (adil.py:29-31,631-633) */
#define ADIL_ROWS_L2BALL     2 /* adil.py:626-629 radius * v / max(||v||_2, radius)            */
#define ADIL_ROWS_SOFTSHRINK 3 /* utils.py:159-161 get_prox_l1 (radius = lambda)               */

/* atom (dictionary) projection modes */
#define ADIL_ATOMS_NONE     0
#define ADIL_ATOMS_CLAMP1   1 /* adil.py:33-35,642  clamp(D, -1, 1)                    */
#define ADIL_ATOMS_L2BALL   2 /* utils.py:52-54     d_k / max(||d_k||_2, 1)            */
#define ADIL_ATOMS_L2SPHERE 3 /* utils.py:49-51     d_k / ||d_k||_2                    */
#define ADIL_ATOMS_L1BALL   4 /* utils.py:55-56     project_onto_l1_ball(d[:,:,:,k], 1): every (channel, atom)
                               *                    column of H*W pixels onto the unit l1 ball (adil_project_atoms only) */

/* adil_grad / adil_grad_dict_step flags */
#define ADIL_GRAD_ACCUMULATE_DD 1 /* adil_grad: dD2 += g^T v instead of dD2 = (minibatches larger than one pass:
                                   * chunks of the batch accumulate in stream order) */
#define ADIL_GRAD_KEEP_PARTIALS 2 /* leave the per-CTA partial code gradients [nslabs][B][K] in `scratch` instead of
                                   * reducing them into dvb (dvb may be NULL): adil_code_step adds them up itself,
                                   * which saves the reduction launch between the two kernels */

/* kernel implementation selector for adil_synth / adil_grad* (adil_set_impl) */
#define ADIL_IMPL_AUTO 0 /* tcgen05 when the shape qualifies, else FMA */
#define ADIL_IMPL_FMA  1 /* CUDA-core FMA kernels                       */
#define ADIL_IMPL_TC   2 /* tcgen05 split-TF32 kernels (error if the shape does not qualify) */

/* AdamW hyper-parameters of ONE update (torch.optim.AdamW as used at adil.py:154,250-251,531,588).
 * `step` is the 1-based step count t of this update; the bias corrections 1-beta^t are evaluated in double
 * precision on the host exactly as torch/optim/adam.py does. */
typedef struct adil_adamw {
  double lr;
  double beta1;
  double beta2;
  double eps;
  double weight_decay;
  long long step;
} adil_adamw_t;

int adil_version(void);
const char* adil_last_error(void);

/* Number of SMs / compute capability of the current device (needs a GPU). */
int adil_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Opt-in: keep [base, base + bytes) -- the dictionary D2, which adil_synth and adil_grad_dict_step both read every step
 * with 28 ms of classifier traffic in between -- in the persisting set-aside of the L2 cache: sets the device's
 * persisting-L2 limit to min(bytes, device maximum) and an access-policy window (hit ratio = limit / window) on `stream`,
 * so that the accesses of every kernel launched in that stream afterwards mark the lines persisting.  base == NULL or
 * bytes == 0 removes the window and resets the persisting lines.  The reference has no counterpart (cuBLAS / foreach
 * kernels re-read D from HBM); off by default -- measured effect in DESIGN.md section 7.4. */
int adil_l2_persist(const void* base, size_t bytes, void* stream);

/* Select the kernel family used by adil_synth / adil_grad / adil_grad_dict_step (process-wide). */
int adil_set_impl(int impl);
int adil_get_impl(void);
/* Bit 0: adil_synth runs on the tcgen05 path for this shape; bit 1: adil_grad / adil_grad_dict_step do. */
int adil_tc_supported(int B, int P, int K);

/* Perturbation synthesis (replaces adil.py:25-26 tensordot + add, the Normalize module of
 * demo_dL_attack.py:16-25, and the clamps of adil.py:481-484,563-567,617-623):
 *     delta[b,:] = sum_k v[v_index ? v_index[b] : b, k] * D2[:, k]
 *     out[b,:]   = f( x[x_index ? x_index[b] : b, :] + delta[b,:] )   (x == NULL: out = f(delta))
 * out, delta_out: [B,P] (either may be NULL, not both).  mean/std: HOST arrays of C floats (may be NULL when
 * NORMALIZE is not set).
 * codes_out (may be NULL): [B,K], receives the gathered code rows v[v_index[b], :] of the batch.  Handing this block
 * to adil_grad / adil_grad_dict_step of the same step as `v` with v_index == NULL (the rows do not change in between:
 * autograd saves the same tensor, adil.py:25,185) spares the backward kernel a gather of its own: a contiguous,
 * 16-byte-aligned block arrives in its shared memory as one bulk copy. */
int adil_synth(float* out, float* delta_out, const float* x, const int64_t* x_index, const float* D2, const float* v,
               const int64_t* v_index, float* codes_out, int B, int P, int K, int C, int hw, const float* mean_host,
               const float* std_host, float eps, int flags, void* stream);

/* Scratch (bytes) the grad entry points need for the deterministic cross-CTA reduction of the code gradients. */
size_t adil_grad_scratch_bytes(int B, int K);

/* Backward contractions (replaces the autograd backward of adil.py:25-26 + Normalize: adil.py:185,281,308,606):
 *     gx = g / std[c]      (std_host == NULL: gx = g)
 *     dD2[p,k] = sum_b gx[b,p] * v[v_index[b],k]          (skipped when dD2 == NULL)
 *     dvb[b,k] = sum_p gx[b,p] * D2[p,k]                  (skipped when dvb == NULL and KEEP_PARTIALS is not set)
 * g: [B,P] gradient w.r.t. the classifier input.  dvb: [B,K] in batch order (the caller scatters by v_index).
 * flags: ADIL_GRAD_*.  nslabs_out (host int, may be NULL): number of partial slabs left in scratch (KEEP_PARTIALS).
 * B is limited per call (adil_grad_max_batch); larger minibatches are passed in chunks with ACCUMULATE_DD.
 * delta / l2_coef: the l2 penalty 0.5 * l2_coef * ||D v||^2 of the regularised variants (adil_regularized.py:112-114,
 * 273-274): with delta = the synthesised perturbation [B,P] (adil_synth's delta_out) the contractions run on
 * gx + l2_coef * delta.  delta == NULL or l2_coef == 0: no penalty.  (The penalised form runs on the CUDA-core kernels.) */
int adil_grad(float* dD2, float* dvb, const float* g, const float* D2, const float* v, const int64_t* v_index, int B,
              int P, int K, int C, int hw, const float* std_host, const float* delta, float l2_coef, int flags,
              int* nslabs_out, void* scratch, size_t scratch_bytes, void* stream);

/* Largest B one adil_grad / adil_grad_dict_step call accepts for this shape with the current kernel family
 * (tcgen05 path: 128 images per pass; FMA path: bounded by shared memory; fused < 0: the FMA path's limit). */
int adil_grad_max_batch(int P, int K, int hw, int fused);

/* Single-GPU fusion of adil_grad with the dictionary AdamW step and projection (adil.py:185-188 for D):
 * dD2 never touches HBM; D2, m, s are updated in place.  atoms_mode: ADIL_ATOMS_NONE or ADIL_ATOMS_CLAMP1.
 * dvb is computed against the PRE-update D2, like autograd does.  flags: ADIL_GRAD_KEEP_PARTIALS. */
int adil_grad_dict_step(float* D2, float* m, float* s, float* dvb, const float* g, const float* v,
                        const int64_t* v_index, int B, int P, int K, int C, int hw, const float* std_host,
                        const adil_adamw_t* hp, int atoms_mode, int flags, int* nslabs_out, void* scratch,
                        size_t scratch_bytes, void* stream);

/* Dictionary AdamW step + elementwise projection on n contiguous elements (a [P_begin,P_end) x K slice):
 * replaces optimise.step() on d + update_d (adil.py:186,188 ; 310-311).  Used after the dD all-reduce on
 * multi-GPU runs.  atoms_mode: ADIL_ATOMS_NONE or ADIL_ATOMS_CLAMP1. */
int adil_dict_step(float* D2, float* m, float* s, const float* dD2, long long n, const adil_adamw_t* hp,
                   int atoms_mode, void* stream);

/* Dictionary step with any per-atom projection (the regularised variants: adil_regularized.py:23-28,141-146,283-285,
 * 467-469 -- `d = d - step * grad_d ; d = constraint_dict(d)`):
 *     hp != NULL: AdamW update with these hyper-parameters;  hp == NULL: plain gradient step D2 -= step * dD2
 *     then atoms_mode: NONE, CLAMP1, L2BALL, L2SPHERE (column norms over all P rows: two passes, fixed summation order)
 * D2, dD2 (and m, s with AdamW): [P,K].  scratch: adil_project_atoms_scratch_bytes(K) bytes. */
int adil_dict_step_atoms(float* D2, float* m, float* s, const float* dD2, int P, int K, const adil_adamw_t* hp, float step,
                         int atoms_mode, void* scratch, void* stream);

/* Proximal gradient step on the code rows of ONE minibatch (adil_regularized.py:304, 414-416, 570-573):
 *     v[v_index[b], :] = prox(v[v_index[b], :] - step * dvb[b, :])
 * rows_mode / radius as adil_project_rows (SOFTSHRINK with radius = step * lambda is the l1 prox).  Rows outside the
 * batch are untouched; a row named twice takes the update of its last slot. */
int adil_code_prox_step(float* v, const float* dvb, const int64_t* v_index, int B, int N, int K, float step,
                        int rows_mode, float radius, void* stream);

/* The dictionary side of one multi-GPU minibatch step as ONE kernel over peer-mapped memory (NVLink 5 / NVSwitch P2P;
 * the intent of the reference's DDP variant, adil.py:379-383):
 *     slice gradient = sum over ranks q of dD_peers[q][slice]           peer loads, summed in rank order
 *     AdamW + projection on this rank's slice of the dictionary          moments m, s exist for the slice only
 *     the new slice is stored into D_peers[q][slice] for every rank q    peer stores
 * i.e. reduce-scatter + optimizer step + all-gather without a collective library call or a staging buffer.
 * D_peers / dD_peers: HOST arrays of `world` device pointers to every rank's [rows_total, K] dictionary / gradient
 * buffer, mapped into this process (e.g. torch symmetric memory); entry `rank` is the local buffer.  The slice is
 * elements [slice_begin, slice_begin + slice_elems) of those buffers (multiples of 4).  The caller places the launch
 * between two cross-rank barriers on `stream`: every rank's gradient complete before, every rank's stores landed
 * before the dictionary is read again.  atoms_mode: ADIL_ATOMS_NONE or ADIL_ATOMS_CLAMP1.  world <= ADIL_MAX_PEERS. */
#define ADIL_MAX_PEERS 16
/* D_mc / dD_mc (optional, both or neither): MULTICAST addresses of the same two buffers (NVLS: one address that names
 * the buffer of every rank; torch symmetric memory's multicast_ptr).  When given, the gradient slice is read with
 * multimem.ld_reduce -- the NVSwitch adds the ranks' values in flight, each rank receives 4PK/R bytes instead of
 * (R-1)/R * 4PK -- and the new dictionary slice is written with multimem.st, which the switch replicates to every rank. */
int adil_dict_step_peer(const void* const* D_peers, const void* const* dD_peers, float* m, float* s,
                        long long slice_begin, long long slice_elems, int rank, int world, const adil_adamw_t* hp,
                        int atoms_mode, const void* D_mc, const void* dD_mc, void* stream);

/* Code AdamW step over ALL N rows (dense gradient, zero outside the batch -- adil.py:154,186) fused with the
 * scatter of dvb by v_index (duplicates accumulate, like index_put_(accumulate=True)) and the row projection
 * (adil.py:187 update_v).  v, m, s: [N,K]; dvb: [B,K] (NULL: zero gradient).
 * partial / nslabs: instead of dvb, the per-CTA slabs [nslabs][B][K] a backward call left in its scratch
 * (ADIL_GRAD_KEEP_PARTIALS): the gradient of batch slot b is the sum over the slabs, added up in a fixed order. */
int adil_code_step(float* v, float* m, float* s, const float* dvb, const int64_t* v_index, int B, int N, int K,
                   const adil_adamw_t* hp, int rows_mode, float radius, const float* partial, int nslabs,
                   void* stream);

/* Row projection only (adil.py:625-633 projection_v ; utils.py:21-41 ; utils.py:159-161).  In place. */
int adil_project_rows(float* v, int N, int K, int rows_mode, float radius, void* stream);

/* Per-atom projection of D2 [P,K] (adil.py:635-642 projection_d ; utils.py:44-57 constraint_dict).  C = number of
 * channels (P = C*hw; only the l1ball mode uses it: utils.py:23 views an atom as [C, H*W] rows).
 * scratch: device buffer of adil_project_atoms_scratch_bytes(K) bytes (unused for CLAMP1). */
size_t adil_project_atoms_scratch_bytes(int K);
int adil_project_atoms(float* D2, int P, int K, int C, int atoms_mode, void* scratch, void* stream);

/* Elementwise AdamW + clamp(+-bound) on n elements: the z update of forward_supervised_DDrague
 * (adil.py:531,554-555).  bound <= 0: no clamp. */
int adil_adamw_clamp(float* p, float* m, float* s, const float* grad, long long n, const adil_adamw_t* hp,
                     float bound, void* stream);

/* Per-image error reductions of a batch of adversarial images against the clean ones: the sums behind
 * compute_rmse / compute_mse (performance.py:249-266) and the l_inf norm printed by forward_unsupervised
 * (adil.py:503-505), in ONE pass over both arrays:
 *   err2[i] = sum_p (adv[i,p] - clean[i,p])^2      ref2[i] = sum_p clean[i,p]^2      linf[i] = max_p |adv - clean|
 * adv, clean: [n, P] (P % 4 == 0, 16-byte aligned); err2 / ref2 / linf: [n] each, any of them may be NULL.
 * Fixed summation order (no atomics): bit-reproducible.  scratch: adil_image_errors_scratch_bytes(n) bytes. */
size_t adil_image_errors_scratch_bytes(int n);
int adil_image_errors(float* err2, float* ref2, float* linf, const float* adv, const float* clean, int n, int P,
                      void* scratch, size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADIL_B200_H_ */
