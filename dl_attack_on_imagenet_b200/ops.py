"""Tensor-level wrappers over the C ABI (include/adil_b200.h).

PyTorch owns every buffer; the wrappers hand raw device pointers and the current CUDA stream to the kernels.
All tensors must be contiguous fp32 (indices: int64) on the same CUDA device -- anything else raises; there
is no CPU path.
"""
import ctypes

import torch

from . import _lib
from ._lib import AdamwParams

# flags / modes (include/adil_b200.h)
SYNTH_NORMALIZE, SYNTH_CLAMP_DELTA, SYNTH_CLAMP01 = 1, 2, 4
ROWS_NONE, ROWS_L1BALL, ROWS_L2BALL, ROWS_SOFTSHRINK = 0, 1, 2, 3
ATOMS_NONE, ATOMS_CLAMP1, ATOMS_L2BALL, ATOMS_L2SPHERE, ATOMS_L1BALL = 0, 1, 2, 3, 4
GRAD_ACCUMULATE_DD, GRAD_KEEP_PARTIALS = 1, 2
IMPL_AUTO, IMPL_FMA, IMPL_TC = 0, 1, 2

_scratch = {}


class KernelTimer(object):
    """Opt-in device timing of the library's launches (the reference has only a wall-clock bracket around whole
    attacks, performance.py:136-144).  Inside `with ops.kernel_timer() as kt:` every wrapper below brackets its C-ABI
    call with CUDA events on the launching stream; `kt.summary()` (after a synchronize) gives calls and mean / total
    milliseconds per entry point.  Outside the context the wrappers record nothing."""

    def __init__(self):
        self.events = {}
        self.launches = 0

    def __enter__(self):
        global _timer
        self._prev, _timer = _timer, self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = self._prev
        return False

    def summary(self):
        out = {}
        for name, pairs in self.events.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            out[name] = {"calls": len(ms), "ms_mean": sum(ms) / len(ms), "ms_total": sum(ms),
                         "ms_min": min(ms), "ms_max": max(ms)}
        return out


_timer = None


def kernel_timer():
    return KernelTimer()


class _Timed(object):
    __slots__ = ("name", "device", "start", "n")

    def __init__(self, name, device, launches=1):
        self.name, self.device, self.n = name, device, launches

    def __enter__(self):
        if _timer is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if _timer is not None and exc[0] is None:
            end = torch.cuda.Event(enable_timing=True)
            end.record(torch.cuda.current_stream(self.device))
            _timer.events.setdefault(self.name, []).append((self.start, end))
            _timer.launches += self.n
        return False


def _f32(t, name, allow_none=False):
    if t is None:
        if allow_none:
            return None
        raise ValueError("%s is required" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: the ADiL kernels have no CPU fallback" % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("%s must be contiguous float32 (got %s, contiguous=%s)" % (name, t.dtype, t.is_contiguous()))
    return t


def _idx(t, name, device):
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(list(t) if not hasattr(t, '__array__') else t, dtype=torch.long)
    if t.dtype != torch.long:
        t = t.long()
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _idx_any(t, name, device, P, K, bit):
    """Index array for the kernels that take host OR device indices (synth: bit 1, grad*: bit 2).  A CPU index tensor
    of at most 128 entries stays on the host when the tcgen05 path applies -- the indices then travel as kernel
    parameters (no H2D copy, no cold miss in the kernel) -- anything else goes to the device."""
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(list(t) if not hasattr(t, '__array__') else t, dtype=torch.long)
    if t.dtype != torch.long:
        t = t.long()
    if (not t.is_cuda) and 0 < t.numel() <= 128 and get_impl() != IMPL_FMA and (
            _lib.lib().adil_tc_supported(int(t.numel()), int(P), int(K)) & bit):
        return t.contiguous()
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _host3(vals, C):
    if vals is None:
        return None
    arr = (ctypes.c_float * C)(*[float(a) for a in vals])
    return arr


def _get_scratch(device, nbytes, tag):
    """Scratch buffer per (device, stream, purpose): two streams calling the same kernel concurrently never share
    one (the partial code-gradient slabs of a backward call live here until the code step has consumed them)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


class CodePartials(object):
    """Per-CTA partial code gradients [nslabs, B, K] a backward call left in its scratch buffer
    (`keep_partials=True`); `code_step` adds them up itself.  Valid until the next backward call on the same
    stream."""

    def __init__(self, buf, nslabs, B, K):
        self.buf, self.nslabs, self.B, self.K = buf, int(nslabs), int(B), int(K)

    def reduce(self):
        """[B, K] code gradient (fixed summation order) -- for callers that want dvb after all."""
        n = self.nslabs * self.B * self.K
        return self.buf[:4 * n].view(torch.float32).view(self.nslabs, self.B, self.K).sum(dim=0)


def adamw_params(step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """torch.optim.AdamW defaults (adil.py:154); `step` is the 1-based count of this update."""
    return AdamwParams(float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step))


def set_impl(impl):
    _lib.check(_lib.lib().adil_set_impl(int(impl)), "adil_set_impl")


def get_impl():
    return _lib.lib().adil_get_impl()


def tc_supported(B, P, K):
    return bool(_lib.lib().adil_tc_supported(int(B), int(P), int(K)))


def l2_persist(t):
    """Opt-in (adil_l2_persist): keep tensor `t` (the dictionary) in the persisting set-aside of the L2 cache for the
    kernels of the current stream; `None` removes the window."""
    if t is None:
        dev = torch.device("cuda", torch.cuda.current_device())
        _lib.check(_lib.lib().adil_l2_persist(None, 0, _stream(dev)), "adil_l2_persist")
        return
    _f32(t, "t")
    _lib.check(_lib.lib().adil_l2_persist(_ptr(t), t.numel() * 4, _stream(t.device)), "adil_l2_persist")


def device_info():
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().adil_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)), "adil_device_info")
    return sm.value, ma.value, mi.value


def synth(D2, v, v_index=None, x=None, x_index=None, mean=None, std=None, eps=0.0, flags=0, out=None,
          delta_out=None, want_out=True, n_channels=None, codes_out=None):
    """out[b] = f(x[x_index[b]] + D2 @ v[v_index[b]])  -- adil.py:25-26 fused with Normalize / clamps.

    D2: [P,K]; v: [N,K]; x: [B,P] (or [Nx,P] with x_index).  Returns (out, delta_out) ([B,P] or None each).
    codes_out: optional [B,K] buffer that receives v[v_index] -- pass it as `v` (with v_index=None) to the backward
    call of the same step, which then fetches the rows with one bulk copy instead of gathering them again."""
    D2 = _f32(D2, "D2")
    v = _f32(v, "v")
    dev = D2.device
    P, K = D2.shape
    v_index = _idx_any(v_index, "v_index", dev, P, K, 1)
    x_index = _idx_any(x_index, "x_index", dev, P, K, 1)
    B = v_index.numel() if v_index is not None else v.shape[0]
    if x is not None:
        x = _f32(x, "x")
        xb = x_index.numel() if x_index is not None else x.shape[0]
        if xb != B or x[0].numel() != P:
            raise ValueError("x rows (%d x %d) do not match B=%d, P=%d" % (xb, x[0].numel(), B, P))
    if want_out and out is None:
        out = torch.empty((B, P), dtype=torch.float32, device=dev)
    if out is not None:
        _f32(out, "out")
    if delta_out is not None:
        _f32(delta_out, "delta_out")
    if codes_out is not None:
        _f32(codes_out, "codes_out")
        if tuple(codes_out.shape) != (B, K):
            raise ValueError("codes_out must be [B=%d, K=%d], got %s" % (B, K, tuple(codes_out.shape)))
    C = n_channels if n_channels is not None else (len(mean) if mean is not None else 1)
    hw = P // C
    mean_h, std_h = _host3(mean, C), _host3(std, C)
    with _Timed("adil_synth", dev, (B + 127) // 128):
        rc = _lib.lib().adil_synth(_ptr(out), _ptr(delta_out), _ptr(x), _ptr(x_index), _ptr(D2), _ptr(v),
                                   _ptr(v_index), _ptr(codes_out), B, P, K, C, hw, mean_h, std_h, float(eps),
                                   int(flags), _stream(dev))
    _lib.check(rc, "adil_synth")
    return out, delta_out


def _grad_scratch(dev, B, K):
    nbytes = _lib.lib().adil_grad_scratch_bytes(int(B), int(K))
    return _get_scratch(dev, nbytes, "grad"), nbytes


def grad_max_batch(P, K, hw=None, fused=False):
    """Largest minibatch one backward kernel call takes for this shape (larger ones are chunked by the wrappers)."""
    return int(_lib.lib().adil_grad_max_batch(int(P), int(K), int(hw if hw else P), 1 if fused else 0))


def _host_index_retry(call, v_index, dev):
    """Run `call(index)`; a host index array refused by the C side (-4: the shape does not run on the tcgen05 path
    after all) is moved to the device and the call repeated once."""
    rc = call(v_index)
    if rc == -4 and v_index is not None and not v_index.is_cuda:
        rc = call(v_index.to(dev))
    return rc


def _n_launches_with_reduction(B, P, K):
    """launches of a backward call that returns the reduced code gradient: the contraction kernel (beyond 128 atoms ONE
    launch covers both column windows) + the slab reduction (one launch, both windows)"""
    return 2


def _grad_call(dD2, dvb, g, D2, v, v_index, B, P, K, std, flags, dev, delta=None, l2_coef=0.0):
    """One adil_grad call (B within the per-call limit).  Returns nslabs (KEEP_PARTIALS) or 0."""
    C = len(std) if std is not None else 1
    scratch, nbytes = _grad_scratch(dev, B, K)
    nslabs = ctypes.c_int(0)
    v_index = _idx_any(v_index, "v_index", dev, P, K, 2)
    with _Timed("adil_grad", dev, 1 if (flags & GRAD_KEEP_PARTIALS or dvb is None) else _n_launches_with_reduction(B, P, K)):
        rc = _host_index_retry(
            lambda ix: _lib.lib().adil_grad(_ptr(dD2), _ptr(dvb), _ptr(g), _ptr(D2), _ptr(v), _ptr(ix), B, P, K, C,
                                            P // C, _host3(std, C), _ptr(delta), float(l2_coef), int(flags),
                                            ctypes.byref(nslabs), _ptr(scratch), nbytes, _stream(dev)), v_index, dev)
    _lib.check(rc, "adil_grad")
    return scratch, nslabs.value


def grad(g, D2, v, v_index=None, std=None, want_dD=True, want_dv=True, dD2=None, dvb=None, accumulate=False,
         keep_partials=False, delta=None, l2_coef=0.0):
    """Backward contractions (adil.py:185): returns (dD2 [P,K] or None, dvb [B,K] or None).

    delta, l2_coef: l2 penalty 0.5 * l2_coef * ||D v||^2 of the regularised variants (adil_regularized.py:112-114):
    `delta` [B,P] is the synthesised perturbation (synth's delta_out); the contractions then run on gx + l2_coef * delta.

    accumulate: dD2 += instead of dD2 = .  keep_partials: the second result is a `CodePartials` handle for
    `code_step` instead of the reduced dvb (one launch fewer).  Minibatches beyond the per-call limit of the kernels
    (128 images on the tcgen05 path) are processed in chunks that accumulate into dD2."""
    g = _f32(g, "g")
    D2 = _f32(D2, "D2")
    v = _f32(v, "v")
    dev = D2.device
    P, K = D2.shape
    if v_index is not None and not torch.is_tensor(v_index):
        v_index = torch.as_tensor(list(v_index) if not hasattr(v_index, '__array__') else v_index, dtype=torch.long)
    B = v_index.numel() if v_index is not None else v.shape[0]
    if g.numel() != B * P:
        raise ValueError("g has %d elements, expected B*P = %d" % (g.numel(), B * P))
    keep_partials = keep_partials and K <= 128       # (more than 128 atoms run as two column windows: reduced dvb only)
    if accumulate and (dD2 is None or not want_dD):
        raise ValueError("grad: accumulate needs an existing dD2")
    if want_dD and dD2 is None:
        dD2 = torch.empty((P, K), dtype=torch.float32, device=dev)
    if want_dD:
        _f32(dD2, "dD2")
    C = len(std) if std is not None else 1
    penalised = delta is not None and l2_coef != 0.0
    if penalised:
        delta = _f32(delta, "delta")
        if delta.numel() != B * P:
            raise ValueError("delta has %d elements, expected B*P = %d" % (delta.numel(), B * P))
        delta = delta.view(B, P)
        bmax = min(grad_max_batch(P, K, P // C, False), int(_lib.lib().adil_grad_max_batch(int(P), int(K), int(P // C), -1)))
    else:
        delta = None
        bmax = grad_max_batch(P, K, P // C, False)
    if bmax < 1:
        raise RuntimeError("grad: shape P=%d K=%d is not supported by the selected kernel family" % (P, K))
    g2 = g.view(B, P)
    if B <= bmax:
        if want_dv and not keep_partials and dvb is None:
            dvb = torch.empty((B, K), dtype=torch.float32, device=dev)
        flags = (GRAD_ACCUMULATE_DD if accumulate else 0) | (GRAD_KEEP_PARTIALS if (want_dv and keep_partials) else 0)
        scratch, nslabs = _grad_call(dD2 if want_dD else None, dvb if (want_dv and not keep_partials) else None, g2,
                                     D2, v, v_index, B, P, K, std, flags, dev, delta, l2_coef)
        second = None
        if want_dv:
            second = CodePartials(scratch, nslabs, B, K) if keep_partials else dvb
        return (dD2 if want_dD else None), second
    # ---- chunked: B beyond one pass --------------------------------------------------------------------------
    if want_dv and dvb is None:
        dvb = torch.empty((B, K), dtype=torch.float32, device=dev)
    for i, b0 in enumerate(range(0, B, bmax)):
        b1 = min(B, b0 + bmax)
        if v_index is not None:
            vi, ix = v, v_index[b0:b1]
        else:
            vi, ix = v[b0:b1], None
        flags = GRAD_ACCUMULATE_DD if (want_dD and (accumulate or i > 0)) else 0
        _grad_call(dD2 if want_dD else None, dvb[b0:b1] if want_dv else None, g2[b0:b1], D2, vi, ix, b1 - b0, P, K, std,
                   flags, dev, delta[b0:b1] if penalised else None, l2_coef)
    return (dD2 if want_dD else None), (dvb if want_dv else None)


def grad_dict_step(D2, m, s, g, v, v_index, hp, std=None, atoms_mode=ATOMS_CLAMP1, want_dv=True, dvb=None,
                   keep_partials=False):
    """Single-GPU fused backward + dictionary AdamW + clamp (adil.py:185-188 for D).  Returns dvb [B,K] (or a
    `CodePartials` handle with keep_partials) or None.  Minibatches beyond the per-call limit run as chunked plain
    contractions that accumulate dD, followed by the stand-alone dictionary step."""
    D2 = _f32(D2, "D2")
    m = _f32(m, "m")
    s = _f32(s, "s")
    g = _f32(g, "g")
    v = _f32(v, "v")
    dev = D2.device
    P, K = D2.shape
    if v_index is not None and not torch.is_tensor(v_index):
        v_index = torch.as_tensor(list(v_index) if not hasattr(v_index, '__array__') else v_index, dtype=torch.long)
    B = v_index.numel() if v_index is not None else v.shape[0]
    if g.numel() != B * P:
        raise ValueError("g has %d elements, expected B*P = %d" % (g.numel(), B * P))
    C = len(std) if std is not None else 1
    keep_partials = keep_partials and K <= 128       # (more than 128 atoms run as two column windows: reduced dvb only)
    bmax = grad_max_batch(P, K, P // C, True)
    if bmax < 1:
        raise RuntimeError("grad_dict_step: shape P=%d K=%d is not supported by the selected kernel family" % (P, K))
    if B > bmax:
        dD2 = _get_scratch(dev, 4 * P * K, "dD")[:4 * P * K].view(torch.float32).view(P, K)
        _, dvb = grad(g, D2, v, v_index, std, want_dD=True, want_dv=want_dv, dD2=dD2, dvb=dvb)
        dict_step(D2, m, s, dD2, hp, atoms_mode)
        return dvb if want_dv else None
    if want_dv and not keep_partials and dvb is None:
        dvb = torch.empty((B, K), dtype=torch.float32, device=dev)
    flags = GRAD_KEEP_PARTIALS if (want_dv and keep_partials) else 0
    scratch, nbytes = _grad_scratch(dev, B, K)
    nslabs = ctypes.c_int(0)
    v_index = _idx_any(v_index, "v_index", dev, P, K, 2)
    with _Timed("adil_grad_dict_step", dev, _n_launches_with_reduction(B, P, K) if (want_dv and not keep_partials) else 1):
        rc = _host_index_retry(
            lambda ix: _lib.lib().adil_grad_dict_step(_ptr(D2), _ptr(m), _ptr(s),
                                                      _ptr(dvb) if (want_dv and not keep_partials) else None, _ptr(g),
                                                      _ptr(v), _ptr(ix), B, P, K, C, P // C, _host3(std, C),
                                                      ctypes.byref(hp), int(atoms_mode), flags, ctypes.byref(nslabs),
                                                      _ptr(scratch), nbytes, _stream(dev)), v_index, dev)
    _lib.check(rc, "adil_grad_dict_step")
    if not want_dv:
        return None
    return CodePartials(scratch, nslabs.value, B, K) if keep_partials else dvb


def dict_step(D2, m, s, dD2, hp, atoms_mode=ATOMS_CLAMP1):
    """AdamW + clamp on (a slice of) the dictionary (adil.py:186,188); all four tensors have the same numel."""
    D2, m, s, dD2 = _f32(D2, "D2"), _f32(m, "m"), _f32(s, "s"), _f32(dD2, "dD2")
    n = D2.numel()
    if not (m.numel() == n and s.numel() == n and dD2.numel() == n):
        raise ValueError("dict_step: size mismatch")
    with _Timed("adil_dict_step", D2.device):
        rc = _lib.lib().adil_dict_step(_ptr(D2), _ptr(m), _ptr(s), _ptr(dD2), n, ctypes.byref(hp), int(atoms_mode),
                                       _stream(D2.device))
    _lib.check(rc, "adil_dict_step")


def dict_step_atoms(D2, dD2, atoms_mode, hp=None, m=None, s=None, step=0.0):
    """Dictionary step with any per-atom projection (adil_regularized.py:27-28,141-146,283-285): AdamW (hp, m, s) or,
    with hp None, the plain gradient step D2 -= step * dD2; then NONE / CLAMP1 / L2BALL / L2SPHERE over the atoms."""
    D2, dD2 = _f32(D2, "D2"), _f32(dD2, "dD2")
    K = D2.shape[-1]
    P = D2.numel() // K
    if dD2.numel() != D2.numel():
        raise ValueError("dict_step_atoms: size mismatch")
    if hp is not None:
        m, s = _f32(m, "m"), _f32(s, "s")
    nbytes = _lib.lib().adil_project_atoms_scratch_bytes(K)
    scratch = _get_scratch(D2.device, nbytes, "atoms")
    with _Timed("adil_dict_step_atoms", D2.device, 3):
        rc = _lib.lib().adil_dict_step_atoms(_ptr(D2), _ptr(m) if hp is not None else None,
                                             _ptr(s) if hp is not None else None, _ptr(dD2), P, K,
                                             ctypes.byref(hp) if hp is not None else None, float(step), int(atoms_mode),
                                             _ptr(scratch), _stream(D2.device))
    _lib.check(rc, "adil_dict_step_atoms")
    return D2


def dict_step_peer(D_ptrs, dD_ptrs, m, s, slice_begin, slice_elems, rank, hp, atoms_mode=ATOMS_CLAMP1, device=None,
                   D_mc=0, dD_mc=0):
    """Fused reduce-scatter + AdamW + clamp + all-gather over peer-mapped buffers (adil_dict_step_peer).  D_ptrs / dD_ptrs:
    device pointers (ints) of every rank's padded dictionary / gradient buffer as mapped into this process.  D_mc /
    dD_mc: multicast addresses of the two buffers (0: none) -- then the reduction happens inside the NVSwitch
    (multimem.ld_reduce) and the result is replicated by it (multimem.st)."""
    m, s = _f32(m, "m"), _f32(s, "s")
    world = len(D_ptrs)
    if len(dD_ptrs) != world:
        raise ValueError("dict_step_peer: %d dictionary pointers, %d gradient pointers" % (world, len(dD_ptrs)))
    arr_D = (ctypes.c_void_p * world)(*[int(p) for p in D_ptrs])
    arr_g = (ctypes.c_void_p * world)(*[int(p) for p in dD_ptrs])
    dev = m.device if device is None else device
    with _Timed("adil_dict_step_peer", dev):
        rc = _lib.lib().adil_dict_step_peer(arr_D, arr_g, _ptr(m), _ptr(s), int(slice_begin), int(slice_elems), int(rank),
                                            world, ctypes.byref(hp), int(atoms_mode),
                                            ctypes.c_void_p(int(D_mc)) if D_mc else None,
                                            ctypes.c_void_p(int(dD_mc)) if dD_mc else None, _stream(dev))
    _lib.check(rc, "adil_dict_step_peer")


def code_prox_step(v, dvb, v_index, step, rows_mode=ROWS_SOFTSHRINK, radius=0.0):
    """v[v_index] = prox(v[v_index] - step * dvb) on the rows of one minibatch (adil_regularized.py:304,414-416)."""
    v, dvb = _f32(v, "v"), _f32(dvb, "dvb")
    N, K = v.shape
    v_index = _idx(v_index, "v_index", v.device)
    B = dvb.shape[0]
    if v_index is not None and v_index.numel() != B:
        raise ValueError("code_prox_step: v_index has %d entries, dvb %d rows" % (v_index.numel(), B))
    rc = _lib.lib().adil_code_prox_step(_ptr(v), _ptr(dvb), _ptr(v_index), B, N, K, float(step), int(rows_mode),
                                        float(radius), _stream(v.device))
    _lib.check(rc, "adil_code_prox_step")
    return v


def code_step(v, m, s, dvb, v_index, hp, rows_mode=ROWS_L1BALL, radius=0.0):
    """AdamW on every row of v (zero gradient outside the batch) + row projection (adil.py:186-187).  `dvb`: [B,K]
    code gradient, a `CodePartials` handle of a backward call (reduced here), or None (zero gradient)."""
    v, m, s = _f32(v, "v"), _f32(m, "m"), _f32(s, "s")
    N, K = v.shape
    v_index = _idx(v_index, "v_index", v.device)
    partial, nslabs = None, 0
    if isinstance(dvb, CodePartials):
        if dvb.K != K:
            raise ValueError("code_step: partials have K=%d, v has K=%d" % (dvb.K, K))
        partial, nslabs, B, dvb = dvb.buf, dvb.nslabs, dvb.B, None
    else:
        dvb = _f32(dvb, "dvb", allow_none=True)
        B = 0 if dvb is None else dvb.shape[0]
    if (dvb is not None or partial is not None) and v_index is not None and v_index.numel() != B:
        raise ValueError("code_step: v_index has %d entries, the gradient %d rows" % (v_index.numel(), B))
    with _Timed("adil_code_step", v.device):
        rc = _lib.lib().adil_code_step(_ptr(v), _ptr(m), _ptr(s), _ptr(dvb), _ptr(v_index), B, N, K, ctypes.byref(hp),
                                       int(rows_mode), float(radius), _ptr(partial), int(nslabs), _stream(v.device))
    _lib.check(rc, "adil_code_step")


def project_rows(v, rows_mode, radius):
    """In-place row projection (adil.py:625-633, utils.py:21-41,159-161)."""
    v = _f32(v, "v")
    N, K = v.shape[0], v[0].numel()
    rc = _lib.lib().adil_project_rows(_ptr(v), N, K, int(rows_mode), float(radius), _stream(v.device))
    _lib.check(rc, "adil_project_rows")
    return v


def project_atoms(D, atoms_mode):
    """In-place per-atom projection of D[..., K] (adil.py:635-642, utils.py:44-57).  The l1ball mode follows
    utils.py:23,56: atom k is viewed as [D.shape[0], -1] and every row is projected onto the unit l1 ball."""
    D = _f32(D, "D")
    K = D.shape[-1]
    P = D.numel() // K
    C = D.shape[0] if (D.dim() >= 3 and atoms_mode == ATOMS_L1BALL) else 1
    nbytes = _lib.lib().adil_project_atoms_scratch_bytes(K)
    scratch = _get_scratch(D.device, nbytes, "atoms")
    rc = _lib.lib().adil_project_atoms(_ptr(D), P, K, int(C), int(atoms_mode), _ptr(scratch), _stream(D.device))
    _lib.check(rc, "adil_project_atoms")
    return D


def adamw_clamp(p, m, s, g, hp, bound=0.0):
    """Elementwise AdamW + clamp(+-bound): z update of forward_supervised_DDrague (adil.py:554-555)."""
    p, m, s, g = _f32(p, "p"), _f32(m, "m"), _f32(s, "s"), _f32(g, "g")
    n = p.numel()
    rc = _lib.lib().adil_adamw_clamp(_ptr(p), _ptr(m), _ptr(s), _ptr(g), n, ctypes.byref(hp), float(bound),
                                     _stream(p.device))
    _lib.check(rc, "adil_adamw_clamp")


def image_errors(adv, clean):
    """Per-image sum_p (adv-clean)^2, sum_p clean^2 and max_p |adv-clean| in one pass (the reductions of
    performance.py:249-266 and adil.py:503-505).  adv, clean: [n, ...] CUDA fp32; returns three [n] tensors."""
    a, c = _f32(adv.detach().contiguous(), "adv"), _f32(clean.detach().contiguous(), "clean")
    if a.shape != c.shape:
        raise ValueError("image_errors: adv %s and clean %s differ in shape" % (tuple(a.shape), tuple(c.shape)))
    n = a.shape[0]
    P = a.numel() // max(n, 1)
    out = torch.empty(3, n, device=a.device, dtype=torch.float32)
    if n == 0:
        return out[0], out[1], out[2]
    nbytes = _lib.lib().adil_image_errors_scratch_bytes(n)
    scratch = _get_scratch(a.device, nbytes, "image_errors")
    rc = _lib.lib().adil_image_errors(_ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(a), _ptr(c), n, P,
                                      _ptr(scratch), nbytes, _stream(a.device))
    _lib.check(rc, "adil_image_errors")
    return out[0], out[1], out[2]


class SynthFunction(torch.autograd.Function):
    """autograd bridge: forward = adil_synth, backward = adil_grad.  Lets user code differentiate through the
    fused synthesis like through adil.py:24-27 (the drivers in adil.py call the kernels directly instead)."""

    @staticmethod
    def forward(ctx, D, v, x, v_index, mean, std, flags):
        K = D.shape[-1]
        D2 = D.reshape(-1, K)
        B = v_index.numel()
        out, _ = synth(D2, v, v_index, x.reshape(B, -1), None, mean, std, 0.0, flags)
        ctx.save_for_backward(D, v, v_index)
        ctx.std = std if (flags & SYNTH_NORMALIZE) else None
        return out.reshape(x.shape)

    @staticmethod
    def backward(ctx, gout):
        D, v, v_index = ctx.saved_tensors
        K = D.shape[-1]
        B = v_index.numel()
        need_D, need_v = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_D or need_v):
            return None, None, None, None, None, None, None
        dD2, dvb = grad(gout.contiguous().reshape(B, -1), D.reshape(-1, K), v, v_index, ctx.std, want_dD=need_D,
                        want_dv=need_v)
        gv = None
        if need_v:
            gv = torch.zeros_like(v)
            gv.index_add_(0, v_index, dvb)
        return (dD2.reshape(D.shape) if need_D else None), gv, None, None, None, None, None
