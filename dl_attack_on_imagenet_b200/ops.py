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
ATOMS_NONE, ATOMS_CLAMP1, ATOMS_L2BALL, ATOMS_L2SPHERE = 0, 1, 2, 3
IMPL_AUTO, IMPL_FMA, IMPL_TC = 0, 1, 2

_scratch = {}


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
    key = (device.index, tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def adamw_params(step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2):
    """torch.optim.AdamW defaults (adil.py:154); `step` is the 1-based count of this update."""
    return AdamwParams(float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step))


def set_impl(impl):
    _lib.check(_lib.lib().adil_set_impl(int(impl)), "adil_set_impl")


def get_impl():
    return _lib.lib().adil_get_impl()


def tc_supported(B, P, K):
    return bool(_lib.lib().adil_tc_supported(int(B), int(P), int(K)))


def device_info():
    sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.lib().adil_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi)), "adil_device_info")
    return sm.value, ma.value, mi.value


def synth(D2, v, v_index=None, x=None, x_index=None, mean=None, std=None, eps=0.0, flags=0, out=None,
          delta_out=None, want_out=True, n_channels=None):
    """out[b] = f(x[x_index[b]] + D2 @ v[v_index[b]])  -- adil.py:25-26 fused with Normalize / clamps.

    D2: [P,K]; v: [N,K]; x: [B,P] (or [Nx,P] with x_index).  Returns (out, delta_out) ([B,P] or None each)."""
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
    C = n_channels if n_channels is not None else (len(mean) if mean is not None else 1)
    hw = P // C
    mean_h, std_h = _host3(mean, C), _host3(std, C)
    rc = _lib.lib().adil_synth(_ptr(out), _ptr(delta_out), _ptr(x), _ptr(x_index), _ptr(D2), _ptr(v), _ptr(v_index),
                               B, P, K, C, hw, mean_h, std_h, float(eps), int(flags), _stream(dev))
    _lib.check(rc, "adil_synth")
    return out, delta_out


def _grad_scratch(dev, B, K):
    nbytes = _lib.lib().adil_grad_scratch_bytes(int(B), int(K))
    return _get_scratch(dev, nbytes, "grad"), nbytes


def grad(g, D2, v, v_index=None, std=None, want_dD=True, want_dv=True, dD2=None, dvb=None):
    """Backward contractions (adil.py:185): returns (dD2 [P,K] or None, dvb [B,K] or None)."""
    g = _f32(g, "g")
    D2 = _f32(D2, "D2")
    v = _f32(v, "v")
    dev = D2.device
    P, K = D2.shape
    v_index = _idx_any(v_index, "v_index", dev, P, K, 2)
    B = v_index.numel() if v_index is not None else v.shape[0]
    if g.numel() != B * P:
        raise ValueError("g has %d elements, expected B*P = %d" % (g.numel(), B * P))
    if want_dD and dD2 is None:
        dD2 = torch.empty((P, K), dtype=torch.float32, device=dev)
    if want_dv and dvb is None:
        dvb = torch.empty((B, K), dtype=torch.float32, device=dev)
    C = len(std) if std is not None else 1
    scratch, nbytes = _grad_scratch(dev, B, K)
    rc = _lib.lib().adil_grad(_ptr(dD2) if want_dD else None, _ptr(dvb) if want_dv else None, _ptr(g), _ptr(D2),
                              _ptr(v), _ptr(v_index), B, P, K, C, P // C, _host3(std, C), _ptr(scratch), nbytes,
                              _stream(dev))
    _lib.check(rc, "adil_grad")
    return (dD2 if want_dD else None), (dvb if want_dv else None)


def grad_dict_step(D2, m, s, g, v, v_index, hp, std=None, atoms_mode=ATOMS_CLAMP1, want_dv=True, dvb=None):
    """Single-GPU fused backward + dictionary AdamW + clamp (adil.py:185-188 for D).  Returns dvb [B,K] or None."""
    D2 = _f32(D2, "D2")
    m = _f32(m, "m")
    s = _f32(s, "s")
    g = _f32(g, "g")
    v = _f32(v, "v")
    dev = D2.device
    P, K = D2.shape
    v_index = _idx_any(v_index, "v_index", dev, P, K, 2)
    B = v_index.numel() if v_index is not None else v.shape[0]
    if g.numel() != B * P:
        raise ValueError("g has %d elements, expected B*P = %d" % (g.numel(), B * P))
    if want_dv and dvb is None:
        dvb = torch.empty((B, K), dtype=torch.float32, device=dev)
    C = len(std) if std is not None else 1
    scratch, nbytes = _grad_scratch(dev, B, K)
    rc = _lib.lib().adil_grad_dict_step(_ptr(D2), _ptr(m), _ptr(s), _ptr(dvb) if want_dv else None, _ptr(g), _ptr(v),
                                        _ptr(v_index), B, P, K, C, P // C, _host3(std, C), ctypes.byref(hp),
                                        int(atoms_mode), _ptr(scratch), nbytes, _stream(dev))
    _lib.check(rc, "adil_grad_dict_step")
    return dvb if want_dv else None


def dict_step(D2, m, s, dD2, hp, atoms_mode=ATOMS_CLAMP1):
    """AdamW + clamp on (a slice of) the dictionary (adil.py:186,188); all four tensors have the same numel."""
    D2, m, s, dD2 = _f32(D2, "D2"), _f32(m, "m"), _f32(s, "s"), _f32(dD2, "dD2")
    n = D2.numel()
    if not (m.numel() == n and s.numel() == n and dD2.numel() == n):
        raise ValueError("dict_step: size mismatch")
    rc = _lib.lib().adil_dict_step(_ptr(D2), _ptr(m), _ptr(s), _ptr(dD2), n, ctypes.byref(hp), int(atoms_mode),
                                   _stream(D2.device))
    _lib.check(rc, "adil_dict_step")


def code_step(v, m, s, dvb, v_index, hp, rows_mode=ROWS_L1BALL, radius=0.0):
    """AdamW on every row of v (zero gradient outside the batch) + row projection (adil.py:186-187)."""
    v, m, s = _f32(v, "v"), _f32(m, "m"), _f32(s, "s")
    dvb = _f32(dvb, "dvb", allow_none=True)
    N, K = v.shape
    v_index = _idx(v_index, "v_index", v.device)
    B = 0 if dvb is None else dvb.shape[0]
    if dvb is not None and v_index is not None and v_index.numel() != B:
        raise ValueError("code_step: v_index has %d entries, dvb %d rows" % (v_index.numel(), B))
    rc = _lib.lib().adil_code_step(_ptr(v), _ptr(m), _ptr(s), _ptr(dvb), _ptr(v_index), B, N, K, ctypes.byref(hp),
                                   int(rows_mode), float(radius), _stream(v.device))
    _lib.check(rc, "adil_code_step")


def project_rows(v, rows_mode, radius):
    """In-place row projection (adil.py:625-633, utils.py:21-41,159-161)."""
    v = _f32(v, "v")
    N, K = v.shape[0], v[0].numel()
    rc = _lib.lib().adil_project_rows(_ptr(v), N, K, int(rows_mode), float(radius), _stream(v.device))
    _lib.check(rc, "adil_project_rows")
    return v


def project_atoms(D, atoms_mode):
    """In-place per-atom projection of D[..., K] (adil.py:635-642, utils.py:44-57)."""
    D = _f32(D, "D")
    K = D.shape[-1]
    P = D.numel() // K
    nbytes = _lib.lib().adil_project_atoms_scratch_bytes(K)
    scratch = _get_scratch(D.device, nbytes, "atoms")
    rc = _lib.lib().adil_project_atoms(_ptr(D), P, K, int(atoms_mode), _ptr(scratch), _stream(D.device))
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
        dD2, dvb = grad(gout.contiguous().reshape(B, -1), D.reshape(-1, K), v, v_index, ctx.std)
        gv = torch.zeros_like(v)
        gv.index_add_(0, v_index, dvb)
        return dD2.reshape(D.shape), gv, None, None, None, None, None
