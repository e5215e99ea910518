"""Host logic of the regularised drivers (dl_attack_on_imagenet_b200.adil_regularized) on the CPU box.

The drivers have no CPU path: every arithmetic step is a C-ABI call.  Here those calls (`ops.synth`, `ops.grad`,
`ops.code_prox_step`, `ops.dict_step_atoms`, `ops.project_atoms`) are replaced by the test suite's CPU checker
(oracle/adil_oracle.py primitives), so that what runs is the drivers' own control flow -- gradient accumulation across
backward passes, line searches / backtracking tests, step-size adaptation, stopping rules, saved files -- against the
outputs of the reference's own functions (tests/golden/*_reference_golden.npz).  The kernels themselves are checked on the
GPU (tests/test_adil_gpu.py runs the same cases through the real library).
"""
import os

import numpy as np
import pytest
import torch

from oracle import adil_oracle as O
from oracle.make_golden import tiny_data

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class _CpuOps(object):
    """the slice of dl_attack_on_imagenet_b200.ops the regularised drivers use, emulated with the CPU checker"""
    SYNTH_NORMALIZE, ROWS_SOFTSHRINK = 1, 3
    ATOMS_NONE, ATOMS_CLAMP1, ATOMS_L2BALL, ATOMS_L2SPHERE, ATOMS_L1BALL = 0, 1, 2, 3, 4

    @staticmethod
    def synth(D2, v, v_index=None, x=None, mean=None, std=None, flags=0, delta_out=None, n_channels=None, **kw):
        out, delta = O.synth(x, D2, v, v_index, mean=mean, std=std, flags=flags)
        if delta_out is not None:
            delta_out.copy_(delta)
        return out, delta_out

    @staticmethod
    def grad(g, D2, v, v_index, std, want_dD=True, want_dv=True, dD2=None, dvb=None, accumulate=False, delta=None,
             l2_coef=0.0, **kw):
        vb = v[v_index] if v_index is not None else v
        gx = g if std is None else g / O.channel_vec(std, len(std), D2.shape[0] // len(std), g.dtype)
        if delta is not None and l2_coef != 0.0:
            gx = gx + l2_coef * delta                      # the penalty 0.5 * l2 * ||D v||^2 seen from the perturbation
        rD = None
        if want_dD:
            rD = gx.t() @ vb
            if dD2 is not None:
                rD = dD2.add_(rD) if accumulate else dD2.copy_(rD)
        return rD, (gx @ D2 if want_dv else None)

    @staticmethod
    def code_prox_step(v, dvb, v_index, step, rows_mode, radius):
        assert rows_mode == _CpuOps.ROWS_SOFTSHRINK
        v[v_index] = O.softshrink(v[v_index] - step * dvb, radius)
        return v

    @staticmethod
    def dict_step_atoms(D2, dD2, atoms_mode, step=0.0, **kw):
        D2.sub_(step * dD2)
        if atoms_mode != _CpuOps.ATOMS_NONE:
            _CpuOps.project_atoms(D2.view(3, -1, 1, D2.shape[-1]), atoms_mode)
        return D2

    @staticmethod
    def project_atoms(D, atoms_mode):
        D.copy_(O.project_atoms(D.clone(), atoms_mode))
        return D


@pytest.fixture()
def reg(monkeypatch):
    import dl_attack_on_imagenet_b200.adil_regularized as reg
    monkeypatch.setattr(reg, "ops", _CpuOps)
    monkeypatch.setattr(reg, "_device_of", lambda model: torch.device("cpu"))
    torch.set_num_threads(1)
    return reg


def _data():
    from dl_attack_on_imagenet_b200.utils import QuickAttackDataset
    xtr, ytr, _, _ = tiny_data()
    return QuickAttackDataset(xtr, ytr)


@pytest.mark.parametrize("tag,kw,ended", [
    ("lcv_untargeted", dict(targeted=False, niter=8, lambda_l1=0.01, lambda_l2=0.5, batch_size=4, step_size=0.05), [0] * 8),
    ("lcv_backtrack", dict(targeted=False, niter=8, lambda_l1=0.05, lambda_l2=0.5, batch_size=4, step_size=10.0), [11, 11]),
    ("lcv_linesearch", dict(targeted=False, niter=8, lambda_l1=0.05, lambda_l2=0.5, batch_size=4, step_size=1.0),
     [5, 0, 3, 2, 5, 0, 2, 0])])
def test_learn_coding_vectors_control_flow(reg, tag, kw, ended):
    g = np.load(os.path.join(GOLD, "lcv_reference_golden.npz"))
    trace = []
    v = reg.learn_coding_vectors(_data(), O.tiny_classifier(seed=0), dictionary=torch.from_numpy(g[tag + "_D"]), trace=trace, **kw)
    assert [t[0] for t in trace] == ended
    assert np.allclose([t[2] for t in trace], g[tag + "_loss"][1:], rtol=1e-6, atol=2e-5)
    vref = torch.from_numpy(g[tag + "_v"])
    assert (v - vref).abs().max() <= 1e-5 * max(1.0, vref.abs().max().item())


@pytest.mark.parametrize("tag,kw,halvings", [
    ("su_untargeted", dict(targeted=False, nepochs=5, batchsize=4, lambdaCoding=0.01, l2_fool=0.5, stepsize=0.05, n_atom=6,
                           dict_set='l2ball'), [(0, 0), (0, 0), (0, 0), (0, 1), (0, 1)]),
    ("su_targeted", dict(targeted=True, nepochs=5, batchsize=3, lambdaCoding=0.02, l2_fool=2.0, stepsize=0.02, n_atom=5,
                         dict_set='l2sphere'), [(0, 0)] * 5),
    ("su_backtrack", dict(targeted=False, nepochs=5, batchsize=4, lambdaCoding=0.05, l2_fool=0.5, stepsize=2.0, n_atom=6,
                          dict_set='l2ball'), [(5, 3), (0, 3), (0, 2), (0, 2), (0, 0)])])
def test_sadil_updated_control_flow(reg, tmp_path, tag, kw, halvings):
    g = np.load(os.path.join(GOLD, "sadil_updated_reference_golden.npz"))
    trace, path = [], str(tmp_path / "su.bin")
    D, v = reg.sadil_updated(_data(), O.tiny_classifier(seed=0), dictionary=torch.from_numpy(g[tag + "_D0"]), trace=trace,
                             model_file=path, **kw)
    assert [t[:2] for t in trace] == halvings
    assert np.allclose([t[2] for t in trace], g[tag + "_loss"][1:], rtol=1e-5, atol=5e-5)
    assert (D - torch.from_numpy(g[tag + "_D"])).abs().max() <= 1e-4
    assert (v - torch.from_numpy(g[tag + "_v"])).abs().max() <= 1e-4
    Df, label, pred, vf, lossf = torch.load(path, weights_only=True)
    assert torch.equal(Df, D) and torch.equal(vf, v) and len(lossf) == len(g[tag + "_loss"])
    assert label == g[tag + "_label"].tolist() and pred == g[tag + "_pred"].tolist()


@pytest.mark.parametrize("tag,kw,accepted", [
    ("fb_untargeted", dict(targeted=False, niter=8, lambdaCoding=0.01, l2_fool=0.5, batchsize=4, step_size=0.05, n_atom=6,
                           dict_set='l2ball'), [0] * 8),
    ("fb_backtrack", dict(targeted=False, niter=8, lambdaCoding=0.05, l2_fool=0.5, batchsize=4, step_size=10.0, n_atom=6,
                          dict_set='l2ball'), [4, 5, 0, 0, 0, 0, 0, 0])])
def test_full_batch_adil_control_flow(reg, monkeypatch, tag, kw, accepted):
    g = np.load(os.path.join(GOLD, "adil_fb_reference_golden.npz"))
    D0 = torch.from_numpy(g[tag + "_D0"])
    calls, project = {"n": 0}, _CpuOps.project_atoms

    def first_draw(D, mode):                               # the reference learns D from its own draw: hand that draw in
        calls["n"] += 1
        return D.copy_(D0) if calls["n"] == 1 else project(D, mode)

    monkeypatch.setattr(_CpuOps, "project_atoms", staticmethod(first_draw))
    trace = []
    D, v, loss = reg.adil(_data(), O.tiny_classifier(seed=0), trace=trace, **kw)
    assert trace == accepted
    assert np.allclose(loss, g[tag + "_loss"], rtol=1e-5, atol=5e-5)
    assert (D - torch.from_numpy(g[tag + "_D"])).abs().max() <= 1e-4
    assert (v - torch.from_numpy(g[tag + "_v"])).abs().max() <= 1e-4


def test_sadil_control_flow(reg, tmp_path):
    g = np.load(os.path.join(GOLD, "adil_reference_golden.npz"))
    tag = "sadil_untargeted"
    D, v, loss = reg.sadil(_data(), O.tiny_classifier(seed=0), nepochs=3, dictionary=torch.from_numpy(g[tag + "_D0"]),
                           model_file=str(tmp_path / "s.bin"), targeted=False, batchsize=4, lambdaCoding=0.01, l2_fool=0.5,
                           stepsize=0.05, n_atom=6, dict_set='l2ball')
    assert np.allclose(loss, g[tag + "_loss"], rtol=1e-5, atol=2e-5)
    assert (D - torch.from_numpy(g[tag + "_D"])).abs().max() <= 1e-5
    assert (v - torch.from_numpy(g[tag + "_v"])).abs().max() <= 1e-5
