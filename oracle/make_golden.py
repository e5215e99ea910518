"""TEST INFRASTRUCTURE -- generates tests/golden/adil_reference_golden.npz by running the UNMODIFIED
reference (`/root/reference`, imported through oracle/ref_shim.py) on tiny seeded problems.

Run in the build container only (the reference tree does not travel to the GPU box):

    python oracle/make_golden.py

The fixtures pin `oracle/adil_oracle.py` (tests/test_oracle_golden.py) and, through it, the CUDA
kernels.  Every array is fp32/int64 and a few KB; all inputs needed to replay a case are stored
next to the reference's outputs, so the tests never need the reference itself.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle.adil_oracle import IndexedTensorDataset, tiny_classifier  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "adil_reference_golden.npz")

# tiny problem used for all driver-level cases
C, H, W, K, N, NVAL, B = 3, 8, 8, 6, 10, 3, 4
EPS = 8.0 / 255.0


def tiny_data():
    g1 = torch.Generator().manual_seed(1)
    g2 = torch.Generator().manual_seed(2)
    xtr = torch.rand(N, C, H, W, generator=g1)
    ytr = torch.randint(0, 10, (N,), generator=g1)
    xva = torch.rand(NVAL, C, H, W, generator=g2)
    yva = torch.randint(0, 10, (NVAL,), generator=g2)
    return xtr, ytr, xva, yva


class QuickDataset(torch.utils.data.Dataset):
    """(x, y) pairs, like the reference's QuickAttackDataset (utils.py:177-186)."""

    def __init__(self, images, labels):
        self.images, self.labels = images, labels

    def __len__(self):
        return len(self.images)

    def __getitem__(self, item):
        return self.images[item], self.labels[item]


def main():
    torch.set_num_threads(1)
    ref = ref_shim.load_reference()
    ru = ref_shim.load_reference_utils()
    out = {}

    # ---- known-answer vectors and random cases for the projections ------------------
    x_kat = torch.tensor([[.5, -.3, .1], [.1, .1, -.1], [0, 0, 0], [.25, -.25, 0], [2, 0, 0], [-.4, .4, .4]])
    out["l1_kat_in"] = x_kat.numpy()
    out["l1_kat_out"] = ru.project_onto_l1_ball(x_kat.clone(), eps=0.5).numpy()
    g = torch.Generator().manual_seed(7)
    xr = torch.randn(37, 13, generator=g) * 0.05
    xr[3] = 0
    xr[5, :] = 0.01                      # ties
    xr[6, 2:] = 0                        # sparse
    out["l1_rand_in"] = xr.numpy()
    out["l1_rand_out_eps"] = ru.project_onto_l1_ball(xr.clone(), eps=EPS).numpy()
    out["l1_rand_out_1"] = ru.project_onto_l1_ball(xr.clone(), eps=1.0).numpy()
    xk200 = torch.rand(9, 200, generator=g)
    out["l1_k200_in"] = xk200.numpy()
    out["l1_k200_out"] = ru.project_onto_l1_ball(xk200.clone(), eps=EPS).numpy()

    model = tiny_classifier(seed=0)
    xtr, ytr, xva, yva = tiny_data()

    class Holder(ref.ADIL):                                   # ADIL without the fit side effect of __init__
        def __init__(self, model, **kw):
            ref.Attack.__init__(self, "ADIL", model.eval())
            for k_, v_ in kw.items():
                setattr(self, k_, v_)

    h_l2 = Holder(model, norm='l2', eps=0.5, n_atoms=3)
    out["l2rows_kat_out"] = h_l2.projection_v(x_kat.clone()).numpy()
    h_l2b = Holder(model, norm='l2', eps=EPS, n_atoms=13)
    out["l2rows_rand_out"] = h_l2b.projection_v(xr.clone()).numpy()

    d_kat = torch.arange(24, dtype=torch.float32).reshape(1, 2, 3, 4) / 10 - 1
    out["atoms_kat_in"] = d_kat.numpy()
    out["atoms_kat_l2ball"] = ru.constraint_dict(d_kat.clone(), 'l2ball').numpy()
    d_r = torch.randn(3, 5, 4, 7, generator=g) * 0.2
    d_r[..., 2] *= 0.01                                       # an atom inside the unit ball
    out["atoms_rand_in"] = d_r.numpy()
    out["atoms_rand_l2ball"] = ru.constraint_dict(d_r.clone(), 'l2ball').numpy()
    out["atoms_rand_l2sphere"] = ru.constraint_dict(d_r.clone(), 'l2sphere').numpy()
    out["atoms_rand_l1ball"] = ru.constraint_dict(d_r.clone(), 'l1ball').numpy()

    s_in = torch.tensor([.25, -.05, .1, -.3])
    out["shrink_in"] = s_in.numpy()
    out["shrink_out"] = ru.get_prox_l1(0.1)(s_in).numpy()

    logits = torch.tensor([[1., 5., 2.], [-3., -1., -2.], [100., 0., 0.]])
    lab = torch.tensor([1, 1, 0])
    h_f = Holder(model, kappa=50)
    out["floss_logits"] = logits.numpy()
    out["floss_labels"] = lab.numpy()
    out["floss_out"] = h_f.f_loss(logits, lab).numpy()
    h_f._targeted = True
    out["floss_out_targeted"] = h_f.f_loss(logits, lab).numpy()

    # ---- teacher-forced steps: reference Attack_dict_model + AdamW + update_v + update_d -------------------
    torch.manual_seed(11)
    D0 = -1 + 2 * torch.rand(C, H, W, K)
    v0 = ru.project_onto_l1_ball(torch.rand(N, K), eps=EPS)
    adm = ref.Attack_dict_model(D0.clone(), v0.clone(), EPS)
    opt = torch.optim.AdamW(adm.parameters(), lr=0.01)
    out["tf_D0"], out["tf_v0"] = D0.numpy(), v0.numpy()
    crit = torch.nn.CrossEntropyLoss(reduction='sum')
    gen = torch.Generator().manual_seed(5)
    for step in range(3):
        idx = torch.randperm(N, generator=gen)[:B]
        x = xtr[idx]
        label = model(x).argmax(dim=-1)
        opt.zero_grad()
        xadv = None

        def hooked(inp):                                       # capture d loss / d (x + dv) at the model input
            nonlocal xadv
            xadv = inp
            xadv.retain_grad()
            return model(inp)
        output = adm(x, idx, hooked)
        loss = -crit(output, label)
        loss.backward()
        out["tf_idx_%d" % step] = idx.numpy()
        out["tf_gin_%d" % step] = xadv.grad.detach().clone().numpy()      # gradient w.r.t. the UN-normalised input
        out["tf_dD_%d" % step] = adm.d.grad.detach().clone().numpy()
        out["tf_dv_%d" % step] = adm.v.grad.detach().clone().numpy()
        out["tf_xadv_%d" % step] = xadv.detach().clone().numpy()
        opt.step()
        adm.update_v()
        adm.update_d()
        out["tf_D_%d" % step] = adm.d.data.clone().numpy()
        out["tf_v_%d" % step] = adm.v.data.clone().numpy()

    # ---- whole-fit runs through the reference constructor ---------------------------------------------------
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    os.makedirs("trained_dicts", exist_ok=True)
    try:
        def fit(tag, seed, **kw):
            torch.manual_seed(seed)
            tr = IndexedTensorDataset(xtr, ytr)
            va = IndexedTensorDataset(xva, yva)
            atk = ref.ADIL(model, eps=EPS, n_atoms=K, batch_size=B, data_train=tr, data_val=va,
                           model_name=tag, **kw)
            rl = torch.load(atk.model_file, weights_only=False)
            out[tag + "_D"] = rl[0].numpy()
            out[tag + "_v"] = rl[1].numpy()
            out[tag + "_loss"] = np.asarray(rl[2], dtype=np.float64)
            out[tag + "_fool"] = np.asarray(rl[3], dtype=np.float64)
            out[tag + "_valfool"] = np.asarray(float(rl[4]), dtype=np.float64)
            return atk

        atk_a = fit("fit_gd_ce", 1234, steps=3, norm='linf', loss='ce', method='gd')
        fit("fit_gd_logits", 1235, steps=3, norm='linf', loss='logits', method='gd', kappa=50)
        fit("fit_gd_l2", 1236, steps=2, norm='l2', loss='ce', method='gd')
        fit("fit_alter_ce", 1237, steps=4, steps_in=2, norm='linf', loss='ce', method='alter')

        # ---- inference paths on the gd dictionary ---------------------------------------------------------
        Dfit = torch.from_numpy(out["fit_gd_ce_D"])
        atk_a.steps_inference = 5
        adv = atk_a(xva, yva)                                  # forward -> forward_supervised_DDrague
        out["ddrague_adv"] = adv.detach().numpy()
        adv2 = atk_a.forward_supervised_AdamW(xva, yva, Dfit.clone(), 'eval')
        out["coder_adv"] = adv2.detach().numpy()
        out["coder_fooled"] = np.asarray(int(atk_a.forward_supervised_AdamW(xva, yva, Dfit.clone(), 'train')))
        atk_a.attack = 'unsupervised'
        atk_a.trials = 3
        torch.manual_seed(99)
        advu, dvn = atk_a(xva, yva)
        out["unsup_adv"] = advu.detach().numpy()
        out["unsup_dvnorm"] = np.asarray(dvn, dtype=np.float64)
        torch.manual_seed(98)
        out["sample_sphere_linf"] = atk_a.sample_sphere(5).numpy()
        atk_a.norm = 'l2'
        torch.manual_seed(97)
        out["sample_sphere_l2"] = atk_a.sample_sphere(5).numpy()
    finally:
        os.chdir(cwd)

    # ---- regularised variant: SADiL through the reference's own function (adil_regularized.py:200-312) ---------------
    import importlib
    reg = importlib.import_module("attacks.attacks_classes.adil_regularized")
    tmp2 = tempfile.mkdtemp()
    for tag, kw in (("sadil_untargeted", dict(targeted=False, batchsize=4, lambdaCoding=0.01, l2_fool=0.5, stepsize=0.05,
                                               n_atom=K, dict_set='l2ball')),
                    ("sadil_targeted", dict(targeted=True, batchsize=3, lambdaCoding=0.02, l2_fool=2.0, stepsize=0.02,
                                            n_atom=5, dict_set='l2sphere'))):
        torch.manual_seed(4321)
        state = torch.get_rng_state()
        D0 = ru.constraint_dict(torch.randn(3, H, W, kw["n_atom"]), constr_set=kw["dict_set"])   # the draw sadil makes
        torch.set_rng_state(state)
        Dr, vr, _ = reg.sadil(QuickDataset(xtr, ytr), model, nepochs=3, device=torch.device("cpu"),
                              model_file=os.path.join(tmp2, tag + ".bin"), **kw)
        _, loss_r = torch.load(os.path.join(tmp2, tag + ".bin"), weights_only=False)
        out[tag + "_D0"] = D0.numpy()
        out[tag + "_D"] = Dr.detach().numpy()
        out[tag + "_v"] = vr.detach().numpy()
        out[tag + "_loss"] = np.asarray(loss_r, dtype=np.float64)

    out["meta_torch_version"] = np.asarray(torch.__version__)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
