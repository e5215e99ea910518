"""Pins the oracle to the UNMODIFIED reference at real image size: BASELINE.json configs[0] (ResNet-18 random-init,
32 synthetic 3x224x224 images, 10 atoms, l_inf 8/255, batch 32, 20 iterations, CPU).  The reference trajectory is
the fixture tests/golden/adil_reference_imagenet.npz written by oracle/make_golden_imagenet.py in the build
container (loss/img -1.727425, -1.747301, -1.758778 ... -1.958733)."""
import os

import numpy as np
import pytest
import torch

from oracle import adil_oracle as O
from dl_attack_on_imagenet_b200.data import build_classifier, synthetic_images

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "adil_reference_imagenet.npz")
EPS = 8.0 / 255.0


def checksum(t):
    t = t.detach().double()
    return np.asarray([t.sum().item(), t.abs().sum().item()])


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN, allow_pickle=False))


def test_fixture_holds_the_survey_trajectory(gold):
    loss = gold["cfg1_loss"]
    assert len(loss) == 20
    assert np.allclose(loss[:3], [-1.727425, -1.747301, -1.758778], atol=2e-6) and abs(loss[-1] - (-1.958733)) < 2e-6
    fool = gold["cfg1_fool"]
    assert (fool[:18] == 0).all() and np.allclose(fool[18:], [0.03125, 0.0625])
    for name in ("fr_resnet18", "fr_vgg11", "fr_densenet121"):
        n_img, K, B, steps = gold[name + "_meta"][:4]
        assert (n_img, K, B) == (200, 10, 100) and len(gold[name + "_fool"]) == steps


def test_oracle_reproduces_the_reference_on_config_1(gold):
    """Free-running oracle == free-running reference at 3x224x224 through a real ResNet-18: same RNG draws (initial
    state, shuffling, the iterated-but-stubbed validation loader), same arithmetic -- per-step checksums of D and v and
    the loss trajectory agree to fp32 rounding of the classifier (bit-exact on the build container's CPU)."""
    n_threads = torch.get_num_threads()
    torch.set_num_threads(os.cpu_count() or 1)
    try:
        model = build_classifier('resnet18', seed=0)
        x, y = synthetic_images(32, seed=1)
        xv, yv = synthetic_images(1, seed=2)
        tr, va = O.IndexedTensorDataset(x, y), O.IndexedTensorDataset(xv, yv)
        torch.manual_seed(1234)
        st = O.init_state(3, 224, 224, 32, 10, EPS)
        assert np.array_equal(np.concatenate([checksum(st.D2), checksum(st.v)]), gold["cfg1_init"])
        steps = []

        def on_step(st_, index, xb, xin, g, lval, phase):
            if phase == 'after':
                steps.append(np.concatenate([checksum(st_.D2), checksum(st_.v)]))
        st, loss, fool, _ = O.learn_dictionary_a(model, tr, EPS, 20, 10, 32, state=st, val=va, val_coder=False,
                                                 on_step=on_step)
    finally:
        torch.set_num_threads(n_threads)
    ref_steps = gold["cfg1_steps"]
    assert len(steps) == len(ref_steps) == 20
    # checksums: sum(D), sum|D| over 1.5 M entries, sum(v), sum|v| over 320 -- relative agreement
    got = np.stack(steps)
    assert np.abs(got - ref_steps).max() <= 1e-6 * np.abs(ref_steps).max()
    assert np.abs(np.asarray(loss) - gold["cfg1_loss"]).max() <= 1e-6
    assert np.array_equal(np.asarray(fool), gold["cfg1_fool"])
    pix = torch.randperm(3 * 224 * 224, generator=torch.Generator().manual_seed(5))[:256]
    assert np.abs(st.D2[pix].numpy() - gold["cfg1_D_sub"]).max() <= 1e-6
    assert np.abs(st.v.numpy() - gold["cfg1_v"]).max() <= 1e-7
