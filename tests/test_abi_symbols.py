"""The C-ABI shared library builds, loads without a GPU and exports every symbol include/adil_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "adil_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adil_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def library():
    from dl_attack_on_imagenet_b200.build import build_library
    path = build_library()
    return ctypes.CDLL(path)


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("adil_synth", "adil_grad", "adil_grad_dict_step", "adil_dict_step", "adil_code_step",
                 "adil_project_rows", "adil_project_atoms", "adil_adamw_clamp", "adil_last_error", "adil_version"):
        assert must in syms


def test_every_declared_symbol_is_exported(library):
    for name in declared_symbols():
        assert hasattr(library, name), name


def test_binding_covers_header_exactly():
    from dl_attack_on_imagenet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_host_only_entry_points(library):
    from dl_attack_on_imagenet_b200 import _lib
    lib = _lib.lib()
    assert lib.adil_version() >= 211
    # tcgen05 coverage by atom count (bit 0: synthesis, bit 1: backward): one window up to 128 atoms, two column windows
    # in one launch up to 256 (K % 8 == 0), the synthesis codes fit tensor memory up to 224 atoms
    for K, want in ((50, 3), (128, 3), (136, 3), (200, 3), (224, 3), (232, 2), (256, 2), (260, 0), (204, 1)):
        assert lib.adil_tc_supported(100, 150528, K) == want, K
    assert lib.adil_grad_max_batch(150528, 50, 50176, 1) == 128 and lib.adil_grad_max_batch(150528, 300, 50176, 0) == 0
    assert lib.adil_grad_scratch_bytes(100, 50) >= 148 * 100 * 50 * 4
    assert lib.adil_grad_scratch_bytes(0, 50) == 0
    assert lib.adil_project_atoms_scratch_bytes(64) > 0
    assert lib.adil_image_errors_scratch_bytes(100) >= 100 * 3 * 4 and lib.adil_image_errors_scratch_bytes(0) == 0
    assert lib.adil_set_impl(7) != 0 and b"bad impl" in lib.adil_last_error()
    assert lib.adil_set_impl(0) == 0


def test_argument_errors_are_reported_without_a_gpu():
    """Shape validation happens before any CUDA call, so it is checkable on the CPU box."""
    from dl_attack_on_imagenet_b200 import _lib
    lib = _lib.lib()
    one = ctypes.c_void_p(16)
    rc = lib.adil_synth(one, None, None, None, one, one, None, None, 4, 10, 3, 1, 10, None, None, 0.0, 0, None)
    assert rc < 0 and b"multiple of 4" in lib.adil_last_error()
    rc = lib.adil_synth(one, None, None, None, one, one, None, None, 4, 12, 300, 1, 12, None, None, 0.0, 0, None)
    assert rc < 0 and b"ADIL_MAX_ATOMS" in lib.adil_last_error()
    rc = lib.adil_grad(None, None, one, one, one, None, 4, 12, 3, 1, 12, None, None, 0.0, 0, None, None, 0, None)
    assert rc < 0 and b"nothing to compute" in lib.adil_last_error()
    rc = lib.adil_grad(one, None, one, one, one, None, 4, 12, 3, 1, 12, None, None, 0.0, 64, None, None, 0, None)
    assert rc < 0 and b"unknown flags" in lib.adil_last_error()
    rc = lib.adil_grad(None, None, one, one, one, None, 4, 12, 3, 1, 12, None, None, 0.0, 2, None, None, 0, None)
    assert rc < 0 and b"nslabs_out" in lib.adil_last_error()
    rc = lib.adil_dict_step_atoms(one, None, None, one, 12, 3, None, 0.1, 4, None, None)
    assert rc < 0 and b"unsupported atoms_mode" in lib.adil_last_error()
    rc = lib.adil_code_prox_step(one, one, None, 4, 8, 3, 0.1, 9, 0.1, None)
    assert rc < 0 and b"bad rows_mode" in lib.adil_last_error()
    rc = lib.adil_code_step(one, one, one, one, None, 4, 8, 3, ctypes.byref(_lib.AdamwParams(0.01, 0.9, 0.999, 1e-8, 0.01, 1)),
                            1, 0.1, one, 3, None)
    assert rc < 0 and b"partial slabs exclude dvb" in lib.adil_last_error()
    rc = lib.adil_project_atoms(one, 12, 3, 5, 4, one, None)
    assert rc < 0 and b"l1ball" in lib.adil_last_error()
    rc = lib.adil_image_errors(one, one, one, one, one, 4, 10, one, 1 << 20, None)
    assert rc < 0 and b"P % 4" in lib.adil_last_error()
    rc = lib.adil_image_errors(one, one, one, one, one, 4, 12, one, 8, None)
    assert rc == -2 and b"scratch" in lib.adil_last_error()


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly: the product path never computes on the host."""
    from dl_attack_on_imagenet_b200 import ops
    D2 = torch.zeros(12, 3)
    v = torch.zeros(2, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.synth(D2, v, x=torch.zeros(2, 12))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.project_rows(v, ops.ROWS_L1BALL, 0.1)
    from dl_attack_on_imagenet_b200 import ADIL
    from oracle.adil_oracle import tiny_classifier
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ADIL(tiny_classifier(), eps=0.03, model_name="cpu_should_fail")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dl_attack_on_imagenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def test_hot_tensor_core_kernels_have_no_register_spills():
    """The benchmarked instantiations of the tcgen05 kernels must not touch local memory: a spill reload queues behind the
    global stores of the AdamW pass and cost the fused step 5-10 us whenever an unrelated edit made ptxas spill one
    loop-carried value.  Checked on the built library with cuobjdump (no GPU needed)."""
    import re
    import shutil
    import subprocess
    from dl_attack_on_imagenet_b200.build import build_library
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = build_library()
    res = subprocess.run([cuobjdump, "-res-usage", lib], capture_output=True, text=True).stdout
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True).stdout
    # (single-window kernels of configs 2-4, the column-window kernels of config 5, the synthesis kernels)
    hot = re.compile(r"(grad_kernelILi64ELb[01]ELb0ELi[34]E|grad_kernelILi32ELb1ELb0ELi3E|grad_kernelILi(64|48|32)ELb[01]ELb1ELi[34]E|"
                     r"synth_kernelILi(64|48)ELb1E)")
    seen = 0
    for m in re.finditer(r"Function (\S+?):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        if hot.search(m.group(1)):
            seen += 1
            assert int(m.group(3)) == 0 and int(m.group(5)) == 0, "%s uses a stack frame / local memory: %s" % (m.group(1), m.group(0))
    assert seen >= 6
    cur, local_ops = None, {}
    for line in sass.splitlines():
        f = re.search(r"Function : (\S+)", line)
        if f:
            cur = f.group(1)
        elif cur and hot.search(cur) and re.search(r"\b(LDL|STL)\b", line):
            local_ops[cur] = local_ops.get(cur, 0) + 1
    assert not local_ops, "local-memory instructions in hot kernels: %s" % local_ops
