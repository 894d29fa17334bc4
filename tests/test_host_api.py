"""CPU-side tests: the host mirror of the reference API, index bookkeeping, the C-ABI
surface, and the basis arithmetic through the host harness.  No GPU needed."""
import ctypes
import os
import re
import subprocess
import warnings

import numpy as np
import pytest

import zernike_oracle as zo
from motif_learn_b200 import _lib
from motif_learn_b200.features import (ZPs, zmoments, nm2j, nm2j_complex, construct_complex_matrix,
                                       construct_real_matrix, construct_rot_maps_matrix, clear_border)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_cuda():
    import torch
    return not torch.cuda.is_available()


# ---- the reference's own test-suite for this path, against our names -------------------------
# (tests/features/test_zmoments.py:5-88 of jiadongdan/motif-learn)
def test_scalar_inputs():
    assert nm2j(0, 0) == 0
    assert nm2j(1, -1) == 1
    assert nm2j(2, 0) == 4
    assert nm2j(3, 1) == 8
    assert nm2j(4, -4) == 10
    assert nm2j(5, 3) == 19


def test_array_inputs():
    np.testing.assert_array_equal(nm2j([0, 1, 2, 2, 3], [0, -1, 0, 2, 3]), [0, 1, 4, 5, 9])


def test_edge_cases():
    assert nm2j(0, 0) == 0
    assert nm2j(1000, 1000) == ((1000 + 2) * 1000 + 1000) // 2


def test_invalid_inputs():
    with pytest.raises(ValueError, match="Radial order `n` must be non-negative."):
        nm2j(-1, 0)
    with pytest.raises(ValueError, match="Azimuthal frequency `m` must satisfy \\|m\\| ≤ n."):
        nm2j(2, 3)
    with pytest.raises(ValueError):
        nm2j(0, 1)
    with pytest.raises(ValueError, match="`n - \\|m\\|` must be even."):
        nm2j(1, 0)
    with pytest.raises(ValueError):
        nm2j(3, 2)
    with pytest.raises(ValueError, match="`n` and `m` must have the same shape."):
        nm2j([1, 2], [0])
    assert nm2j(2.0, 0.0) == 4
    with pytest.raises(ValueError):
        nm2j(2.5, 0)
    with pytest.raises(ValueError):
        nm2j(2, 0.5)


def test_large_array_and_types():
    n_values = np.arange(0, 100) * 2
    m_values = np.where(np.arange(0, 100) % 2 == 0, 0, 1) * 2
    j = nm2j(n_values, m_values)
    assert len(j) == 100 and j[0] == 0
    assert j[1] == ((n_values[1] + 2) * n_values[1] + m_values[1]) // 2
    assert isinstance(nm2j(2, 0), int)
    assert isinstance(nm2j([2], [0]), np.ndarray)


def test_zmoments_select_filters_by_absolute_m_values():
    data = np.arange(12, dtype=float).reshape(2, 6)
    n = np.array([0, 1, 1, 2, 2, 3])
    m = np.array([0, -1, 1, -2, 2, 3])
    z = zmoments(data=data, n=n, m=m)
    selected = z.select([1, -2])
    np.testing.assert_array_equal(selected.m, np.array([-1, 1, -2, 2]))
    np.testing.assert_array_equal(selected.n, np.array([1, 1, 2, 2]))
    np.testing.assert_array_equal(selected.data, data[:, [1, 2, 3, 4]])


# ---- index helpers against the golden vectors of the live reference --------------------------
def test_index_helpers_golden(golden):
    g = golden("index.npz")
    np.testing.assert_array_equal(nm2j(g["n"], g["m"]), g["j"])
    np.testing.assert_array_equal(nm2j_complex(g["n"], np.abs(g["m"])), g["jc"])
    np.testing.assert_array_equal(construct_complex_matrix(g["n12"], g["m12"]), g["cmat12"])
    np.testing.assert_array_equal(construct_rot_maps_matrix([1, 2, 3], [2, 3, 4, 6]), g["rotmat_a"])
    np.testing.assert_array_equal(construct_rot_maps_matrix([2, 3, 4, 6], g["m12"]), g["rotmat_12"])
    np.testing.assert_array_equal(construct_rot_maps_matrix([1, 2, 3], [2, 3, 4, 6]),
                                  [[1, 1, 1, 1], [1, -1, 1, 1], [-0.5, 1, -0.5, 1]])


def test_complex_real_matrices_roundtrip():
    n, m = zo.mode_table(9)
    cm = construct_complex_matrix(n, m)
    np.testing.assert_array_equal(cm, zo.complex_matrix(n, m))
    pick = np.where(cm == 1j, 0, cm)
    n_c = pick.dot(np.abs(n)).real.astype(int)
    m_c = pick.dot(np.abs(m)).real.astype(int)
    inv, n_r, m_r = construct_real_matrix(n_c, m_c)
    np.testing.assert_array_equal(n_r, n)
    np.testing.assert_array_equal(m_r, m)
    z = np.random.default_rng(0).normal(size=(5, len(n)))
    np.testing.assert_allclose((inv @ (cm @ z.T)).real.T, z, atol=1e-15)
    with pytest.raises(ValueError, match="Azimuthal frequency m must be non-negative."):
        nm2j_complex(2, -2)


def test_mode_table_and_zps_attributes():
    z = ZPs(12, 64)
    n, m = zo.mode_table(12)
    np.testing.assert_array_equal(z.n, n)
    np.testing.assert_array_equal(z.m, m)
    np.testing.assert_array_equal(nm2j(z.n, z.m), np.arange(91))


def test_zps_validation_messages():
    with pytest.raises(ValueError, match="n_max must be non-negative."):
        ZPs(-1, 8)
    with pytest.raises(ValueError, match="size must be positive."):
        ZPs(2, 0)
    with pytest.raises(ValueError, match=r"n_max=9 exceeds size=8\. This will produce meaningless results\. "
                                         r"Use n_max <= 4 for accurate moments\."):
        ZPs(9, 8)
    with pytest.warns(UserWarning, match=r"n_max=6 exceeds recommended limit of size/2≈4\."):
        ZPs(6, 8)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        ZPs(4, 8)
    z = ZPs(4, 8)
    with pytest.raises(ValueError, match="Images must be 2D or 3D array."):
        z.transform(np.zeros(5))
    with pytest.raises(ValueError, match=r"For batch processing, image size \(7x8\) must match polynomial size \(8x8\)"):
        z.transform(np.zeros((3, 7, 8)))
    with pytest.raises(ValueError, match=r"For FFT convolution, image size \(6x20\) must be at least as large as "
                                         r"polynomial size \(8x8\)"):
        z.transform(np.zeros((6, 20)))


def test_zps_is_sklearn_estimator():
    from sklearn.base import clone
    z = ZPs(10, 32, precision="fp32")
    assert z.get_params() == {"n_max": 10, "size": 32, "precision": "fp32", "output": "auto", "value_max": None}
    c = clone(z)
    assert (c.n_max, c.size, c.precision) == (10, 32, "fp32")
    assert z.fit(None) is z


def test_zmoments_validation_messages():
    with pytest.raises(ValueError, match="`n` and `m` must have the same shape."):
        zmoments(np.zeros((2, 3)), [0, 1, 1], [0, -1])
    with pytest.raises(ValueError, match="Data shape mismatch: expected 3 moments but got 4"):
        zmoments(np.zeros((2, 4)), [0, 1, 1], [0, -1, 1])
    with pytest.raises(ValueError, match="Data shape mismatch: expected 3 moments but got 2"):
        zmoments(np.zeros((2, 5, 5)), [0, 1, 1], [0, -1, 1])
    with pytest.raises(ValueError, match="Data must be 2D or 3D array."):
        zmoments(np.zeros(3), [0, 1, 1], [0, -1, 1])
    z = zmoments(np.zeros((2, 3)), [0, 1, 1], [0, -1, 1])
    with pytest.raises(ValueError, match="m=0 must be included in m_unselect."):
        z.rot_maps([2], m_unselect=(1,))


def test_zmoments_ctor_sorts_modes():
    data = np.arange(8, dtype=float).reshape(2, 4)
    z = zmoments(data, n=[2, 1, 2, 0], m=[2, 1, -2, 0])
    np.testing.assert_array_equal(z.n, [0, 1, 2, 2])
    np.testing.assert_array_equal(z.m, [0, 1, -2, 2])
    np.testing.assert_array_equal(z.data, data[:, [3, 1, 2, 0]])
    assert z.valid_mask is None and not z.is_complex
    u = z.unselect([0, 1])
    np.testing.assert_array_equal(u.m, [-2, 2])


def test_valid_mask_golden(golden):
    g = golden("lattice.npz")
    n, m = zo.mode_table(1)
    z = zmoments(np.zeros((3,) + g["map_valid"].shape), n, m, patch_size=48)
    np.testing.assert_array_equal(z.valid_mask, g["map_valid"])
    z = zmoments(np.zeros((3,) + g["map2_valid"].shape), n, m, patch_size=33)
    np.testing.assert_array_equal(z.valid_mask, g["map2_valid"])


def test_clear_border_golden(golden):
    g = golden("lattice.npz")
    for k in (32, 33):
        np.testing.assert_array_equal(clear_border(g["pts"], g["img"].shape, k), g[f"kept_{k}"])


# ---- the C-ABI surface -------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "zernike_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(zb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in zernike_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared            # the ctypes binding covers all of them
    assert _lib.load().zb200_abi_version() == 1
    assert _lib.load().zb200_num_modes(12) == 91 and _lib.load().zb200_num_complex_modes(12) == 49
    assert _lib.load().zb200_num_complex_modes(20) == 121


@pytest.mark.skipif(not _no_cuda(), reason="checks the no-device behaviour")
def test_no_silent_cpu_fallback():
    z = ZPs(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        z.transform(np.zeros((2, 8, 8), dtype=np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        z.transform(np.zeros((16, 16), dtype=np.float32))
    zm = zmoments(np.ones((2, 3)), [0, 1, 1], [0, -1, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        zm.to_complex()
    handle = ctypes.c_void_p()
    rc = _lib.load().zb200_plan_create(4, 8, ctypes.byref(handle))
    assert rc == _lib.ENODEV and "no CPU fallback" in _lib.last_error()
    # the rows either side of the path (peak detection, PCA) refuse to run without the device as well
    from motif_learn_b200.features import local_max, pca
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        local_max(np.zeros((16, 16), dtype=np.float32), 2.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pca(np.zeros((30, 4), dtype=np.float32), 2)


# ---- basis arithmetic (the code the CUDA generator runs) via the host harness -------------------
@pytest.fixture(scope="module")
def basis_tool(tmp_path_factory):
    out = tmp_path_factory.mktemp("tool") / "basis_host"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "motif-learn_b200", "csrc"),
                    "-o", str(out), os.path.join(ROOT, "tests", "host_tools", "basis_host.cpp")], check=True)
    return str(out)


@pytest.mark.parametrize("n_max,size,tol", [(4, 8, 1e-14), (6, 9, 1e-14), (5, 11, 1e-14), (10, 32, 1e-12),
                                            (12, 48, 1e-11), (12, 64, 1e-11), (20, 64, 1e-8), (12, 33, 1e-11)])
def test_basis_recurrence_matches_reference_algorithm(basis_tool, tmp_path, n_max, size, tol):
    path = tmp_path / "b.bin"
    subprocess.run([basis_tool, str(n_max), str(size), str(path)], check=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, ref = zo.zernike_basis(n_max, size)
    got = np.fromfile(path).reshape(ref.shape)
    np.testing.assert_array_equal(got[0] != 0, ref[0] != 0)      # identical disk mask (incl. ties)
    assert np.abs(got - ref).max() < tol


def test_basis_recurrence_is_closer_to_exact_than_reference(basis_tool, tmp_path):
    path = tmp_path / "b.bin"
    subprocess.run([basis_tool, "20", "64", str(path)], check=True)
    pts = [(r, c) for r in range(1, 64, 8) for c in range(3, 64, 10)]
    _, _, ex = zo.zernike_basis_exact(20, 64, pts)
    _, _, ref = zo.zernike_basis(20, 64)
    got = np.fromfile(path).reshape(ref.shape)
    e_got = max(np.abs(got[:, r, c] - ex[:, i]).max() for i, (r, c) in enumerate(pts))
    e_ref = max(np.abs(ref[:, r, c] - ex[:, i]).max() for i, (r, c) in enumerate(pts))
    assert e_got < 1e-12 < e_ref


def test_header_is_plain_c99_and_links(tmp_path):
    """include/zernike_b200.h is a C header (no C++/torch types cross the boundary): a C99 translation unit that
    calls the meta entry points compiles with -pedantic, links against the shared library and runs."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "zernike_b200.h"\n'
                   'int main(void) { int32_t n[3], m[3];\n'
                   '  if (zb200_abi_version() != ZB200_ABI_VERSION) return 1;\n'
                   '  if (zb200_num_modes(12) != 91 || zb200_num_complex_modes(12) != 49) return 2;\n'
                   '  if (zb200_mode_table(1, n, m) != 3 || m[1] != -1 || m[2] != 1) return 3;\n'
                   '  printf("ok\\n"); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-lzernike_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True)
    assert out.stdout.strip() == "ok"
