"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
of the live reference.  Runs on the B200 box: ``pytest -m gpu``.

Tolerances (DESIGN.md "Numerics"):
  * basis (fp64)            |ours - ref| <= 5e-12 (n_max<=12), 1e-8 (n_max=20: the REFERENCE's own
                            float-factorial error, ours is 2e-14 from exact)
  * fp32-grade contraction  allclose(rtol=1e-4, atol=1e-6*max|ref|)      (fp32 SIMT, tf32x3, f16x3)
  * 1xTF32 / 1xF16 map      |err| <= 1e-3*max|ref|                        (stated bound, fast modes)
  * n-fold scores           |err| <= 1e-5 (fp32-grade) on items with non-zero norm, NaN pattern equal
  * gather / index work     bit-exact
"""
import hashlib
import warnings

import numpy as np
import pytest

import zernike_oracle as zo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def api():
    from motif_learn_b200 import features
    return features


def fp32_close(got, ref, rtol=1e-4, scale=1e-6):
    ref = np.asarray(ref)
    np.testing.assert_allclose(np.asarray(got), ref, rtol=rtol, atol=scale * np.abs(ref).max())


def precisions(api, n_max, size):
    from motif_learn_b200 import _lib
    z = api.ZPs(n_max, size)
    out = ["fp32"]
    if _lib.load().zb200_plan_supports(z._plan, _lib.PREC_TF32X3, _lib.OUT_REAL):
        out += ["tf32x3", "tf32"]
    return out


def map_precisions(api, n_max, size):
    from motif_learn_b200 import _lib
    z = api.ZPs(n_max, size)
    out = ["fp32"]
    if _lib.load().zb200_plan_supports_map(z._plan, _lib.PREC_TF32X3):
        out += ["tf32x3", "tf32"]
    if _lib.load().zb200_plan_supports_map(z._plan, _lib.PREC_F16X3):
        out += ["f16x3", "f16"]
    return out


# ---- K1 basis ------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_max,size,tol", [(4, 8, 1e-13), (6, 9, 1e-13), (5, 11, 1e-13), (10, 32, 2e-12),
                                            (12, 48, 5e-12), (12, 64, 5e-12), (20, 64, 1e-8), (12, 33, 5e-12),
                                            (0, 1, 0), (3, 5, 1e-13)])
def test_basis_vs_oracle(api, n_max, size, tol):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z = api.ZPs(n_max, size)
    n, m, ref = zo.zernike_basis(n_max, size)
    got = z.polynomials
    assert got.shape == ref.shape and got.dtype == np.float64
    np.testing.assert_array_equal(z.n, n)
    np.testing.assert_array_equal(z.m, m)
    np.testing.assert_array_equal(got[0] != 0, ref[0] != 0)        # plane 0 is the disk mask (incl. ties)
    assert np.all(got[:, ref[0] == 0] == 0)                         # exactly zero outside the disk
    assert np.abs(got - ref).max() <= tol


def test_basis_golden_and_exact(api, golden):
    g = golden("basis.npz")
    np.testing.assert_allclose(api.ZPs(10, 32).polynomials, g["full_10_32"], rtol=0, atol=2e-12)
    for n_max, size, tol in [(12, 48, 5e-12), (12, 64, 5e-12), (20, 64, 1e-8)]:
        flat = api.ZPs(n_max, size).polynomials.ravel()
        np.testing.assert_allclose(flat[g[f"idx_{n_max}_{size}"]], g[f"val_{n_max}_{size}"], rtol=0, atol=tol)
    # SURVEY 8c known answers of the reference
    p = api.ZPs(10, 32).polynomials
    assert abs(p.sum() - 430.9346046698769) < 1e-8 and abs(p[4, 16, 16] + 1.7248414389629714) < 1e-12
    assert np.count_nonzero(p[0]) == 740 and np.all(p[0][p[0] != 0] == 1.0)
    # closer to the exact polynomial than the reference algorithm itself
    pts = [(r, c) for r in range(1, 64, 8) for c in range(3, 64, 10)]
    _, _, ex = zo.zernike_basis_exact(20, 64, pts)
    got = api.ZPs(20, 64).polynomials
    assert max(np.abs(got[:, r, c] - ex[:, i]).max() for i, (r, c) in enumerate(pts)) < 1e-12


# ---- K2 gather -----------------------------------------------------------------------------------
def test_gather_golden(api, golden, torch):
    g = golden("lattice.npz")
    img, pts = g["img"], g["pts"]
    for k in (32, 33):
        kp = api.KeyPoints(pts, img, k)
        np.testing.assert_array_equal(kp.pts, g[f"kept_{k}"])
        patches = kp.extract_patches()
        assert patches.dtype == np.float32 and patches.shape == (len(kp.pts), k, k)
        assert hashlib.sha256(np.ascontiguousarray(patches).tobytes()).hexdigest()[:16] == str(g[f"patch_sha_{k}"])
        np.testing.assert_array_equal(patches, zo.extract_patches(img, kp.pts, k))
    flat = api.KeyPoints(pts, img, 32).extract_patches(flat=True)
    np.testing.assert_array_equal(flat.shape, g["flat_shape_32"])
    # device-resident image keeps the result in HBM; half-to-even rounding of .5 coordinates
    dimg = torch.from_numpy(img).cuda()
    odd = np.array([[40.5, 41.5], [100.5, 64.5], [77.49999, 90.50001]])
    kp = api.KeyPoints(odd, dimg, 16)
    out = kp.extract_patches()
    assert out.is_cuda
    np.testing.assert_array_equal(out.cpu().numpy(), zo.extract_patches(img, odd, 16))
    empty = api.KeyPoints(np.zeros((0, 2)), img, 16).extract_patches()
    assert empty.shape == (0, 16, 16)


# ---- K3 projection ---------------------------------------------------------------------------------
def test_projection_golden(api, golden):
    g = golden("patches_nfold.npz")
    p = g["patches"]
    for n_max, key in [(12, "z12_data"), (20, "z20_data")]:
        for prec in precisions(api, n_max, 64):
            got = api.ZPs(n_max, 64, precision=prec).transform(p)
            assert got.data.dtype == np.float64 and got.data.shape == g[key].shape
            if prec == "tf32":
                assert np.abs(got.data - g[key]).max() <= 1e-3 * np.abs(g[key]).max()
            else:
                fp32_close(got.data, g[key])
    got = api.ZPs(12, 64).fit_transform(p)
    np.testing.assert_allclose(got.data[0, :4], [2.985956997797e-01, 6.946528841105e-03, 6.458243044575e-03,
                                                 2.461357425825e-03], rtol=1e-4)


def test_projection_config1_vs_oracle(api, torch):
    """BASELINE config 1: ZPs(n_max=10) on 10,000 32x32 patches of a hexagonal lattice."""
    from motif_learn_b200.datasets import honeycomb_image
    img, pts = honeycomb_image(1536, bond=12.0, seed=0)
    patches = api.KeyPoints(pts, img, 32).extract_patches()[:10000]
    assert patches.shape == (10000, 32, 32)
    _, _, v = zo.zernike_basis(10, 32)
    ref = zo.project_patches(patches.astype(np.float64), v)
    for prec in precisions(api, 10, 32):
        z = api.ZPs(10, 32, precision=prec)
        host = z.transform(patches).data                                   # host pipeline (chunked)
        dev = z.transform(torch.from_numpy(patches).cuda()).data            # device-resident
        assert dev.is_cuda and dev.dtype == torch.float32
        for got in (host, dev.cpu().numpy()):
            if prec == "tf32":
                assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
            else:
                fp32_close(got, ref)


@pytest.mark.parametrize("n_max,size,count", [(12, 64, 1000), (20, 64, 300), (8, 33, 257), (3, 6, 5), (12, 48, 129),
                                              (0, 4, 3), (21, 64, 130), (23, 64, 100)])
def test_projection_shapes_vs_oracle(api, n_max, size, count):
    rng = np.random.default_rng(n_max * 100 + size)
    patches = rng.random((count, size, size), dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, v = zo.zernike_basis(n_max, size)
        ref = zo.project_patches(patches.astype(np.float64), v)
        for prec in precisions(api, n_max, size):
            got = api.ZPs(n_max, size, precision=prec).transform(patches).data
            if prec == "tf32":
                assert np.abs(got - ref).max() <= 1e-3 * np.abs(ref).max()
            else:
                fp32_close(got, ref)                      # the stated gate: rtol 1e-4, atol 1e-6 * max|ref|


def test_projection_edge_cases(api, torch):
    z = api.ZPs(4, 8)
    out = z.transform(np.zeros((0, 8, 8), dtype=np.float32))
    assert out.data.shape == (0, 15)
    one = z.transform(np.ones((1, 8, 8), dtype=np.float64))        # float64 input is accepted
    _, _, v = zo.zernike_basis(4, 8)
    fp32_close(one.data, zo.project_patches(np.ones((1, 8, 8)), v))
    # linearity on the device path (size-independent property)
    rng = np.random.default_rng(5)
    a = torch.from_numpy(rng.random((300, 8, 8), dtype=np.float32)).cuda()
    b = torch.from_numpy(rng.random((300, 8, 8), dtype=np.float32)).cuda()
    za, zb, zab = (z.transform(t).data for t in (a, b, 2 * a - 3 * b))
    assert (zab - (2 * za - 3 * zb)).abs().max().item() < 1e-5


def test_fused_epilogues(api, golden, torch):
    g = golden("patches_nfold.npz")
    p = torch.from_numpy(g["patches"]).cuda()
    for n_max, cabs in [(12, np.abs(g["z12_cdata"])), (20, g["z20_cabs"])]:
        for prec in precisions(api, n_max, 64):
            z = api.ZPs(n_max, 64, precision=prec)
            tol = 1e-3 if prec == "tf32" else 2e-6
            got = z.transform_features(p, "abs").cpu().numpy()
            assert np.abs(got - cabs).max() <= tol * cabs.max()
            zc = z.transform_features(p, "complex").cpu().numpy()
            if n_max == 12:
                assert np.abs(zc - g["z12_cdata"]).max() <= tol * cabs.max()
            mag, ang = z.transform_features(p, "abs_phase")
            rebuilt = (mag * torch.exp(1j * ang)).cpu().numpy()
            assert np.abs(rebuilt - zc).max() <= 10 * tol * cabs.max()


# ---- zmoments algebra --------------------------------------------------------------------------------
def test_algebra_float64_matches_reference_golden(api, golden):
    """numpy float64 in -> float64 kernels -> numpy out: matches the live reference tightly."""
    g = golden("patches_nfold.npz")
    n, m = zo.mode_table(12)
    z = api.zmoments(g["z12_data"], n, m, patch_size=64)
    zc = z.to_complex()
    assert zc.is_complex and zc.data.dtype == np.complex128
    np.testing.assert_array_equal(zc.n, g["z12_cn"])
    np.testing.assert_array_equal(zc.m, g["z12_cm"])
    np.testing.assert_allclose(zc.data, g["z12_cdata"], rtol=1e-14, atol=0)
    assert zc.to_complex() is zc and z.to_real() is z
    back = zc.to_real()
    np.testing.assert_array_equal(back.m, m)
    np.testing.assert_allclose(back.data, g["z12_real_back"], rtol=1e-14, atol=0)
    np.testing.assert_allclose(z.normalize().data, g["z12_norm2"], rtol=1e-13)
    np.testing.assert_allclose(z.normalize(order=1).data, g["z12_norm1"], rtol=1e-13)
    np.testing.assert_allclose(z.normalize(order=np.inf).data, g["z12_norminf"], rtol=1e-13)
    np.testing.assert_allclose(z.normalize(order=3).data, zo.normalize(g["z12_data"], 3), rtol=1e-12)
    np.testing.assert_allclose(z.rotate(30.0).data, g["z12_rot30"], rtol=1e-12, atol=1e-16)
    np.testing.assert_allclose(z.rot_maps([2, 3, 4, 6]), g["z12_rot"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(z.rot_maps([3, 6], p=1), g["z12_rot_p1"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(z.rot_maps([2, 4], m_unselect=(0, 1, 2)), g["z12_rot_unsel"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(z.rot_maps([2, 3], p=None), zo.rot_maps(g["z12_data"], n, m, [2, 3], p=None),
                               rtol=1e-11, atol=1e-16)
    np.testing.assert_allclose(z.rot_maps([2, 3], p=np.inf), zo.rot_maps(g["z12_data"], n, m, [2, 3], p=np.inf),
                               rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(z.mirror_map(), g["z12_mirror"], rtol=1e-11)     # float64 data: float64 cos/sin table
    sel = z.select([2, -3])
    np.testing.assert_array_equal(sel.m, g["z12_sel_m"])
    np.testing.assert_array_equal(sel.data, g["z12_sel_data"])
    uns = z.unselect([0, 1])
    np.testing.assert_array_equal(uns.m, g["z12_unsel_m"])
    np.testing.assert_array_equal(uns.n, g["z12_unsel_n"])
    # complex normalisation and zero-norm -> NaN like numpy
    np.testing.assert_allclose(zc.normalize().data, zo.normalize(g["z12_cdata"]), rtol=1e-13)
    zero = api.zmoments(np.zeros((2, len(n))), n, m)
    assert np.isnan(zero.rot_maps([3])).all() and np.isnan(zero.normalize().data).all()


def test_algebra_device_float32_and_planar_layout(api, golden, torch):
    g = golden("lattice.npz")
    n, m = zo.mode_table(10)
    ref = g["z10_data"]
    z = api.zmoments(torch.from_numpy(ref.astype(np.float32)).cuda(), n, m, patch_size=32)
    zc = z.to_complex()
    assert zc.data.is_cuda and zc.data.dtype == torch.complex64
    assert np.abs(zc.data.cpu().numpy() - g["z10_cdata"]).max() < 1e-6
    rot = z.rot_maps([2, 3, 4, 6])
    assert rot.is_cuda and np.abs(rot.cpu().numpy() - g["z10_rot"]).max() < 1e-5
    # same data laid out as a (M, H, W) map: 96 patches -> an 8 x 12 "image"
    planar = np.ascontiguousarray(ref.T.reshape(len(n), 8, 12))
    zp = api.zmoments(planar, n, m, patch_size=32)
    np.testing.assert_allclose(zp.rot_maps([2, 3, 4, 6]), g["z10_rot"].T.reshape(4, 8, 12), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(zp.to_complex().data, g["z10_cdata"].T.reshape(-1, 8, 12), rtol=1e-14)
    np.testing.assert_allclose(zp.mirror_map(), zo.mirror_map(planar, n, m), rtol=1e-11)
    np.testing.assert_allclose(zp.to_complex().to_real().data, planar, rtol=1e-14)
    assert zp.valid_mask.shape == (8, 12)
    # ctor permutation on device data
    perm = np.random.default_rng(0).permutation(len(n))
    zs = api.zmoments(z.data[:, torch.from_numpy(perm).cuda()], n[perm], m[perm])
    np.testing.assert_array_equal(zs.m, m)
    assert torch.equal(zs.data, z.data)


# ---- K4 dense map --------------------------------------------------------------------------------------
def test_map_golden(api, golden, torch):
    g = golden("lattice.npz")
    img = g["map_img"]
    ys, xs = g["map_ys"], g["map_xs"]
    precs = map_precisions(api, 12, 48)
    assert "tf32x3" in precs and "f16x3" in precs   # the tensor-core maps must be available for 48-px windows
    for prec in precs:
        z = api.ZPs(12, 48, precision=prec)
        zm = z.transform(img)
        assert zm.data.shape == (91, 160, 192) and zm.data.dtype == np.float64
        np.testing.assert_array_equal(zm.valid_mask, g["map_valid"])
        fused = z.symmetry_map(img, [2, 3, 4, 6])
        assert fused.shape == (4, 160, 192)
        if prec in ("tf32", "f16"):               # fast modes: stated bounds
            assert np.abs(zm.data[:, ys, xs] - g["map_moments"]).max() <= 1e-3 * np.abs(g["map_moments"]).max()
            assert np.abs(fused - g["map_rot"]).max() < 2e-3
            continue
        fp32_close(zm.data[:, ys, xs], g["map_moments"])
        rot = zm.rot_maps([2, 3, 4, 6])
        assert np.abs(rot - g["map_rot"]).max() < 1e-5
        assert np.abs(np.abs(zm.to_complex().data)[:, ys, xs] - g["map_cabs_pts"]).max() < 1e-6
        assert np.abs(zm.mirror_map()[ys, xs] - g["map_mirror_pts"]).max() < 1e-5
        # fused map -> scores (never materialises the moments)
        assert np.abs(fused - g["map_rot"]).max() < 1e-5
        assert np.abs(z.symmetry_map(img, [3, 6], p=1) - zo.rot_maps(zm.data, z.n, z.m, [3, 6], p=1)).max() < 1e-5
        # row bands (image-tile sharding) reproduce the full result exactly
        dimg = torch.from_numpy(img.astype(np.float32)).cuda()
        full = z.symmetry_map(dimg, [2, 3, 4, 6])
        parts = [z.symmetry_map(dimg, [2, 3, 4, 6], row0=r0, rows=r) for r0, r in [(0, 50), (50, 37), (87, 73)]]
        assert torch.equal(torch.cat(parts, dim=1), full)
        band = z._transform_map(dimg, row0=100, rows=9).data
        assert torch.equal(band, z.transform(dimg).data[:, 100:109])


def test_map_odd_window_golden(api, golden):
    g = golden("lattice.npz")
    img = g["map2_img"]
    z = api.ZPs(8, 33)
    zm = z.transform(img)
    fp32_close(zm.data[:, ::7, ::5], g["map2_moments"])
    assert np.abs(zm.rot_maps([3, 6]) - g["map2_rot"]).max() < 1e-5
    assert np.abs(z.symmetry_map(img, [3, 6]) - g["map2_rot"]).max() < 1e-5
    np.testing.assert_array_equal(zm.valid_mask, g["map2_valid"])


def test_map_vs_oracle_fft_nonmultiple_width(api):
    """Width not a multiple of the 128-pixel tile nor of 4; zero-norm pixels give NaN scores."""
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image((150, 203), bond=12.0, seed=4, angle=11.0, vacancy_frac=0.02)
    img[:, 150:] = 0.0                                   # a dead region: windows there are all-zero
    n, m, v = zo.zernike_basis(6, 20)
    z = api.ZPs(6, 20)
    ref = zo.moment_map_fft(img.astype(np.float64), v, n)
    fp32_close(z.transform(img).data, ref)
    got = z.symmetry_map(img, [2, 3, 6])
    want = zo.rot_maps(zo.moment_map_direct(img.astype(np.float64), v), n, m, [2, 3, 6])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    assert np.isnan(want).any()
    ok = ~np.isnan(want) & (np.abs(ref[3:]).sum(axis=0) > 1e-3)[None]
    assert np.abs(got - want)[ok].max() < 1e-4


def test_map_equals_gather_plus_projection_at_config2_size(api, torch):
    """BASELINE config 2 geometry (2048x2048, n_max=12, 48-px window): the dense map at a pixel
    equals the projection of the patch gathered at that pixel (K4 == K2+K3), everywhere incl.
    tile seams, and the fused scores equal rot_maps of those projections."""
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image(2048, bond=12.0, seed=0)
    dimg = torch.from_numpy(img).cuda()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.integers(30, 2018, 400), [127, 128, 129, 1023, 1024, 2000]])
    ys = np.concatenate([rng.integers(30, 2018, 400), [127, 128, 129, 1023, 1024, 2000]])
    pts = np.stack([xs, ys], axis=1).astype(np.float64)
    patches = zo.extract_patches(img, pts, 48)
    n, m, v = zo.zernike_basis(12, 48)
    zref = zo.project_patches(patches.astype(np.float64), v)
    want = zo.rot_maps(zref, n, m, [2, 3, 4, 6])
    for prec in [pr for pr in map_precisions(api, 12, 48) if pr not in ("tf32", "f16")]:
        z = api.ZPs(12, 48, precision=prec)
        scores = z.symmetry_map(dimg, [2, 3, 4, 6])
        assert scores.shape == (4, 2048, 2048)
        got = scores[:, torch.from_numpy(ys).cuda(), torch.from_numpy(xs).cuda()].cpu().numpy().T
        assert np.abs(got - want).max() < 1e-5
        band = z._transform_map(dimg, row0=1000, rows=48).data
        sel = (ys >= 1000) & (ys < 1048)
        if sel.any():
            gotm = band[:, torch.from_numpy(ys[sel] - 1000).cuda(), torch.from_numpy(xs[sel]).cuda()].cpu().numpy().T
            fp32_close(gotm, zref[sel])


# ---- remaining BASELINE configurations -------------------------------------------------------------
def test_config3_complex_n20_on_peak_patches(api, torch):
    """BASELINE config 3 family: complex ZPs n_max=20 on 64x64 patches extracted at atom peaks.
    Parity on a 4096-patch subset against the oracle (the full 1M stack is a throughput case)."""
    from motif_learn_b200.datasets import honeycomb_image
    img, pts = honeycomb_image(1024, bond=12.0, seed=5, angle=17.0, jitter=0.3, noise=0.01)
    dimg = torch.from_numpy(img).cuda()
    kp = api.KeyPoints(pts, dimg, 64)
    patches = kp.extract_patches()
    assert patches.is_cuda and patches.shape[0] >= 4096
    patches = patches[:4096].contiguous()
    host = patches.cpu().numpy()
    np.testing.assert_array_equal(host, zo.extract_patches(img, kp.pts[:4096], 64))
    n, m, v = zo.zernike_basis(20, 64)
    ref = zo.project_patches(host.astype(np.float64), v)
    refc, n_c, m_c = zo.to_complex(ref, n, m)
    z = api.ZPs(20, 64)
    zm = z.transform(patches)
    fp32_close(zm.data.cpu().numpy(), ref)
    zc = zm.to_complex()
    np.testing.assert_array_equal(zc.m, m_c)
    assert np.abs(zc.data.cpu().numpy() - refc).max() <= 3e-6 * np.abs(refc).max()
    feats = z.transform_features(patches, "abs")          # falls back to real + packing kernels at n_max=20
    assert feats.shape == (4096, 121)
    assert np.abs(feats.cpu().numpy() - np.abs(refc)).max() <= 3e-6 * np.abs(refc).max()
    mag, ph = z.transform_features(patches, "abs_phase")
    big = np.abs(refc) > 1e-3 * np.abs(refc).max()
    dphi = np.angle(np.exp(1j * (ph.cpu().numpy() - np.angle(refc))))
    assert np.abs(dphi[big]).max() < 1e-3


def test_config4_tiled_map_with_defects(api, torch):
    """BASELINE config 4 geometry (64-px window, frame with vacancies/dopants) on a 1024x1536 frame:
    row bands computed independently (as the GPUs of one box would) equal the oracle on three bands
    including the seams, and equal the single-call result bit-exactly."""
    from motif_learn_b200.datasets import honeycomb_image
    from motif_learn_b200.parallel import row_band
    H, W = 1024, 1536
    img, _ = honeycomb_image((H, W), bond=12.0, seed=11, vacancy_frac=0.01, dopant_frac=0.005)
    dimg = torch.from_numpy(img).cuda()
    z = api.ZPs(12, 64)
    full = z.symmetry_map(dimg, [2, 3, 4, 6])
    bands = [z.symmetry_map(dimg, [2, 3, 4, 6], row0=r0, rows=r) for r0, r in (row_band(H, k, 4) for k in range(4))]
    assert torch.equal(torch.cat(bands, dim=1), full)
    n, m, v = zo.zernike_basis(12, 64)
    got = full.cpu().numpy()
    for r0 in (0, 224, 960):                               # top border, a band seam (256 +- 32), bottom border
        lo, hi = max(0, r0 - 32), min(H, r0 + 64 + 32)
        sub = img[lo:hi].astype(np.float64)
        ref = zo.rot_maps(zo.moment_map_fft(sub, v, n, chunk=16), n, m, [2, 3, 4, 6])
        # rows whose 64-px window lies inside the cropped strip (or touches the true frame border)
        a = 0 if lo == 0 else 32
        b = (hi - lo) if hi == H else (hi - lo) - 32
        assert np.nanmax(np.abs(got[:, lo + a:lo + b] - ref[:, a:b])) < 1e-5


def test_config5_frame_batch_gather_then_features(api, torch):
    """BASELINE config 5 family on 3 frames: peaks -> clear_border -> gather -> ZPs(12,64) -> |Zc|,
    frame-sharded exactly like the 8-GPU run would (each frame independent)."""
    from motif_learn_b200.datasets import honeycomb_image
    z = api.ZPs(12, 64)
    n, m, v = zo.zernike_basis(12, 64)
    for seed in (0, 1, 2):
        img, pts = honeycomb_image(512, bond=12.0, seed=seed, angle=10.0 * seed, jitter=0.3)
        kp = api.KeyPoints(pts, torch.from_numpy(img).cuda(), 64)
        np.testing.assert_array_equal(kp.pts, zo.clear_border(pts, img.shape, 64))
        patches = kp.extract_patches()
        ref = zo.project_patches(zo.extract_patches(img, kp.pts, 64).astype(np.float64), v)
        fp32_close(z.transform(patches).data.cpu().numpy(), ref)
        refabs = np.abs(zo.to_complex(ref, n, m)[0])
        got = z.transform_features(patches, "abs").cpu().numpy()
        assert np.abs(got - refabs).max() <= 3e-6 * refabs.max()


def test_fused_gather_projection(api, torch):
    """K2 fused into K3 (zb200_project_peaks_f32): same numbers as gather + projection, bit-for-bit
    against the unfused tensor-core path, and fp32-grade against the oracle; border windows read zeros."""
    from motif_learn_b200.datasets import honeycomb_image
    img, pts = honeycomb_image((600, 777), bond=12.0, seed=9, angle=23.0, jitter=0.4, noise=0.01)
    dimg = torch.from_numpy(img).cuda()
    for n_max, k in ((12, 64), (10, 32), (12, 48)):
        kept = zo.clear_border(pts, img.shape, k)
        z = api.ZPs(n_max, k)
        n, m, v = zo.zernike_basis(n_max, k)
        patches = zo.extract_patches(img, kept, k)
        ref = zo.project_patches(patches.astype(np.float64), v)
        fused = z.transform_peaks(dimg, kept, fused=True)
        assert fused.data.is_cuda and fused.data.shape == ref.shape
        fp32_close(fused.data.cpu().numpy(), ref)
        # the tf32x3 fused kernel equals the unfused kernel of the same arithmetic bit for bit
        zt = api.ZPs(n_max, k, precision="tf32x3")
        unfused = zt.transform(torch.from_numpy(patches).cuda()).data
        assert torch.equal(zt.transform_peaks(dimg, kept, fused=True).data, unfused)
        refabs = np.abs(zo.to_complex(ref, n, m)[0])
        assert np.abs(z.transform_peaks(dimg, kept, "abs", fused=True).cpu().numpy() - refabs).max() <= 3e-6 * refabs.max()
        zc = z.transform_peaks(dimg, kept, "complex", fused=True).cpu().numpy()
        refc = zo.to_complex(ref, n, m)[0]
        fp32_close(np.concatenate([zc.real, zc.imag], axis=1), np.concatenate([refc.real, refc.imag], axis=1))
        # default route: 64-pixel windows gather inside the mirror-folded kernel (frame maximum as the range bound),
        # the other shapes run gather kernel + projection; the two-kernel route stays available
        default = z.transform_peaks(dimg, kept).data
        fp32_close(default.cpu().numpy(), ref)
        two = z.transform_peaks(dimg, kept, fused=False).data
        assert torch.equal(two, z.transform(torch.from_numpy(patches).cuda()).data)
        fp32_close(two.cpu().numpy(), ref)
        if k == 64:
            assert torch.equal(default, fused.data)
        host = z.transform_peaks(img, kept, fused=True)           # numpy frame in -> float64 numpy out
        assert isinstance(host.data, np.ndarray) and host.data.dtype == np.float64
        fp32_close(host.data, ref)
    # windows hanging over the frame edge: zero extension, like the dense map
    z = api.ZPs(12, 64)
    edge = np.array([[3.0, 5.0], [770.0, 590.0], [400.2, 10.7]])
    got = z.transform_peaks(dimg, edge, fused=True).data.cpu().numpy()
    dense = z.transform(dimg).data
    c = np.rint(edge).astype(int)
    want = dense[:, torch.from_numpy(c[:, 1]).cuda(), torch.from_numpy(c[:, 0]).cuda()].cpu().numpy().T
    # two fp32-grade KERNELS against each other (tf32x3 projection vs f16x3 dense map), each within the gate of
    # the oracle: the difference may reach twice the atol term
    fp32_close(got, want, rtol=1e-4, scale=2e-6)
    assert z.transform_peaks(dimg, np.zeros((0, 2))).data.shape == (0, 91)


def test_fused_patch_scores(api, golden, torch):
    """rot_maps fused into the projection epilogue (zb200_project_patches_scores_f32)."""
    g = golden("patches_nfold.npz")
    p = g["patches"]
    for prec in precisions(api, 12, 64):
        z = api.ZPs(12, 64, precision=prec)
        tol = 2e-3 if prec == "tf32" else 1e-5
        got = z.symmetry_scores(p, [2, 3, 4, 6])
        assert got.shape == g["z12_rot"].shape and np.abs(got - g["z12_rot"]).max() < tol
        assert np.abs(z.symmetry_scores(p, [3, 6], p=1) - g["z12_rot_p1"]).max() < tol
        assert np.abs(z.symmetry_scores(p, [2, 4], m_unselect=(0, 1, 2)) - g["z12_rot_unsel"]).max() < tol
        dev = z.symmetry_scores(torch.from_numpy(p).cuda(), [2, 3, 4, 6])
        assert dev.is_cuda and np.abs(dev.cpu().numpy() - g["z12_rot"]).max() < tol
    big = np.random.default_rng(3).random((3000, 64, 64), dtype=np.float32)
    _, _, v = zo.zernike_basis(12, 64)
    n, m = zo.mode_table(12)
    want = zo.rot_maps(zo.project_patches(big.astype(np.float64), v), n, m, [2, 3, 4, 5, 6, 8])
    assert np.abs(api.ZPs(12, 64).symmetry_scores(big, [2, 3, 4, 5, 6, 8]) - want).max() < 1e-5
    z20 = api.ZPs(20, 64)                                   # 231 modes: unfused fallback on the fp32-grade path
    want20 = zo.rot_maps(g["z20_data"], *zo.mode_table(20), [3, 6])
    assert np.abs(z20.symmetry_scores(p, [3, 6]) - want20).max() < 1e-5


def test_render_atoms_golden(golden, torch):
    """"next" row f1: the GPU frame renderer against frames drawn by the live reference."""
    from motif_learn_b200.datasets import render_atoms_gpu
    g = golden("render.npz")
    sigma = float(g["sigma"])
    n_a = int((g["amps"] == 1.0).sum())
    img = render_atoms_gpu((160, 160), g["pts"][:n_a], 1.0, sigma)                 # A sub-lattice ...
    img = render_atoms_gpu((160, 160), g["pts"][n_a:], 0.5, sigma, out=img)        # ... then B, like to_image
    assert img.is_cuda and img.dtype == torch.float32
    np.testing.assert_allclose(img.cpu().numpy(), g["img"], rtol=0, atol=3e-7)
    img2 = render_atoms_gpu((120, 160), g["pts"], g["amps2"], sigma)
    np.testing.assert_allclose(img2.cpu().numpy(), g["img2"], rtol=0, atol=3e-7)
    np.testing.assert_allclose(img2.cpu().numpy(), zo.render_atoms((120, 160), g["pts"], g["amps2"], sigma), rtol=0,
                               atol=3e-7)
    many = np.random.default_rng(0).random((5000, 2)) * [700, 300]                # list overflow path (dense atoms)
    dense = render_atoms_gpu((300, 700), many, 0.3, 2.0)
    np.testing.assert_allclose(dense.cpu().numpy(), zo.render_atoms((300, 700), many, 0.3, 2.0), rtol=0, atol=2e-6)
    assert float(render_atoms_gpu((16, 16), np.zeros((0, 2)), 1.0, 3.0).abs().max()) == 0.0


# ---- "next" row f2: peak detection ------------------------------------------------------------------------
def test_local_max_golden(api, golden, torch):
    """CUDA local_max against the reference-generated goldens: identical peaks, identical order."""
    g = golden("peaks.npz")
    for tag in ("a", "b"):
        img = g[f"img_{tag}"]
        i = 0
        while f"pts_{tag}{i}" in g.files:
            r, thr = g[f"arg_{tag}{i}"]
            thr = None if np.isnan(thr) else float(thr)
            got = api.local_max(img, float(r), thr)
            assert got.dtype == np.int64
            np.testing.assert_array_equal(got, g[f"pts_{tag}{i}"])
            dev = api.local_max(torch.from_numpy(img).cuda(), float(r), thr, as_tensor=True)
            assert dev.is_cuda and torch.equal(dev.cpu().long(), torch.from_numpy(g[f"pts_{tag}{i}"]))
            i += 1


def test_local_max_vs_oracle_frame_and_edge_cases(api):
    """A 1024^2 noisy lattice frame against the oracle; constant frame, plateaus, border maxima, ties."""
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image(1024, bond=12.0, seed=3, angle=5.0, jitter=0.2, noise=0.02)
    for r, thr in ((6.0, None), (6.0, 0.25), (2.5, 0.1)):
        np.testing.assert_array_equal(api.local_max(img, r, thr), zo.local_max(img, r, thr))
    assert api.local_max(np.full((9, 11), 3.0, dtype=np.float32), 2.0).shape == (0, 2)
    assert api.local_max(np.full((9, 11), 3.0, dtype=np.float32), 2.0, threshold=1.0).shape == (0, 2)
    small = np.zeros((7, 7), dtype=np.float32)
    small[0, 3] = 5.0
    small[3, 3] = small[3, 4] = 2.0
    np.testing.assert_array_equal(api.local_max(small, 1.0), [[3, 3]])
    np.testing.assert_array_equal(api.local_max(small, 0.5), [[3, 3], [4, 3]])
    # exact ties everywhere (noise-free lattice): our raster-order rule == the oracle's, and the result is a
    # valid suppression (kept peaks pairwise farther than r apart)
    clean, _ = honeycomb_image(256, bond=12.0, seed=0)
    got = api.local_max(clean, 5.0)
    np.testing.assert_array_equal(got, zo.local_max(clean, 5.0))
    d = np.sqrt(((got[:, None, :] - got[None, :, :]) ** 2).sum(-1)) + np.eye(len(got)) * 1e9
    assert d.min() > 5.0


def test_local_max_feeds_keypoints_and_zps(api, torch):
    """The path end to end on the device: frame -> local_max -> KeyPoints -> ZPs, against the oracle."""
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image(512, bond=12.0, seed=8, jitter=0.2, noise=0.01)
    pts = api.local_max(img, 5.0, 0.3)
    ref_pts = zo.local_max(img, 5.0, 0.3)
    np.testing.assert_array_equal(pts, ref_pts)
    kp = api.KeyPoints(pts, torch.from_numpy(img).cuda(), 32)
    z = api.ZPs(10, 32).transform(kp.extract_patches())
    n, m, v = zo.zernike_basis(10, 32)
    kept = zo.clear_border(ref_pts, img.shape, 32)
    ref = zo.project_patches(zo.extract_patches(img, kept, 32).astype(np.float64), v)
    fp32_close(z.data.cpu().numpy(), ref)


def test_config4_full_size_map_property(api, torch):
    """BASELINE config 4 at FULL size (4096x4096 frame with defects, 64-px window, n_max=12): the fused
    symmetry map at sampled pixels (incl. span seams at multiples of 512, row-pair seams, frame borders)
    equals rot_maps of the oracle projection of the window gathered at that pixel; four row bands (the 4-GPU
    sharding, odd band starts included) reproduce the single-call result bit for bit."""
    from motif_learn_b200.datasets import honeycomb_image
    S = 4096
    img, _ = honeycomb_image(S, bond=12.0, seed=2, vacancy_frac=0.01, dopant_frac=0.005)
    dimg = torch.from_numpy(img).cuda()
    z = api.ZPs(12, 64)
    full = z.symmetry_map(dimg, [2, 3, 4, 6])
    assert full.shape == (4, S, S)
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.integers(34, S - 34, 300), [511, 512, 513, 2047, 2048, 4000, 40, 4055]])
    ys = np.concatenate([rng.integers(34, S - 34, 300), [1000, 1001, 2047, 2048, 2049, 33, 4060, 4061]])
    pts = np.stack([xs, ys], axis=1).astype(np.float64)
    n, m, v = zo.zernike_basis(12, 64)
    zref = zo.project_patches(zo.extract_patches(img, pts, 64).astype(np.float64), v)
    want = zo.rot_maps(zref, n, m, [2, 3, 4, 6])
    got = full[:, torch.from_numpy(ys).cuda(), torch.from_numpy(xs).cuda()].cpu().numpy().T
    assert np.abs(got - want).max() < 1e-5
    bands = [(0, 1023), (1023, 1026), (2049, 999), (3048, 1048)]
    parts = [z.symmetry_map(dimg, [2, 3, 4, 6], row0=r0, rows=r) for r0, r in bands]
    assert torch.equal(torch.cat(parts, dim=1), full)


def test_map_many_modes_two_passes(api):
    """n_max >= 15 gives more than 128 padded modes: the tensor-core map runs one pass per mode block and the
    scores come from the materialised moments (n_max=16 -> 153 modes, n_max=20 -> 231 modes)."""
    from motif_learn_b200 import _lib
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image((140, 600), bond=12.0, seed=9, angle=3.0)
    for n_max, k in ((16, 40), (20, 64)):
        z = api.ZPs(n_max, k)
        assert _lib.load().zb200_plan_supports_map(z._plan, _lib.PREC_F16X3)
        n, m, v = zo.zernike_basis(n_max, k)
        ref = zo.moment_map_fft(img.astype(np.float64), v, n, chunk=32)
        got = z.transform(img).data
        assert got.shape == ref.shape
        fp32_close(got, ref)
        want = zo.rot_maps(ref, n, m, [2, 3, 4, 6])
        assert np.nanmax(np.abs(z.symmetry_map(img, [2, 3, 4, 6]) - want)) < 1e-5


# ---- "next" row f4: PCA of the feature matrix ---------------------------------------------------------
def test_pca_golden_and_live_sklearn(api, golden, torch):
    """CUDA pca against the goldens of the live reference (scikit-learn PCA.fit_transform), and against
    scikit-learn itself on a tall matrix of device-computed features (float64 accumulation: 1e-10)."""
    g = golden("pca.npz")
    for c, key in ((2, "pca2"), (5, "pca5")):
        got = api.pca(g["feats"], c)
        assert got.shape == g[key].shape and got.dtype == np.float64
        np.testing.assert_allclose(got, g[key], rtol=0, atol=1e-10)
    from sklearn.decomposition import PCA
    rng = np.random.default_rng(0)
    base = rng.normal(size=(20000, 6)) @ rng.normal(size=(6, 91)) + 0.05 * rng.normal(size=(20000, 91))
    X = (base + rng.normal(size=91)).astype(np.float32)
    ref = PCA(n_components=3).fit_transform(X.astype(np.float64))
    np.testing.assert_allclose(api.pca(X, 3), ref, rtol=0, atol=1e-9 * np.abs(ref).max())
    dev = api.pca(torch.from_numpy(X).cuda(), 3)
    assert dev.is_cuda and dev.dtype == torch.float64
    np.testing.assert_allclose(dev.cpu().numpy(), ref, rtol=0, atol=1e-9 * np.abs(ref).max())
    # more features than one 96-column tile (n_max=20 real moments: 231 columns)
    Y = (rng.normal(size=(5000, 4)) @ rng.normal(size=(4, 231)) + 0.1 * rng.normal(size=(5000, 231))).astype(np.float32)
    ref = PCA(n_components=2).fit_transform(Y.astype(np.float64))
    np.testing.assert_allclose(api.pca(Y, 2), ref, rtol=0, atol=1e-9 * np.abs(ref).max())


def test_download_as_f64_multi_chunk(torch):
    """zb200_download_as_f64 (the float64 host result every numpy-in call returns): several 32 MiB staging
    chunks plus a ragged tail, threaded widening -- bit-identical to the float32 values."""
    from motif_learn_b200.features._zps import _host_f64
    t = torch.randn(20_000_003, device="cuda", dtype=torch.float32)
    got = _host_f64(t)
    assert got.dtype == np.float64 and got.shape == (20_000_003,)
    np.testing.assert_array_equal(got, t.cpu().numpy().astype(np.float64))
    small = torch.arange(7, device="cuda", dtype=torch.float32).reshape(7, 1)
    np.testing.assert_array_equal(_host_f64(small), np.arange(7, dtype=np.float64).reshape(7, 1))


@pytest.mark.parametrize("n_max,k", [(0, 1), (1, 2), (2, 3), (4, 5), (6, 16), (8, 17), (12, 128)])
def test_map_window_sizes_vs_oracle(api, n_max, k):
    """Every window size runs on the tensor-core map: 1-pixel and odd windows, one / several 16-tap groups,
    the 128-px maximum (two basis k-blocks per window row)."""
    rng = np.random.default_rng(k)
    shape = (k + 3, 140 + k) if k < 100 else (131, 300)
    img = rng.random(shape).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z = api.ZPs(n_max, k)
    n, m, v = zo.zernike_basis(n_max, k)
    ref = zo.moment_map_fft(img.astype(np.float64), v, n)
    fp32_close(z.transform(img).data, ref)


def test_integration_stubs_bind_the_c_abi_directly(torch):
    """The two ctypes stubs of INTEGRATION.md section 2 (what a maintainer of the reference would add), run as
    written against the shared library alone: patch route through zb200_project_patches_host, image route through
    zb200_moment_map_f32 + zb200_download_as_f64."""
    import ctypes as C
    from motif_learn_b200 import _lib as ours
    lib = C.CDLL(ours.LIB_PATH)
    lib.zb200_plan_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.zb200_project_patches_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    lib.zb200_moment_map_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.zb200_download_as_f64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.zb200_last_error.restype = C.c_char_p
    rng = np.random.default_rng(5)
    n, m, v = zo.zernike_basis(8, 24)
    plan = C.c_void_p()
    assert lib.zb200_plan_create(8, 24, C.byref(plan)) == 0, lib.zb200_last_error()
    x = rng.random((300, 24, 24)).astype(np.float32)
    out = np.empty((300, len(n)), dtype=np.float64)
    assert lib.zb200_project_patches_host(plan, x.ctypes.data, 300, 2, out.ctypes.data) == 0, lib.zb200_last_error()
    fp32_close(out, zo.project_patches(x.astype(np.float64), v))
    image = rng.random((70, 90)).astype(np.float32)
    img = torch.from_numpy(image).cuda()
    dev = torch.empty((len(n), 70, 90), dtype=torch.float32, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.zb200_moment_map_f32(plan, img.data_ptr(), 70, 90, 0, 70, 4, dev.data_ptr(), stream) == 0, lib.zb200_last_error()
    host = np.empty((len(n), 70, 90), dtype=np.float64)
    assert lib.zb200_download_as_f64(dev.data_ptr(), dev.numel(), host.ctypes.data, stream) == 0
    fp32_close(host, zo.moment_map_fft(image.astype(np.float64), v, n))
    # third stub of INTEGRATION.md: the notebooks' patch route end to end from a host frame
    lib.zb200_project_peaks_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_int, C.c_int, C.c_void_p]
    frame = np.ascontiguousarray(rng.random((120, 150)), dtype=np.float32)
    pts = zo.clear_border(rng.uniform(0, 120, (40, 2)), frame.shape, 24).astype(np.float64)
    counts = np.array([len(pts)], dtype=np.int64)
    frames = (C.c_void_p * 1)(frame.ctypes.data)
    feats = np.empty((len(pts), len(n)), dtype=np.float64)
    assert lib.zb200_project_peaks_host(plan, frames, 1, 120, 150, pts.ctypes.data, counts.ctypes.data, 2, 0, 1,
                                        feats.ctypes.data) == 0, lib.zb200_last_error()
    fp32_close(feats, zo.project_patches(zo.extract_patches(frame, pts, 24).astype(np.float64), v))


# ---- host-buffer routes (round 2) -----------------------------------------------------------------
def test_transform_peaks_batch_host_route(api, torch):
    """zb200_project_peaks_host: frames + peak lists in, float64 features out, frames pipelined over two streams."""
    from motif_learn_b200.datasets import honeycomb_image
    frames, pts = [], []
    for seed, angle in ((0, 0.0), (1, 17.0), (2, 41.0), (3, 5.0), (4, 29.0)):
        img, p = honeycomb_image((320, 416), bond=12.0, seed=seed, angle=angle, jitter=0.2, noise=0.01)
        frames.append(img)
        pts.append(zo.clear_border(p, img.shape, 48))
    pts[3] = pts[3][:0]                                             # a frame without peaks inside the series
    z = api.ZPs(12, 48)
    n, m, v = zo.zernike_basis(12, 48)
    for kind in ("real", "complex", "abs"):
        got = z.transform_peaks_batch(frames, pts, kind)
        assert len(got) == len(frames)
        for f, (img, p) in enumerate(zip(frames, pts)):
            ref = zo.project_patches(zo.extract_patches(img, p, 48).astype(np.float64), v) if len(p) else np.zeros((0, 91))
            zc = zo.to_complex(ref, n, m)[0]
            if kind == "real":
                assert got[f].data.dtype == np.float64 and got[f].data.shape == ref.shape
                if len(p):
                    fp32_close(got[f].data, ref)
            elif kind == "complex":
                assert got[f].dtype == np.complex128 and got[f].shape == zc.shape
                if len(p):
                    assert np.abs(got[f] - zc).max() <= 2e-6 * np.abs(zc).max()
            else:
                assert got[f].dtype == np.float64
                if len(p):
                    assert np.abs(got[f] - np.abs(zc)).max() <= 2e-6 * np.abs(zc).max()
    # the one-frame call of a numpy user goes through the same entry point and equals the device route
    one = z.transform_peaks(frames[1], pts[1])
    dev = z.transform_peaks(torch.from_numpy(frames[1]).cuda(), pts[1])
    assert isinstance(one.data, np.ndarray) and one.data.dtype == np.float64 and dev.data.is_cuda
    np.testing.assert_array_equal(one.data, dev.data.double().cpu().numpy())
    assert z.transform_peaks_batch([], []) == []
    with pytest.raises(ValueError):
        z.transform_peaks_batch(frames[:2], pts[:1])
    # return types no longer depend on which internal route ran: numpy frame -> numpy for every kind
    for fused in (False, True):
        assert isinstance(z.transform_peaks(frames[0], pts[0], "abs", fused=fused), np.ndarray)
        assert z.transform_peaks(torch.from_numpy(frames[0]).cuda(), pts[0], "abs", fused=fused).is_cuda


def test_host_pipeline_is_thread_safe(api):
    """Two host threads on ONE plan (ctypes drops the GIL): the per-plan staging buffers are guarded by a mutex,
    results must be those of the serial calls (ADVICE r1, zb200_api.cu:314)."""
    import threading
    rng = np.random.default_rng(3)
    a = rng.random((9000, 32, 32), dtype=np.float32)
    b = rng.random((7000, 32, 32), dtype=np.float32)
    z = api.ZPs(10, 32)
    want = {"a": z.transform(a).data, "b": z.transform(b).data}
    got, errs = {}, []

    def work(key, arr):
        try:
            for _ in range(3):
                got[key] = api.ZPs(10, 32).transform(arr).data
        except Exception as exc:        # pragma: no cover
            errs.append(exc)

    threads = [threading.Thread(target=work, args=("a", a)), threading.Thread(target=work, args=("b", b))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs
    np.testing.assert_array_equal(got["a"], want["a"])
    np.testing.assert_array_equal(got["b"], want["b"])


def test_fallback_routes_respect_output(api):
    """transform_features / symmetry_scores on shapes without a fused epilogue used to fail for output='numpy'."""
    rng = np.random.default_rng(11)
    p = rng.random((50, 10, 10), dtype=np.float32)
    z = api.ZPs(6, 10, precision="fp32", output="numpy")
    n, m, v = zo.zernike_basis(6, 10)
    ref = zo.project_patches(p.astype(np.float64), v)
    mag = z.transform_features(p, "abs")
    assert np.abs(mag.cpu().numpy() - np.abs(zo.to_complex(ref, n, m)[0])).max() <= 3e-6 * np.abs(ref).max()
    s = z.symmetry_scores(p, list(range(2, 12)))                    # 10 folds: not fusable
    assert isinstance(s, np.ndarray) and np.abs(s - zo.rot_maps(ref, n, m, list(range(2, 12)))).max() < 1e-5


def _nccl_worker(rank, world, port, out_dir):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from motif_learn_b200 import features, parallel as par
        rng = np.random.default_rng(0)
        patches = torch.from_numpy(rng.random((1001, 32, 32), dtype=np.float32)).cuda()       # replicated input
        z = features.ZPs(10, 32)
        feats = par.transform_patches_sharded(lambda x: z.transform(x).data, patches)          # gather_ragged on NCCL
        lo, hi = par.shard_range(1001, rank, world)
        rows = par.gather_rows(z.transform(patches[lo:hi]).data)                               # ragged gather_rows
        img = torch.from_numpy(rng.random((200, 160), dtype=np.float32)).cuda()
        smap = par.symmetry_map_sharded(lambda r0, r: z.symmetry_map(img, [2, 3], row0=r0, rows=r), 200)
        full = z.symmetry_map(img, [2, 3])
        whole = z.transform(patches).data
        ok = bool(torch.equal(feats, whole) and torch.equal(rows, whole) and torch.equal(smap, full))
        # K5 fused into K3: the projection kernel writes its rows into every rank's copy (CUDA IPC peer memory)
        big = torch.from_numpy(np.random.default_rng(5).random((70001, 32, 32), dtype=np.float32)).cuda()
        lo, hi = par.shard_range(70001, rank, world)
        for kind, cols in (("real", 66), ("abs", 36), ("complex", 72)):
            arr = par.PeerArray(70001, cols)
            arr.begin()
            z.transform_allgather(big[lo:hi], arr, lo, kind)
            arr.fence()
            torch.cuda.synchronize()
            if kind == "real":
                want = z.transform(big).data
            else:
                want = z.transform_features(big, kind)
                want = torch.view_as_real(want).reshape(70001, cols) if kind == "complex" else want
            ok = ok and bool(torch.equal(arr.local, want))
            arr.close()
        # score bands of one frame through the copy engines
        arr = par.PeerArray(2 * 200, 160)
        r0, rr = par.row_band(200, rank, world)
        arr.begin()
        par.push_score_bands(arr, z.symmetry_map(img, [2, 3], row0=r0, rows=rr), r0, 200)
        arr.fence()
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(arr.local.view(2, 200, 160), full))
        arr.local.zero_()
        arr.begin()
        par.symmetry_map_allgather(z, img, [2, 3], arr, r0, rr, 200, n_sub=3)      # sub-bands overlapped with the copies
        arr.fence()
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(arr.local.view(2, 200, 160), full))
        arr.close()
        torch.save({"ok": ok, "n": int(feats.shape[0])}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_feature_gather_on_nccl(torch, tmp_path):
    """K5 on real GPUs: shards of moments and score bands gathered with NCCL equal the single-GPU result bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = torch.load(str(tmp_path / f"r{r}.pt"))
        assert got["ok"] and got["n"] == 1001


def test_series_features_threads_and_streams(api, torch):
    """BASELINE config 5 pipeline: frames spread over worker threads / streams give exactly the per-frame results."""
    from motif_learn_b200.datasets import honeycomb_image
    frames = []
    for seed in range(7):
        img, _ = honeycomb_image((384, 448), bond=12.0, seed=seed, angle=7.0 * seed, jitter=0.3, noise=0.01)
        frames.append(img)
    dev = torch.from_numpy(np.stack(frames)).cuda()
    z = api.ZPs(12, 48)
    n, m, v = zo.zernike_basis(12, 48)
    feats, pts = api.series_features(z, dev, 5.0, 0.3, kind="abs", workers=3)
    torch.cuda.synchronize()
    assert len(feats) == len(pts) == 7
    for f in range(7):
        want_pts = zo.clear_border(zo.local_max(frames[f], 5.0, 0.3), frames[f].shape, 48)
        np.testing.assert_array_equal(pts[f], want_pts)
        ref = np.abs(zo.to_complex(zo.project_patches(zo.extract_patches(frames[f], want_pts, 48).astype(np.float64), v), n, m)[0])
        assert feats[f].is_cuda and np.abs(feats[f].cpu().numpy() - ref).max() <= 3e-6 * ref.max()
    one, pts1 = api.series_features(z, frames, 5.0, 0.3, kind="real", workers=1)      # numpy frames, single worker
    torch.cuda.synchronize()
    for f in range(7):
        np.testing.assert_array_equal(pts1[f], pts[f])
        assert one[f].data.shape == (len(pts[f]), 91)
    assert api.series_features(z, [], 5.0) == ([], [])


def test_lattice_and_tmd_generators_golden(golden, torch):
    """Row f1, second part: GPU HoneyCombLattice (coordinates + frame) and TMDImageSimulator.simulate against the
    live reference's outputs (tests/golden/datasets.npz)."""
    from motif_learn_b200.datasets import HoneyCombLattice, TMDImageSimulator
    g = golden("datasets.npz")
    for tag, kw in (("a", dict(size=160, l=12, seed=3, angle=17.0, jitter=0.2)),
                    ("b", dict(size=200, l=10, seed=1, angle=-33.0, random_shift=False))):
        lat = HoneyCombLattice(**kw)
        ca, cb = lat.coordinates()
        np.testing.assert_allclose(ca, g[f"lat_{tag}_A"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(cb, g[f"lat_{tag}_B"], rtol=0, atol=1e-10)
        pa, pb = lat.get_points()
        assert pa.shape == g[f"lat_{tag}_ptsA"].shape and pb.shape == g[f"lat_{tag}_ptsB"].shape
        np.testing.assert_allclose(pa, g[f"lat_{tag}_ptsA"], rtol=0, atol=1e-10)
        img = lat.to_image(normalize=(tag == "b"))
        assert img.dtype == np.float32 and img.shape == g[f"lat_{tag}_img"].shape
        np.testing.assert_allclose(img, g[f"lat_{tag}_img"], rtol=0, atol=4e-7)
        assert lat.to_image(as_tensor=True).is_cuda
    with pytest.raises(ValueError):
        HoneyCombLattice(size=64, l=12.0, a=5.0)
    for tag, kw, basis in (("a", dict(size=(96, 128), a=20, theta=7.0), None),
                           ("b", dict(size=(80, 80), a=16, theta=30.0), [(0.0, 0.0, 'TM'), (1 / 3, 1 / 3, 'X'), (2 / 3, 2 / 3, 'X')])):
        sim = TMDImageSimulator(basis=basis, **kw)
        sim.add_random_vacancies('X', 3, 2, seed=0)
        sim.add_random_dopants('TM', 4, seed=1)
        img, masks = sim.simulate(return_masks=True)
        assert img.dtype == np.float32
        np.testing.assert_allclose(img, g[f"tmd_{tag}_img"], rtol=0, atol=3e-7)     # FFT round-off + float32 sum order
        for label, m in masks.items():
            np.testing.assert_array_equal(m, g[f"tmd_{tag}_mask_{label}"])
        np.testing.assert_array_equal(sim.lbs, g[f"tmd_{tag}_lbs"])
        assert [str(v) for v in sim.labels] == [str(v) for v in g[f"tmd_{tag}_labels"]]
        np.testing.assert_allclose(sim.pts, g[f"tmd_{tag}_pts"], rtol=0, atol=1e-10)


def test_cluster_labels_golden(golden, torch):
    """Row f4: kmeans_lbs / gmm_lbs on the GPU give the labels of the live reference (scikit-learn, random_state=0):
    identical seeding (same RandomState stream), Lloyd / EM passes in float64 on the device."""
    from motif_learn_b200.clustering import gmm_lbs, kmeans_lbs, sort_lbs
    g = golden("clustering.npz")
    x = g["feats"]
    for data in (x, torch.from_numpy(x).cuda()):
        got = kmeans_lbs(data, 3)
        assert isinstance(got, np.ndarray) and got.dtype == np.int64
        np.testing.assert_array_equal(got, g["kmeans3"])
    np.testing.assert_array_equal(kmeans_lbs(x, 5), g["kmeans5"])     # clusters that split a family: same trajectory
    np.testing.assert_array_equal(gmm_lbs(x, 3), g["gmm3"])
    np.testing.assert_array_equal(sort_lbs(np.array([7, 7, 2, 7, 2, 9])), np.array([0, 0, 1, 0, 1, 2]))
    big = np.concatenate([x] * 40) + np.random.default_rng(0).normal(0, 1e-4, (40 * len(x), x.shape[1])).astype(np.float32)
    np.testing.assert_array_equal(kmeans_lbs(big, 3), zo.kmeans_lbs(big.astype(np.float64), 3))     # 40 000 x 36 vs scikit-learn here
    with pytest.raises(ValueError):
        kmeans_lbs(x, 0)
    with pytest.raises(NotImplementedError):
        gmm_lbs(x, 3, type="diag")


def test_rot_maps_on_complex_moments_documented_difference(api, golden, torch):
    """The reference squares COMPLEX numbers when rot_maps is called on complex moments (its output is complex);
    here complex moments are converted back to the real representation first, so the scores are those of the real
    moments.  Pinned with the live reference's own output so that nobody is surprised."""
    g = golden("clustering.npz")
    zc = api.zmoments(g["cdata"], g["cn"], g["cm"])
    ours = zc.rot_maps([2, 3, 4, 6])
    np.testing.assert_allclose(ours, g["rot_on_real"], rtol=0, atol=1e-12)
    assert np.abs(ours - g["rot_on_complex"]).max() > 0.1


def test_denoise_svd_golden(golden, torch):
    """Row f4: patch-SVD denoiser -- gather kernel on the strided grid, randomized SVD with library linear algebra,
    overlap-add kernel -- against the live reference's frame (one draw of ITS randomized SVD) and the exact rank-r one."""
    from motif_learn_b200.denoise import DenoiseSVD, denoise_svd, extract_patches, reconstruct_patches
    g = golden("denoise.npz")
    img = g["img"]
    p = extract_patches(img, 16, 5)
    assert p.shape == tuple(g["patch_shape"]) and p.dtype == np.float32
    np.testing.assert_array_equal(p[:3], g["patch_head"].astype(np.float32))
    np.testing.assert_array_equal(p, zo.denoise_extract(img, 16, 5).astype(np.float32))
    rec = reconstruct_patches(zo.denoise_extract(img, 16, 5), img.shape, 5)
    np.testing.assert_allclose(rec, g["rec"], rtol=0, atol=1e-13)
    clean, s = denoise_svd(img, 16, 6, extraction_step=5, verbose=False, return_s=True)
    assert clean.dtype == np.float64 and clean.shape == img.shape
    np.testing.assert_allclose(clean, g["exact"], rtol=0, atol=5e-6)          # float32 gather of the float64 frame
    np.testing.assert_allclose(clean, g["clean"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(s, g["s"], rtol=1e-6)
    job = DenoiseSVD(torch.from_numpy(img).cuda(), 6, 16, 5)
    out = job.run()
    assert out.is_cuda and np.abs(out.cpu().numpy() - g["exact"]).max() < 5e-6
    with pytest.raises(ValueError):
        denoise_svd(img, 200, 3)


# ---- fp16-split projection (round 2) ---------------------------------------------------------------------------
@pytest.mark.parametrize("n_max,size,count", [(12, 64, 1500), (20, 64, 300), (10, 32, 700), (12, 48, 257), (3, 6, 5),
                                              (13, 128, 200), (14, 128, 150), (0, 64, 130)])
def test_projection_f16x3_vs_oracle(api, torch, n_max, size, count):
    """ZB200_PREC_F16X3 on patch stacks (value_max given): the strict fp32-grade gate against the oracle, every
    fused epilogue, and invariance under the scale of the data (the power-of-two input scale follows value_max)."""
    rng = np.random.default_rng(n_max * 1000 + size)
    patches = rng.random((count, size, size), dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(n_max, size)
    ref = zo.project_patches(patches.astype(np.float64), v)
    refc = zo.to_complex(ref, n, m)[0]
    z = api.ZPs(n_max, size, value_max=1.0)
    from motif_learn_b200 import _lib
    if not _lib.load().zb200_plan_supports(z._plan, _lib.PREC_F16X3, _lib.OUT_REAL):
        pytest.skip("no tensor-core projection for this shape")
    assert z._precision_code(device_stack=True) == _lib.PREC_F16X3
    dev = torch.from_numpy(patches).cuda()
    fp32_close(z.transform(dev).data.cpu().numpy(), ref)
    zc = z.transform_features(dev, "complex").cpu().numpy()
    fp32_close(np.concatenate([zc.real, zc.imag], axis=1), np.concatenate([refc.real, refc.imag], axis=1))
    mag, ph = z.transform_features(dev, "abs_phase")
    assert np.abs(mag.cpu().numpy() - np.abs(refc)).max() <= 3e-6 * np.abs(refc).max()
    big = np.abs(refc) > 1e-3 * np.abs(refc).max()
    assert np.abs(np.angle(np.exp(1j * (ph.cpu().numpy() - np.angle(refc))))[big]).max() < 1e-3
    # same numbers on rescaled data with a matching bound; the kernel equals tf32x3 to fp32-grade, not bit for bit
    for scale in (3.7e5, 2.1e-7):
        zs = api.ZPs(n_max, size, value_max=scale)
        fp32_close(zs.transform(dev * scale).data.cpu().numpy() / scale, ref)
    # a bound that is too small by more than 4x is LOUD: inf / NaN, never a silently wrong number
    bad = api.ZPs(n_max, size, value_max=1e-6).transform(dev).data
    assert not torch.isfinite(bad).all()
    # without value_max 'auto' is tf32x3 -- or, for the windows the mirror-folded kernel serves, its auto-ranged form;
    # numpy input keeps the host pipeline
    auto = bool(_lib.load().zb200_plan_supports_autorange(z._plan))
    assert auto == (size % 64 == 0 and n_max <= 20)
    assert api.ZPs(n_max, size)._precision_code(device_stack=True) == (_lib.PREC_F16X3 if auto else _lib.PREC_TF32X3)
    fp32_close(api.ZPs(n_max, size).transform(dev).data.cpu().numpy(), ref)
    fp32_close(z.transform(patches).data, ref)
    from sklearn.base import clone
    assert clone(z).value_max == 1.0
    with pytest.raises(ValueError):
        api.ZPs(n_max, size, value_max=-1.0)


@pytest.mark.parametrize("n_max,count", [(12, 5000), (20, 2100), (4, 129)])
def test_folded_projection_auto_range_and_fallback(api, torch, n_max, count):
    """64-pixel windows without value_max: the mirror-folded kernel scales by the largest |x| of 1024 evenly spread
    patches.  Same results as with a bound; values up to 8x the sampled maximum still convert; beyond that the
    overflow flag makes the tf32x3 kernel behind it recompute the stack (nothing silent, no host round trip)."""
    size = 64
    rng = np.random.default_rng(7 + n_max)
    patches = rng.random((count, size, size), dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(n_max, size)
    z = api.ZPs(n_max, size)
    from motif_learn_b200 import _lib
    assert _lib.load().zb200_plan_supports_autorange(z._plan) == 1
    sampled = {(i * count) // min(count, 1024) for i in range(min(count, 1024))}
    hidden = next(i for i in range(count) if i not in sampled) if len(sampled) < count else None
    for boost in (1.0, 6.0, 300.0):
        x = patches.copy()
        if hidden is None and boost > 1.0:
            continue
        if hidden is not None:
            x[hidden] *= boost                              # a patch the range sample never sees
        ref = zo.project_patches(x.astype(np.float64), v)
        refc = zo.to_complex(ref, n, m)[0]
        dev = torch.from_numpy(x).cuda()
        fp32_close(z.transform(dev).data.cpu().numpy(), ref)
        zc = z.transform_features(dev, "complex").cpu().numpy()
        fp32_close(np.concatenate([zc.real, zc.imag], axis=1), np.concatenate([refc.real, refc.imag], axis=1))
        mag = z.transform_features(dev, "abs").cpu().numpy()
        assert np.abs(mag - np.abs(refc)).max() <= 3e-6 * np.abs(refc).max()
    # the host pipelines take the same route
    fp32_close(z.transform(patches[:700]).data, zo.project_patches(patches[:700].astype(np.float64), v))
    # tiny and huge data scales
    for scale in (1e-20, 1e20):
        dev = torch.from_numpy(patches[:300] * np.float32(scale)).cuda()
        fp32_close(z.transform(dev).data.cpu().numpy() / scale, zo.project_patches(patches[:300].astype(np.float64), v))
    assert torch.equal(z.transform(torch.zeros((3, size, size), device="cuda")).data, torch.zeros((3, len(n)), device="cuda"))


def test_symmetry_map_host_route_in_bands(api, torch):
    """numpy frame -> float64 score maps through zb200_symmetry_map_host (frame up once, row bands, each band's
    download + widening overlapped with the next band): bit-identical to the single device call, NaN pattern included."""
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image((333, 290), bond=11.0, seed=4, angle=7.0, jitter=0.2, noise=0.01)
    z = api.ZPs(12, 48)
    for folds, p in (([2, 3, 4, 6], 2), ([3], 1)):
        host = z.symmetry_map(img, folds, p=p)
        assert isinstance(host, np.ndarray) and host.dtype == np.float64 and host.shape == (len(folds), 333, 290)
        dev = z.symmetry_map(torch.from_numpy(img).cuda(), folds, p=p)
        np.testing.assert_array_equal(host.astype(np.float32), dev.cpu().numpy())
    pageable = np.array(img)                                  # a second call reuses the staging buffers
    np.testing.assert_array_equal(z.symmetry_map(pageable, [2, 3, 4, 6]), host if len(folds) == 4 else z.symmetry_map(img, [2, 3, 4, 6]))


def test_folded_projection_ragged_counts_and_properties(api, torch):
    """Tile edges of the folded kernel (128-patch tiles, CTA pairs: counts around 1, 128 and 256, an odd number of
    tiles), every epilogue, linearity, and invariance under a permutation of the patches."""
    rng = np.random.default_rng(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(12, 64)
    z, zh = api.ZPs(12, 64), api.ZPs(12, 64, value_max=1.0)
    base = rng.random((700, 64, 64), dtype=np.float32)
    ref_all = zo.project_patches(base.astype(np.float64), v)
    for count in (1, 2, 127, 128, 129, 255, 256, 257, 385, 700):
        dev = torch.from_numpy(base[:count]).cuda()
        for tr in (z, zh):
            got = tr.transform(dev).data
            assert got.shape == (count, 91)
            fp32_close(got.cpu().numpy(), ref_all[:count])
        refc = zo.to_complex(ref_all[:count], n, m)[0]
        mag, ph = zh.transform_features(dev, "abs_phase")
        assert np.abs(mag.cpu().numpy() - np.abs(refc)).max() <= 3e-6 * np.abs(refc).max()
    dev = torch.from_numpy(base).cuda()
    perm = torch.from_numpy(rng.permutation(700)).cuda()
    assert torch.equal(zh.transform(dev[perm].contiguous()).data, zh.transform(dev).data[perm])    # a patch's result does not depend on its tile
    a, b = dev[:300], dev[300:600]
    za, zb, zab = (zh.transform(t).data for t in (a, b, (0.5 * a + 0.25 * b).contiguous()))
    assert (zab - (0.5 * za + 0.25 * zb)).abs().max().item() < 2e-6 * float(np.abs(ref_all).max())


def test_mirror_map_in_row_bands(api, golden, torch):
    """ZPs.mirror_map streams the frame through K4 + the mirror kernel band by band (the (M,H,W) maps never exist in
    full): equal to the materialised route bit for bit, and to the live reference's mirror_map at the golden pixels."""
    g = golden("lattice.npz")
    img = g["map_img"]
    z = api.ZPs(12, 48)
    dimg = torch.from_numpy(img.astype(np.float32)).cuda()
    full = z.transform(dimg).mirror_map()
    for band in (64, 50, 7):
        assert torch.equal(z.mirror_map(dimg, band_rows=band), full)
    host = z.mirror_map(img.astype(np.float32))
    assert isinstance(host, np.ndarray) and host.shape == img.shape
    np.testing.assert_allclose(host[g["map_ys"], g["map_xs"]], g["map_mirror_pts"], rtol=0, atol=2e-5)


def test_release_plans_rebuilds_on_demand(api):
    z = api.ZPs(6, 16)
    a = z.transform(np.ones((3, 16, 16), dtype=np.float32)).data
    assert api.release_plans() >= 1
    assert api.release_plans() == 0
    b = api.ZPs(6, 16).transform(np.ones((3, 16, 16), dtype=np.float32)).data       # the plan is rebuilt
    np.testing.assert_array_equal(a, b)
