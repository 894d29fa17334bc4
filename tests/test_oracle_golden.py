"""Pin the CPU oracle (oracle/zernike_oracle.py) to the reference.

(i) the reference's own pinned answers (tests/features/test_zmoments.py:5-88 of
jiadongdan/motif-learn), restated; (ii) golden vectors produced by the REAL
reference in the build container (oracle/make_goldens.py).  CPU only.
"""
import numpy as np
import pytest

import zernike_oracle as zo


# ---- (i) the reference's own known answers --------------------------------- #
def test_nm2j_reference_known_answers():
    for (n, m), j in {(0, 0): 0, (1, -1): 1, (2, 0): 4, (3, 1): 8, (4, -4): 10, (5, 3): 19}.items():
        assert zo.nm2j(n, m) == j
    np.testing.assert_array_equal(zo.nm2j([0, 1, 2, 2, 3], [0, -1, 0, 2, 3]), [0, 1, 4, 5, 9])
    assert zo.nm2j(1000, 1000) == ((1000 + 2) * 1000 + 1000) // 2
    assert isinstance(zo.nm2j(2, 0), int) and isinstance(zo.nm2j([2], [0]), np.ndarray)


@pytest.mark.parametrize("args,msg", [
    ((-1, 0), "Radial order `n` must be non-negative."),
    ((2, 3), r"Azimuthal frequency `m` must satisfy \|m\| ≤ n."),
    ((1, 0), r"`n - \|m\|` must be even."),
    (([1, 2], [0]), "`n` and `m` must have the same shape."),
])
def test_nm2j_reference_errors(args, msg):
    with pytest.raises(ValueError, match=msg):
        zo.nm2j(*args)


def test_select_reference_ordering():
    m = np.array([0, -1, 1, -2, 2, 3])
    np.testing.assert_array_equal(zo.select_indices(m, [1, -2]), [1, 2, 3, 4])


# ---- (ii) golden vectors from the live reference ---------------------------- #
def test_index_golden(golden):
    g = golden("index.npz")
    np.testing.assert_array_equal(zo.nm2j(g["n"], g["m"]), g["j"])
    np.testing.assert_array_equal(zo.nm2j_complex(g["n"], np.abs(g["m"])), g["jc"])
    n12, m12 = zo.mode_table(12)
    np.testing.assert_array_equal(n12, g["n12"])
    np.testing.assert_array_equal(m12, g["m12"])
    np.testing.assert_array_equal(zo.complex_matrix(n12, m12), g["cmat12"])
    np.testing.assert_array_equal(zo.rot_weights([1, 2, 3], [2, 3, 4, 6]), g["rotmat_a"])
    np.testing.assert_array_equal(zo.rot_weights([2, 3, 4, 6], m12), g["rotmat_12"])


@pytest.mark.parametrize("n_max,size", [(4, 8), (6, 9), (5, 11), (10, 32)])
def test_basis_full_golden(golden, n_max, size):
    g = golden("basis.npz")
    _, _, v = zo.zernike_basis(n_max, size)
    ref = g[f"full_{n_max}_{size}"]
    assert v.shape == ref.shape
    np.testing.assert_array_equal(v == 0, ref == 0)          # identical disk mask
    np.testing.assert_allclose(v, ref, rtol=0, atol=1e-13)


@pytest.mark.parametrize("n_max,size", [(12, 48), (12, 64), (20, 64), (12, 33)])
def test_basis_sampled_golden(golden, n_max, size):
    g = golden("basis.npz")
    _, _, v = zo.zernike_basis(n_max, size)
    flat = v.ravel()
    np.testing.assert_allclose(flat[g[f"idx_{n_max}_{size}"]], g[f"val_{n_max}_{size}"], rtol=0, atol=1e-12)
    s, a, nz = g[f"stat_{n_max}_{size}"]
    assert abs(flat.sum() - s) < 1e-8 and abs(np.abs(flat).sum() - a) < 1e-7
    assert np.count_nonzero(v[0]) == int(nz)


def test_basis_exact_yardstick():
    # exact rational evaluation of the radial polynomial at sampled pixels: the
    # reference algorithm (float factorial power sum) is within 1e-11 of it at
    # n_max=12 and drifts to ~3e-9 at n_max=20 (parity bound).
    pts = [(r, c) for r in range(0, 48, 5) for c in range(1, 48, 7)]
    _, _, ref = zo.zernike_basis(12, 48)
    _, _, ex = zo.zernike_basis_exact(12, 48, pts)
    got = np.stack([ref[:, r, c] for r, c in pts], axis=1)
    assert np.abs(got - ex).max() < 1e-11
    pts = [(r, c) for r in range(0, 64, 9) for c in range(2, 64, 11)]
    _, _, ref = zo.zernike_basis(20, 64)
    _, _, ex = zo.zernike_basis_exact(20, 64, pts)
    got = np.stack([ref[:, r, c] for r, c in pts], axis=1)
    assert 1e-12 < np.abs(got - ex).max() < 1e-8


def test_patches_golden(golden):
    g = golden("patches_nfold.npz")
    p = g["patches"]
    n, m, v = zo.zernike_basis(12, 64)
    z = zo.project_patches(p, v)
    np.testing.assert_allclose(z, g["z12_data"], rtol=1e-12, atol=1e-15)
    zc, n_c, m_c = zo.to_complex(z, n, m)
    np.testing.assert_array_equal(n_c, g["z12_cn"])
    np.testing.assert_array_equal(m_c, g["z12_cm"])
    np.testing.assert_allclose(zc, g["z12_cdata"], rtol=1e-12, atol=1e-15)
    zr, n_r, m_r = zo.to_real(zc, n_c, m_c)
    np.testing.assert_array_equal(n_r, n)
    np.testing.assert_array_equal(m_r, m)
    np.testing.assert_allclose(zr, g["z12_real_back"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(zo.rot_maps(z, n, m, [2, 3, 4, 6]), g["z12_rot"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(zo.rot_maps(z, n, m, [3, 6], p=1), g["z12_rot_p1"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(zo.rot_maps(z, n, m, [2, 4], m_unselect=(0, 1, 2)), g["z12_rot_unsel"],
                               rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(zo.normalize(z), g["z12_norm2"], rtol=1e-12)
    np.testing.assert_allclose(zo.normalize(z, 1), g["z12_norm1"], rtol=1e-12)
    np.testing.assert_allclose(zo.normalize(z, np.inf), g["z12_norminf"], rtol=1e-12)
    np.testing.assert_allclose(zo.rotate(z, n, m, 30.0)[0], g["z12_rot30"], rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(zo.mirror_map(z, n, m), g["z12_mirror"], rtol=1e-11)
    keep = zo.select_indices(m, [2, -3])
    np.testing.assert_array_equal(m[keep], g["z12_sel_m"])
    np.testing.assert_allclose(z[:, keep], g["z12_sel_data"], rtol=1e-12, atol=1e-15)
    keep = zo.select_indices(m, [0, 1], invert=True)
    np.testing.assert_array_equal(m[keep], g["z12_unsel_m"])
    np.testing.assert_array_equal(n[keep], g["z12_unsel_n"])
    # survey KATs
    np.testing.assert_allclose(z[0, :4], [2.985956997797e-01, 6.946528841105e-03, 6.458243044575e-03,
                                          2.461357425825e-03], rtol=1e-10)
    # n_max = 20 (config-3 family)
    n20, m20, v20 = zo.zernike_basis(20, 64)
    z20 = zo.project_patches(p, v20)
    np.testing.assert_allclose(z20, g["z20_data"], rtol=1e-11, atol=1e-14)
    zc20, _, _ = zo.to_complex(z20, n20, m20)
    np.testing.assert_allclose(np.abs(zc20), g["z20_cabs"], rtol=1e-11, atol=1e-14)


def test_gather_golden(golden):
    g = golden("lattice.npz")
    img, pts = g["img"], g["pts"]
    import hashlib
    for k in (32, 33):
        kept = zo.clear_border(pts, img.shape, k)
        np.testing.assert_array_equal(kept, g[f"kept_{k}"])
        patches = zo.extract_patches(img, kept, k)
        assert patches.dtype == img.dtype and patches.shape[1:] == (k, k)
        assert hashlib.sha256(np.ascontiguousarray(patches).tobytes()).hexdigest()[:16] == str(g[f"patch_sha_{k}"])
        np.testing.assert_array_equal(patches[:4], g[f"patch_head_{k}"])
    flat = zo.extract_patches(img, zo.clear_border(pts, img.shape, 32), 32, flat=True)
    np.testing.assert_array_equal(flat.shape, g["flat_shape_32"])


def test_lattice_patch_moments_golden(golden):
    g = golden("lattice.npz")
    img, pts = g["img"], g["pts"]
    patches = zo.extract_patches(img, zo.clear_border(pts, img.shape, 32), 32)[:96]
    n, m, v = zo.zernike_basis(10, 32)
    z = zo.project_patches(patches, v)
    np.testing.assert_allclose(z, g["z10_data"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(zo.to_complex(z, n, m)[0], g["z10_cdata"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(zo.rot_maps(z, n, m, [2, 3, 4, 6]), g["z10_rot"], rtol=1e-10, atol=1e-13)


def test_map_golden(golden):
    g = golden("lattice.npz")
    img = g["map_img"]
    n, m, v = zo.zernike_basis(12, 48)
    z = zo.moment_map_fft(img, v, n)
    ys, xs = g["map_ys"], g["map_xs"]
    np.testing.assert_allclose(z[:, ys, xs], g["map_moments"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(zo.rot_maps(z, n, m, [2, 3, 4, 6]), g["map_rot"], rtol=1e-8, atol=1e-11)
    np.testing.assert_array_equal(zo.valid_mask(img.shape, 48), g["map_valid"])
    np.testing.assert_allclose(np.abs(zo.to_complex(z, n, m)[0])[:, ys, xs], g["map_cabs_pts"], rtol=1e-9,
                               atol=1e-14)
    np.testing.assert_allclose(zo.mirror_map(z, n, m)[ys, xs], g["map_mirror_pts"], rtol=1e-8)
    # chunked FFT (per-mode independence) is equivalent
    zc = zo.moment_map_fft(img, v, n, chunk=16)
    np.testing.assert_allclose(zc, z, rtol=0, atol=1e-15)
    # direct-form definition agrees with the FFT route
    sub = img[:70, :80]
    zd = zo.moment_map_direct(sub, v)
    zf = zo.moment_map_fft(sub, v, n)
    np.testing.assert_allclose(zd, zf, rtol=0, atol=2e-14)


def test_map_odd_window_golden(golden):
    g = golden("lattice.npz")
    img = g["map2_img"]
    n, m, v = zo.zernike_basis(8, 33)
    z = zo.moment_map_fft(img, v, n)
    np.testing.assert_allclose(z[:, ::7, ::5], g["map2_moments"], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(zo.rot_maps(z, n, m, [3, 6]), g["map2_rot"], rtol=1e-8, atol=1e-11)
    np.testing.assert_array_equal(zo.valid_mask(img.shape, 33), g["map2_valid"])
    np.testing.assert_allclose(zo.moment_map_direct(img, v), z, rtol=0, atol=2e-14)


def test_map_equals_patch_projection(golden):
    # map[:, y, x] == transform(patch centred at (x, y))  
    g = golden("lattice.npz")
    img = g["map_img"]
    n, m, v = zo.zernike_basis(12, 48)
    z = zo.moment_map_fft(img, v, n)
    pts = np.array([[60.0, 50.0], [100.2, 80.7], [150.5, 120.5]])
    patches = zo.extract_patches(img, pts, 48)
    zp = zo.project_patches(patches, v)
    c = np.rint(pts).astype(int)
    np.testing.assert_allclose(z[:, c[:, 1], c[:, 0]].T, zp, rtol=0, atol=1e-14)


def test_render_golden(golden):
    g = golden("render.npz")
    img = zo.render_atoms((160, 160), g["pts"], g["amps"], float(g["sigma"]))
    np.testing.assert_array_equal(img, g["img"])
    img2 = zo.render_atoms((120, 160), g["pts"], g["amps2"], float(g["sigma"]))
    np.testing.assert_array_equal(img2, g["img2"])


# ---- peak detection ("next" row f2) ------------------------------------------------------------------
def test_oracle_local_max_matches_reference_golden(golden):
    """The oracle's local_max against the REAL reference filter run on the restated skimage candidates
    (oracle/make_goldens.py:golden_peaks): same peaks in the same (brightest-first) order."""
    g = golden("peaks.npz")
    for tag in ("a", "b"):
        img = g[f"img_{tag}"]
        np.testing.assert_array_equal(zo.peak_local_max_md1(img, None), g[f"cand_{tag}"])
        i = 0
        while f"pts_{tag}{i}" in g.files:
            r, thr = g[f"arg_{tag}{i}"]
            got = zo.local_max(img, float(r), None if np.isnan(thr) else float(thr))
            np.testing.assert_array_equal(got, g[f"pts_{tag}{i}"])
            i += 1


def test_oracle_peak_local_max_edge_cases():
    assert zo.peak_local_max_md1(np.full((9, 11), 3.0, dtype=np.float32), None).shape == (0, 2)   # trivial image
    assert zo.peak_local_max_md1(np.full((9, 11), 3.0, dtype=np.float32), 1.0).shape == (0, 2)
    img = np.zeros((7, 7), dtype=np.float32)
    img[0, 3] = 5.0          # on the excluded border
    img[3, 3] = img[3, 4] = 2.0   # a two-pixel plateau: both survive (ensure_spacing rejects only d < 1)
    got = zo.peak_local_max_md1(img, None)
    np.testing.assert_array_equal(got, [[3, 3], [3, 4]])
    np.testing.assert_array_equal(zo.local_max(img, 1.0), [[3, 3]])            # (x, y); the raster-first twin wins
    np.testing.assert_array_equal(zo.local_max(img, 0.5), [[3, 3], [4, 3]])


# ---- PCA ("next" row f4) ---------------------------------------------------------------------------
def test_oracle_pca_matches_reference_golden(golden):
    g = golden("pca.npz")
    for c, key in ((2, "pca2"), (5, "pca5")):
        np.testing.assert_allclose(zo.pca(g["feats"], c), g[key], rtol=0, atol=1e-13)


def test_lattice_coords_and_tmd_oracle_match_live_reference(golden):
    """Row f1, second part: the oracle's restatement of HoneyCombLattice._generate_coordinates and of one species of
    TMDImageSimulator.simulate against tests/golden/datasets.npz (outputs of the unmodified reference)."""
    g = golden("datasets.npz")
    for tag, kw in (("a", dict(size=160, l=12, seed=3, angle=17.0, jitter=0.2)),
                    ("b", dict(size=200, l=10, seed=1, angle=-33.0, random_shift=False))):
        ca, cb, _ = zo.honeycomb_coords(**kw)
        np.testing.assert_allclose(ca, g[f"lat_{tag}_A"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(cb, g[f"lat_{tag}_B"], rtol=0, atol=1e-10)
    # TMD frame a: rebuild the two species from the golden masks' atom positions is not possible (masks are rounded),
    # so the oracle blur is checked on the golden delta images themselves: blur(delta) summed == golden frame
    from scipy.signal import fftconvolve
    total = np.zeros((96, 128), dtype=np.float32)
    for label, sigma, amp in (("TM", 2.0, 1.0), ("X", 1.5, 0.6)):
        delta = g[f"tmd_a_mask_{label}"]
        ys, xs = np.nonzero(delta)
        blurred, again = zo.tmd_blur((96, 128), np.stack([xs, ys], axis=1).astype(float), delta[ys, xs], sigma, amp)
        np.testing.assert_array_equal(again, delta)
        total += blurred
    np.testing.assert_allclose(total, g["tmd_a_img"], rtol=0, atol=2e-7)


def test_cluster_labels_oracle_matches_live_reference(golden):
    """Row f4: the oracle's scikit-learn calls give the labels the reference's kmeans_lbs / gmm_lbs gave."""
    g = golden("clustering.npz")
    x = g["feats"].astype(np.float64)
    np.testing.assert_array_equal(zo.kmeans_lbs(x, 3), g["kmeans3"])
    np.testing.assert_array_equal(zo.gmm_lbs(x, 3), g["gmm3"])
    np.testing.assert_array_equal(g["kmeans3"], g["truth"])            # the three patch families, largest first
    # the reference's rot_maps on COMPLEX moments squares complex numbers (its result is complex, not a score)
    assert np.iscomplexobj(g["rot_on_complex"]) and np.abs(g["rot_on_complex"].imag).max() > 0.1


def test_denoise_svd_oracle_matches_live_reference(golden):
    g = golden("denoise.npz")
    img = g["img"]
    p = zo.denoise_extract(img, 16, 5)
    assert tuple(g["patch_shape"]) == p.shape
    np.testing.assert_array_equal(p[:3], g["patch_head"])
    np.testing.assert_allclose(zo.denoise_reconstruct(p, img.shape, 5), g["rec"], rtol=0, atol=1e-13)
    clean, s = zo.denoise_svd_exact(img, 16, 6, 5)
    np.testing.assert_allclose(clean, g["exact"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(clean, g["clean"], rtol=0, atol=1e-6)        # the reference's randomized draw
    np.testing.assert_allclose(s, g["s"], rtol=1e-9)
