"""Generate tests/golden/*.npz from the REAL reference (run in the build container).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_goldens.py
Imports jiadongdan/motif-learn from /root/reference through oracle/ref_shim.py
(the reference cannot travel to the GPU box), runs its own public API on small
seeded inputs produced by its own dataset generators and stores inputs+outputs.
Every array in the fixtures is an output of unmodified reference code.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.load_reference()
from mtflearn.features import ZPs, zmoments, KeyPoints  # noqa: E402
from mtflearn.features import construct_rot_maps_matrix, construct_complex_matrix, nm2j, nm2j_complex  # noqa: E402
from mtflearn.datasets import HoneyCombLattice, get_zps_test_patches  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays")


def golden_index():
    n = np.array([0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 7, 12, 20])
    m = np.array([0, -1, 1, -2, 0, 2, -3, -1, 1, 3, -4, -2, 0, 2, 4, -5, 6, -20])
    z12 = ZPs(12, 16)
    save("index.npz",
         n=n, m=m, j=nm2j(n, m), jc=nm2j_complex(n, np.abs(m)),
         n12=z12.n, m12=z12.m,
         cmat12=construct_complex_matrix(z12.n, z12.m),
         rotmat_a=construct_rot_maps_matrix([1, 2, 3], [2, 3, 4, 6]),
         rotmat_12=construct_rot_maps_matrix([2, 3, 4, 6], z12.m))


def golden_basis():
    arrays = {}
    for n_max, size in [(4, 8), (6, 9), (5, 11), (10, 32)]:
        z = ZPs(n_max, size)
        arrays[f"full_{n_max}_{size}"] = z.polynomials
    rng = np.random.default_rng(1234)
    for n_max, size in [(12, 48), (12, 64), (20, 64), (12, 33)]:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            z = ZPs(n_max, size)
        flat = z.polynomials.ravel()
        idx = rng.choice(flat.size, size=6000, replace=False)
        arrays[f"idx_{n_max}_{size}"] = idx
        arrays[f"val_{n_max}_{size}"] = flat[idx]
        arrays[f"stat_{n_max}_{size}"] = np.array([flat.sum(), np.abs(flat).sum(),
                                                    float(np.count_nonzero(z.polynomials[0]))])
    save("basis.npz", **arrays)


def _zm_outputs(prefix, o, arrays, full=True):
    oc = o.to_complex()
    arrays[prefix + "data"] = o.data
    arrays[prefix + "n"] = o.n
    arrays[prefix + "m"] = o.m
    arrays[prefix + "cdata"] = oc.data
    arrays[prefix + "cn"] = oc.n
    arrays[prefix + "cm"] = oc.m
    arrays[prefix + "rot"] = o.rot_maps([2, 3, 4, 6])
    arrays[prefix + "rot_p1"] = o.rot_maps([3, 6], p=1)
    arrays[prefix + "rot_unsel"] = o.rot_maps([2, 4], m_unselect=(0, 1, 2))
    if full:
        arrays[prefix + "real_back"] = oc.to_real().data
        arrays[prefix + "norm2"] = o.normalize().data
        arrays[prefix + "norm1"] = o.normalize(order=1).data
        arrays[prefix + "norminf"] = o.normalize(order=np.inf).data
        arrays[prefix + "rot30"] = o.rotate(30.0).data
        arrays[prefix + "mirror"] = o.mirror_map()
        sel = o.select([2, -3])
        arrays[prefix + "sel_data"] = sel.data
        arrays[prefix + "sel_m"] = sel.m
        uns = o.unselect([0, 1])
        arrays[prefix + "unsel_m"] = uns.m
        arrays[prefix + "unsel_n"] = uns.n


def golden_patches():
    arrays = {}
    p3 = get_zps_test_patches(size=64, n_fold=3, num_patches=10)
    p4 = get_zps_test_patches(size=64, n_fold=4, num_patches=4, include_center=False)
    p = np.concatenate([p3, p4]).astype(np.float32)
    arrays["patches"] = p
    arrays["patches_sha"] = np.array(sha(p))
    o = ZPs(12, 64).transform(p)
    _zm_outputs("z12_", o, arrays)
    o20 = ZPs(20, 64).transform(p)
    arrays["z20_data"] = o20.data
    arrays["z20_cabs"] = np.abs(o20.to_complex().data)
    arrays["z20_cang"] = np.angle(o20.to_complex().data)
    save("patches_nfold.npz", **arrays)


def golden_lattice():
    arrays = {}
    lat = HoneyCombLattice(size=256, l=12, seed=0)
    img = lat.to_image()
    pts = np.vstack(lat.get_points())
    arrays["img"] = img
    arrays["img_sha"] = np.array(sha(img))
    arrays["pts"] = pts
    for k in (32, 33):
        kp = KeyPoints(pts, img, k)
        patches = kp.extract_patches()
        arrays[f"kept_{k}"] = kp.pts
        arrays[f"patch_sha_{k}"] = np.array(sha(patches))
        arrays[f"patch_head_{k}"] = patches[:4]
        arrays[f"patch_sum_{k}"] = patches.astype(np.float64).sum(axis=(1, 2))
        if k == 32:
            flat = KeyPoints(pts, img, k).extract_patches(flat=True)
            arrays["flat_shape_32"] = np.array(flat.shape)
            o = ZPs(10, 32).transform(patches[:96])
            _zm_outputs("z10_", o, arrays, full=False)
    # dense map: non-square crop, even window, n_max=12 (config-2 family)
    crop = img[:160, :192].astype(np.float64)
    arrays["map_img"] = crop
    o = ZPs(12, 48).transform(crop)
    rng = np.random.default_rng(7)
    ys = np.concatenate([rng.integers(0, 160, 500), [0, 0, 159, 159, 23, 24, 135, 136]])
    xs = np.concatenate([rng.integers(0, 192, 500), [0, 191, 0, 191, 23, 24, 167, 168]])
    arrays["map_ys"], arrays["map_xs"] = ys, xs
    arrays["map_moments"] = o.data[:, ys, xs]
    arrays["map_rot"] = o.rot_maps([2, 3, 4, 6])
    arrays["map_valid"] = o.valid_mask
    arrays["map_mirror_pts"] = o.mirror_map()[ys, xs]
    arrays["map_cabs_pts"] = np.abs(o.to_complex().data)[:, ys, xs]
    # odd window, smaller order
    crop2 = img[40:140, 30:150].astype(np.float64)
    o2 = ZPs(8, 33).transform(crop2)
    arrays["map2_img"] = crop2
    arrays["map2_moments"] = o2.data[:, ::7, ::5]
    arrays["map2_rot"] = o2.rot_maps([3, 6])
    arrays["map2_valid"] = o2.valid_mask
    save("lattice.npz", **arrays)




def golden_render():
    """Synthetic-frame renderer ("next" row f1): atoms + the frame the reference draws from them."""
    from mtflearn.datasets._tapered_gaussian import add_tapered_gaussian
    lat = HoneyCombLattice(size=160, l=12, seed=3, angle=17.0, jitter=0.2)
    img = lat.to_image()
    sigma = lat.l / 4.0
    radius = int(np.ceil(3 * sigma))

    def in_range(c):
        keep = ((c[:, 0] >= -radius) & (c[:, 0] <= lat.size - 1 + radius)
                & (c[:, 1] >= -radius) & (c[:, 1] <= lat.size - 1 + radius))
        return c[keep]

    a, b = in_range(lat._coords_A), in_range(lat._coords_B)
    pts = np.vstack([a, b])
    amps = np.concatenate([np.full(len(a), 1.0), np.full(len(b), 0.5)])
    rng = np.random.default_rng(0)
    amps2 = amps.copy()
    amps2[rng.random(len(amps2)) < 0.05] = 0.0
    amps2[rng.random(len(amps2)) < 0.05] *= 0.8
    img2 = np.zeros((120, 160), dtype=np.float32)
    add_tapered_gaussian(img2, pts, sigma, amplitude=amps2, r_factor=3.0)
    save("render.npz", pts=pts, amps=amps, sigma=np.array(sigma), img=img, amps2=amps2, img2=img2)


def golden_peaks():
    """Peak detection ("next" row f2).  The reference's ``local_max`` calls skimage's ``peak_local_max``, and
    scikit-image is absent from this image: the REAL ``mtflearn.features._local_max_v2.local_max`` (its own
    swap of the columns and its own ``filter_peaks_by_distance``) is run with the oracle's restatement of the
    skimage routine injected in place of the MagicMock stand-in.  Frames carry seeded noise so that no two
    candidate intensities are equal (the reference's visiting order is then unique)."""
    import mtflearn.features._local_max_v2 as lm
    import zernike_oracle as zo
    lm.peak_local_max = lambda image, min_distance=1, threshold_abs=None: zo.peak_local_max_md1(image, threshold_abs)
    arrays = {}
    rng = np.random.default_rng(7)
    lat = HoneyCombLattice(size=192, l=12, seed=5, angle=9.0, jitter=0.3)
    img_a = (lat.to_image() + rng.normal(0.0, 0.02, (192, 192))).astype(np.float32)
    img_b = rng.random((96, 131)).astype(np.float32)                       # pure noise: dense candidates
    for tag, img, cases in (("a", img_a, [(3.0, None), (5.0, 0.3), (7.5, None), (1.0, 0.5)]),
                            ("b", img_b, [(2.0, None), (4.3, 0.6), (12.0, None)])):
        cand = zo.peak_local_max_md1(img, None)
        assert len(np.unique(img[cand[:, 0], cand[:, 1]])) == len(cand)     # no intensity ties
        arrays[f"img_{tag}"] = img
        arrays[f"cand_{tag}"] = cand
        for i, (r, thr) in enumerate(cases):
            arrays[f"pts_{tag}{i}"] = lm.local_max(img, r, thr)
            arrays[f"arg_{tag}{i}"] = np.array([r, np.nan if thr is None else thr])
    save("peaks.npz", **arrays)


def golden_pca():
    """PCA of a feature matrix ("next" row f4): the reference's own ``pca`` (scikit-learn underneath) on the
    rotation-invariant features |Zc| of the three-fold test patches plus lattice patches."""
    from mtflearn.features import pca
    rng = np.random.default_rng(11)
    p3 = get_zps_test_patches(size=32, n_fold=3, num_patches=400)
    p4 = get_zps_test_patches(size=32, n_fold=4, num_patches=400)
    patches = np.concatenate([p3, p4]).astype(np.float32)
    patches += rng.normal(0, 0.01, patches.shape).astype(np.float32)
    feats = np.abs(ZPs(10, 32).fit_transform(patches).to_complex().data).astype(np.float32)
    save("pca.npz", feats=feats, pca2=pca(feats.astype(np.float64), 2), pca5=pca(feats.astype(np.float64), 5))


def golden_datasets():
    """Lattice coordinates and TMD frames ("next" row f1, second part) straight from the reference classes."""
    from mtflearn.datasets import TMDImageSimulator
    arrays = {}
    for tag, kw in (("a", dict(size=160, l=12, seed=3, angle=17.0, jitter=0.2)),
                    ("b", dict(size=200, l=10, seed=1, angle=-33.0, random_shift=False))):
        lat = HoneyCombLattice(**kw)
        img = lat.to_image(normalize=(tag == "b"))
        pa, pb = lat.get_points()
        arrays[f"lat_{tag}_A"], arrays[f"lat_{tag}_B"] = lat._coords_A, lat._coords_B
        arrays[f"lat_{tag}_ptsA"], arrays[f"lat_{tag}_ptsB"] = pa, pb
        arrays[f"lat_{tag}_img"] = img
    for tag, kw, basis in (("a", dict(size=(96, 128), a=20, theta=7.0), None),
                           ("b", dict(size=(80, 80), a=16, theta=30.0), [(0.0, 0.0, 'TM'), (1 / 3, 1 / 3, 'X'), (2 / 3, 2 / 3, 'X')])):
        sim = TMDImageSimulator(basis=basis, **kw)
        sim.add_random_vacancies('X', 3, 2, seed=0)
        sim.add_random_dopants('TM', 4, seed=1)
        img, masks = sim.simulate(return_masks=True)
        arrays[f"tmd_{tag}_img"] = img
        for label, m in masks.items():
            arrays[f"tmd_{tag}_mask_{label}"] = m
        arrays[f"tmd_{tag}_lbs"] = sim.lbs
        arrays[f"tmd_{tag}_pts"] = sim.pts
        arrays[f"tmd_{tag}_labels"] = np.array([str(v) for v in sim.labels])
    save("datasets.npz", **arrays)


def golden_clustering():
    """Cluster labels ("next" row f4): the reference's own kmeans_lbs / gmm_lbs (scikit-learn underneath,
    random_state=0) on rotation-invariant features of three patch families of unequal size; plus what the
    reference's rot_maps returns when it is called on COMPLEX moments (it squares the complex numbers)."""
    from mtflearn.clustering import kmeans_lbs, gmm_lbs
    rng = np.random.default_rng(21)
    fam = [get_zps_test_patches(size=32, n_fold=3, num_patches=500),
           get_zps_test_patches(size=32, n_fold=4, num_patches=300),
           get_zps_test_patches(size=32, n_fold=6, num_patches=200, include_center=False)]
    patches = np.concatenate(fam).astype(np.float32)
    patches += rng.normal(0, 0.02, patches.shape).astype(np.float32)
    perm = rng.permutation(len(patches))
    patches = patches[perm]
    truth = np.concatenate([np.full(500, 0), np.full(300, 1), np.full(200, 2)])[perm]
    o = ZPs(10, 32).fit_transform(patches)
    feats = np.abs(o.to_complex().data)
    oc = o.to_complex()
    save("clustering.npz", feats=feats.astype(np.float32), truth=truth,
         kmeans3=kmeans_lbs(feats, 3), kmeans5=kmeans_lbs(feats, 5), gmm3=gmm_lbs(feats, 3),
         rot_on_complex=oc.rot_maps([2, 3, 4, 6])[:32], rot_on_real=o.rot_maps([2, 3, 4, 6])[:32],
         cdata=oc.data[:32], cn=oc.n, cm=oc.m)


def golden_denoise():
    """Patch-SVD denoiser ("next" row f4): the reference's extract_patches / reconstruct_patches / denoise_svd on a
    noisy lattice frame.  denoise_svd uses random_state=None: the stored frame is ONE draw of its randomized SVD, the
    exact rank-r projection is stored beside it (the yardstick both implementations approximate)."""
    from mtflearn.denoise._denoise_svd import extract_patches, reconstruct_patches, denoise_svd
    rng = np.random.default_rng(5)
    lat = HoneyCombLattice(size=144, l=12, seed=2, angle=11.0)
    img = (lat.to_image() + rng.normal(0, 0.15, (144, 144))).astype(np.float64)[:128, :144]
    patches = extract_patches(img, 16, 5)
    rec = reconstruct_patches(patches * 1.0, img.shape, 5)
    clean, s = denoise_svd(img, 16, 6, extraction_step=5, verbose=False, return_s=True)
    flat = patches.reshape(patches.shape[0], -1)
    u, sv, vt = np.linalg.svd(flat, full_matrices=False)
    exact = reconstruct_patches(((u[:, :6] * sv[:6]) @ vt[:6]).reshape(-1, 16, 16), img.shape, 5)
    save("denoise.npz", img=img, patch_shape=np.array(patches.shape), patch_sha=np.array(sha(patches)), patch_head=patches[:3],
         rec=rec, clean=clean, s=s, exact=exact, s_exact=sv[:8])
    print("randomized vs exact: frame", np.abs(clean - exact).max(), "s", np.abs(s - sv[:6]).max() / sv[0])


if __name__ == "__main__":
    golden_index()
    golden_basis()
    golden_patches()
    golden_lattice()
    golden_render()
    golden_peaks()
    golden_pca()
    golden_datasets()
    golden_clustering()
    golden_denoise()
