"""CPU check of the arithmetic behind the mirror-folded projection kernel (csrc/zb200_project_fold.cu), against the
oracle: the mirror parities of every Zernike plane, the butterfly to the four class inputs, the quarter-window
contraction, and the fp16 operand split with a power-of-two input scale.  No GPU, no library: numpy only -- what is
emulated here is exactly what the kernel's splitter warps and MMAs compute, except for the tensor core's truncating
fp32 accumulation (that part is measured on the device: profiles/r02_fold_chunk_probe.log)."""
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import zernike_oracle as zo  # noqa: E402


def _parities(v):
    """(sj, si) of a plane: its sign under the column mirror j -> k-1-j and under the row mirror i -> k-1-i."""
    sj = int(np.sign((v * v[:, ::-1]).sum()))
    si = int(np.sign((v * v[::-1, :]).sum()))
    return sj, si


@pytest.mark.parametrize("n_max,k", [(12, 64), (20, 64), (6, 128)])
def test_every_plane_has_the_parity_the_kernel_assumes(n_max, k):
    """cos(m theta): even under the row mirror, (-1)^m under the column mirror; sin(m theta): odd, (-1)^(m+1)
    (_zps.py:68-75: x = column, y = row, theta = arctan2(y, x)); symmetric to rounding, so the fold is exact."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(n_max, k)
    for j in range(len(n)):
        cos, odd = m[j] >= 0, bool(m[j] % 2)
        want = (-1 if odd else 1, 1) if cos else (1 if odd else -1, -1)
        assert _parities(v[j]) == want, (n[j], m[j])
        sj, si = want
        scale = np.abs(v[j]).max()
        assert np.abs(v[j] - sj * v[j][:, ::-1]).max() <= 1e-7 * scale       # the reference's own power sum is the limit
        assert np.abs(v[j] - si * v[j][::-1, :]).max() <= 1e-7 * scale


def _split16(f):
    """x1 = the top 11 significand bits of fp32 f (exact in fp16), x2 = RN_f16(f - x1): the splitter's split_pair."""
    t = (f.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    x1 = t.astype(np.float16)
    assert np.array_equal(x1.astype(np.float32), t)
    return x1.astype(np.float64), (f - t).astype(np.float16).astype(np.float64)


@pytest.mark.parametrize("n_max,k,scale", [(12, 64, 1.0), (20, 64, 3.0e4), (5, 128, 2.0e-6)])
def test_butterfly_quarter_window_contraction_equals_the_oracle(n_max, k, scale):
    rng = np.random.default_rng(n_max + k)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(n_max, k)
    x = ((rng.random((96, k, k)) * 0.8 + 0.1) * scale).astype(np.float32)
    ref = zo.project_patches(x.astype(np.float64), v)
    h = k // 2
    # the four mirror images of the upper-left quadrant (what the four TMA boxes of a super-block hold)
    a, b = x[:, :h, :h], x[:, :h, ::-1][:, :, :h]
    c, d = x[:, ::-1, :][:, :h, :h], x[:, ::-1, ::-1][:, :h, :h]
    # power-of-two input scale from a bound of |x| (auto_shift: |x| 2^sh <= 2^11), butterfly in fp32
    e = int(np.floor(np.log2(float(np.abs(x).max())))) + 1
    sc = np.float32(2.0 ** (11 - e))
    p, q, r, s = (a + b) * sc, (a - b) * sc, (c + d) * sc, (c - d) * sc
    f = {(1, 1): p + r, (-1, -1): q - s, (-1, 1): q + s, (1, -1): p - r}      # A_re, A_im, B_re, B_im of the kernel
    assert max(float(np.abs(t).max()) for t in f.values()) < 2.0 ** 14          # four-term fold stays in fp16 range
    area = np.pi * k * k / 4
    out = np.empty_like(ref)
    for j in range(len(n)):
        sj, si = _parities(v[j])
        vq = (v[j][:h, :h] + sj * v[j][:h, ::-1][:, :h] + si * v[j][::-1, :][:h, :h] + sj * si * v[j][::-1, ::-1][:h, :h]) / 4
        b1 = vq.astype(np.float16).astype(np.float64)
        b2 = (vq - b1).astype(np.float16).astype(np.float64)
        x1, x2 = _split16(f[(sj, si)])
        acc = np.einsum("nij,ij->n", x1, b1) + np.einsum("nij,ij->n", x2, b1) + np.einsum("nij,ij->n", x1, b2)
        out[:, j] = acc / (area * float(sc))
    err = np.abs(out - ref)
    gate = 1e-4 * np.abs(ref) + 1e-6 * np.abs(ref).max()
    assert (err / gate).max() < 0.1, (err / gate).max()      # 0.02 measured: the split itself is 50x inside the gate
