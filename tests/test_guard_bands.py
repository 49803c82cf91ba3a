"""CPU check of the two guarded FP64 shortcuts of the CUDA path (patchyparticles_b200/csrc/pp_common.cuh).

image_count:   rint(d / L) (MC.c:234-236) is taken from rint(d * (1 / L)) unless the product is within 1e-7 of a
               half-integer, where the IEEE division is used.
patches_bond:  (d . p) / |d| > cosmax (MC.c:249-261, after normalising d by IEEE divisions) is decided as
               d . p > 0 and (d . p)^2 > cosmax^2 d^2 unless the squared cosine is within a relative 1e-11 of the
               threshold, where the reference's arithmetic is used.
Both claims -- outside the band the shortcut and the reference agree -- are checked here in numpy FP64 (IEEE, same
operation order), on random inputs and on inputs crowded around the thresholds."""
import numpy as np


def test_image_count_shortcut_agrees_outside_its_band():
    rng = np.random.default_rng(3)
    for L in (9.39768761, 15.874010519682, 50.7968, 191.20660252, 3.3570000001):
        inv = 1.0 / L
        k = rng.integers(-3, 4, size=2000000).astype(np.float64)
        # separations spread over several boxes, and a second batch crowded around the half-integer quotients
        d_wide = (rng.random(2000000) - 0.5) * 7 * L
        d_edge = (k + 0.5) * L + (rng.random(2000000) - 0.5) * L * 1e-5 * rng.random(2000000) ** 4
        for d in (d_wide, d_edge):
            q = d * inv
            fast = np.rint(q)
            guarded = 0.5 - np.abs(q - fast) < 1e-7      # these go to the division
            exact = np.rint(d / L)
            assert np.array_equal(fast[~guarded], exact[~guarded])
        assert guarded.sum() > 1000 and (~guarded).sum() > 1000  # the crowded batch exercises both sides of the band


def test_squared_cosine_shortcut_agrees_outside_its_band():
    rng = np.random.default_rng(4)
    cosmax = 0.98
    c2 = cosmax * cosmax
    lo, hi = c2 * (1 - 1e-11), c2 * (1 + 1e-11)
    n = 3000000
    p = rng.normal(size=(n, 3))
    p /= np.linalg.norm(p, axis=1)[:, None]                      # a patch: unit vector up to rounding
    # separations with the cosine crowded around cosmax (both sides, down to 1e-14) and anywhere else
    t = rng.normal(size=(n, 3))
    t -= (t * p).sum(axis=1)[:, None] * p
    t /= np.linalg.norm(t, axis=1)[:, None]
    cos_target = np.where(rng.random(n) < 0.5, cosmax + (rng.random(n) - 0.5) * 10.0 ** rng.uniform(-14, -2, n), rng.uniform(-1, 1, n))
    cos_target = np.clip(cos_target, -1, 1)
    dist = 1.0 + rng.random(n) * 0.119
    d = (cos_target[:, None] * p + np.sqrt(1 - cos_target ** 2)[:, None] * t) * dist[:, None]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]           # utils.h:15 order
    # reference: normalise by divisions, then the dot product
    norm = np.sqrt(d2)
    nx, ny, nz = d[:, 0] / norm, d[:, 1] / norm, d[:, 2] / norm
    ref = ((nx * p[:, 0] + ny * p[:, 1]) + nz * p[:, 2]) > cosmax
    # shortcut
    s = (d[:, 0] * p[:, 0] + d[:, 1] * p[:, 1]) + d[:, 2] * p[:, 2]
    s2 = s * s
    fast = (s > 0) & (s2 > hi * d2)
    border = (s > 0) & (s2 >= lo * d2) & (s2 <= hi * d2)      # these go to the reference's arithmetic
    assert np.array_equal(fast[~border], ref[~border])
    assert border.sum() > 100 and ref.sum() > 100000
