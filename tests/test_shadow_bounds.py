"""CPU check of the error budget behind the FP32 screens of the sweep kernel and of the bond count
(patchyparticles_b200/csrc/pp_sweep_cells.cu, pp_bonds.cu, pp_common.cuh: box_shadow / shadow_sep).

The kernels never take a decision in reduced precision, but they DISMISS pairs from a fixed-point shadow of
the positions (every coordinate as a 32-bit fraction of the box side) and FP32 arithmetic.  That is only safe
while (a) the squared separation computed from the shadows is within the screen band of the exact one and
(b) the orientation screen has no false negatives.  This file restates the shadow arithmetic in numpy (same
quantisation, same operation order, FP32 rounding after every operation the kernel rounds) and measures both on
random configurations.  The GPU parity tests then show the kernel's decisions are the reference's; this test
shows why, with the margin."""
import numpy as np

F = np.float32
SCREEN_BAND = 2e-3          # kScreenBand
CLAIMED_D2_ERROR = 1e-5     # DESIGN.md section 3 (boxes up to a few thousand diameters)


def fma32(a, b, c):
    """fmaf: exact product and sum (float64 holds the 48-bit product of two floats), one rounding to FP32."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def box_fraction(r, box):
    """box_fraction: frac(r * (1 / box)) * 2^32 truncated to a 32-bit unsigned integer (saturating)."""
    t = r * (1.0 / box)
    f = t - np.floor(t)
    return np.minimum(np.floor(f * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)


def shadow_sep(q, p, unit):
    """shadow_sep: (float) (int) (q - p) * unit -- the wrapped difference IS the minimum image."""
    return (((q - p).astype(np.uint32)).view(np.int32).astype(F) * unit).astype(F)


def d2_f32(ax, ay, az):
    return fma32(az, az, fma32(ay, ay, (ax * ax).astype(F)))


def random_rotations(n, rng):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(n, 3, 3)


def test_shadow_distance_error_within_band():
    rng = np.random.default_rng(1)
    for box_side, claimed in ((4.5, CLAIMED_D2_ERROR), (22.0, CLAIMED_D2_ERROR), (191.2, CLAIMED_D2_ERROR), (3000.0, CLAIMED_D2_ERROR), (40000.0, 2e-4)):
        n = 400000
        box = np.array([box_side, box_side * 1.1, box_side * 0.93])
        unit = (box / 4294967296.0).astype(F)
        # the moved particle anywhere in the box (unwrapped coordinates, a few images away), a neighbour within range of
        # it through the periodic boundary, a displacement
        p_exact = (rng.random((n, 3)) * 5 - 2) * box
        sep = rng.normal(size=(n, 3))
        sep *= (rng.random(n) * 1.7 / np.linalg.norm(sep, axis=1))[:, None]
        q_exact = p_exact + sep + rng.integers(-2, 3, size=(n, 3)) * box
        disp = (rng.random((n, 3)) - 0.5) * 0.3
        ps = np.stack([box_fraction(p_exact[:, k], box[k]) for k in range(3)], axis=1)
        qs = np.stack([box_fraction(q_exact[:, k], box[k]) for k in range(3)], axis=1)
        a = np.stack([shadow_sep(qs[:, k], ps[:, k], unit[k]) for k in range(3)], axis=1)
        b = (a - disp.astype(F)).astype(F)
        d2o, d2n = d2_f32(a[:, 0], a[:, 1], a[:, 2]), d2_f32(b[:, 0], b[:, 1], b[:, 2])
        eo = np.abs(d2o.astype(np.float64) - (sep ** 2).sum(axis=1))
        en = np.abs(d2n.astype(np.float64) - ((sep - disp) ** 2).sum(axis=1))
        worst = max(eo.max(), en.max())
        assert worst < claimed, (box_side, worst)      # measured: ~1e-6 at 191, ~3e-6 at 3000, ~4e-5 at 40000
        assert worst < 0.1 * SCREEN_BAND               # an order of magnitude inside the band


def test_orientation_screen_has_no_false_negatives():
    """may_point() must say 'maybe' whenever the exact test (MC.c:249-255) finds a patch of the moved particle
    pointing at the neighbour; how often it says 'maybe' otherwise is only a matter of speed."""
    rng = np.random.default_rng(2)
    cosmax, delta = 0.98, 0.119
    n = 2000000
    m = random_rotations(n, rng)
    base = np.array([[-1, -1, 1], [1, -1, -1], [1, 1, 1], [-1, 1, -1]]) / np.sqrt(3.0)   # system.c:23-31
    patches = np.einsum("nij,kj->nki", m, base)                                          # patch = M . b
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    dist = 1.0 + rng.random(n) * delta
    d = dirs * dist[:, None]
    cos = np.einsum("nki,ni->nk", patches, dirs).max(axis=1)
    points = cos > cosmax                                                                  # exact, FP64
    # the screen: separation perturbed by the shadow error, FP32 matrix, FP32 arithmetic, 2e-3 relative band
    d_f = (d + (rng.random((n, 3)) - 0.5) * 2 * 6e-5).astype(F)                           # far above the shadow error per axis
    mf = m.astype(F)
    d2 = d2_f32(d_f[:, 0], d_f[:, 1], d_f[:, 2])
    u = [fma32(mf[:, 2, k], d_f[:, 2], fma32(mf[:, 1, k], d_f[:, 1], (mf[:, 0, k] * d_f[:, 0]).astype(F))) for k in range(3)]
    ux, uy, uz = u
    pm = np.maximum(np.maximum(uz - ux - uy, ux - uy - uz), np.maximum(ux + uy + uz, uy - ux - uz)).astype(F)
    thr = F(np.float64(cosmax) ** 2 * (1 - 1e-11) * 3.0) * F(1 - SCREEN_BAND)
    maybe = (pm > 0) & ((pm * pm).astype(F) >= (thr * d2).astype(F))
    assert points.sum() > 10000                      # ~4 % of the directions
    assert not np.any(points & ~maybe)               # no false negative
    assert maybe.mean() < 1.2 * points.mean() + 1e-3  # and the band costs almost nothing
    # the closest call: how far inside the band the worst pointing pair sits
    margin = ((pm * pm).astype(np.float64) / (thr.astype(np.float64) * d2.astype(np.float64)))[points].min()
    assert margin > 1.0015
