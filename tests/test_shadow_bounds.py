"""CPU check of the error budget behind the FP32 screens of the thread-per-cell sweep kernel
(patchyparticles_b200/csrc/pp_sweep_cells.cu, pp_common.cuh: cell_shadow).

The kernel never takes a decision in reduced precision, but it DISMISSES pairs from a 16-bit fixed-point
shadow of the positions and FP32 arithmetic.  That is only safe while (a) the squared separation computed
from the shadows is within the screen band of the exact one and (b) the orientation screen has no false
negatives.  This file restates the shadow arithmetic in numpy (same quantisation, same operation order, FP32
rounding after every operation the kernel rounds) and measures both on random configurations.  The GPU parity
tests then show the kernel's decisions are the reference's; this test shows why, with the margin."""
import numpy as np

F = np.float32
SCREEN_BAND = 2e-3          # kScreenBand
CLAIMED_D2_ERROR = 2e-4


def fma32(a, b, c):
    """fmaf: exact product and sum (float64 holds the 48-bit product of two floats), one rounding to FP32."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def shadow(frac_cells):
    """cell_shadow: offset inside the cell in units of cell_w / 65535, rounded to nearest (16 bits)."""
    return np.rint(frac_cells * 65535.0)


def shadow_position(q, unit, off):
    """shadow_position: fmaf(2^23 + q, unit, off - 2^23 * unit) with the bias rounded to FP32 once."""
    bias = (off - 8388608.0 * unit.astype(np.float64)).astype(F)
    return fma32((8388608.0 + q).astype(F), unit, bias)


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
    worst = 0.0
    for cell_w in (1.119000002, 1.3, 1.678):  # smallest legal cell, a typical one, the largest (just below 1.5 r_cut)
        n = 400000
        w = np.full(3, cell_w)
        unit = (w / 65535.0).astype(F)
        # the moved particle in the home cell, a neighbour in one of the 27 cells around it, a displacement
        fp, fq = rng.random((n, 3)), rng.random((n, 3))
        code = rng.integers(-1, 2, size=(n, 3)).astype(np.float64)
        disp = (rng.random((n, 3)) - 0.5) * 0.3
        p_exact, q_exact = fp * w, (fq + code) * w
        keep = np.linalg.norm(q_exact - p_exact, axis=1) < 1.7  # what the screens are about
        fo = np.stack([shadow_position(shadow(fp[:, k]), np.full(n, unit[k]), np.zeros(n)) for k in range(3)], axis=1)
        qs = np.stack([shadow_position(shadow(fq[:, k]), np.full(n, unit[k]), code[:, k] * w[k]) for k in range(3)], axis=1)
        fn = (fo + disp.astype(F)).astype(F)
        a, b = (qs - fo).astype(F), (qs - fn).astype(F)
        d2o, d2n = d2_f32(a[:, 0], a[:, 1], a[:, 2]), d2_f32(b[:, 0], b[:, 1], b[:, 2])
        eo = np.abs(d2o.astype(np.float64) - ((q_exact - p_exact) ** 2).sum(axis=1))
        en = np.abs(d2n.astype(np.float64) - ((q_exact - p_exact - disp) ** 2).sum(axis=1))
        worst = max(worst, eo[keep].max(), en[keep].max())
    assert worst < CLAIMED_D2_ERROR, worst        # measured: ~1e-4 at the largest cell
    assert worst < 0.1 * SCREEN_BAND + 1e-4      # an order of magnitude inside the band


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
    d_f = (d + (rng.random((n, 3)) - 0.5) * 2 * 6e-5).astype(F)                           # worst-case shadow error per axis
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
