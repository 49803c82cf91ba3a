// pp_common.cuh -- device-side building blocks shared by all kernels of the engine.
//
// Exact FP64 arithmetic: the Kern-Frenkel pair test, the patch construction and the cell index are
// evaluated with explicitly rounded operations (__dadd_rn / __dmul_rn / __ddiv_rn / __dsqrt_rn) in
// the reference's operation order, so bond counts and accept/reject decisions are bit-identical to
// the CPU reference built without FMA contraction (MC.c:232-266, MC.c:225-227, cells.c:43-49).
// The library is also compiled with -fmad=false.
#pragma once

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace cg = cooperative_groups;

namespace pp {

constexpr int kMaxExp = 32;

// Per-launch constants (passed by value as a __grid_constant__ kernel parameter).
struct SimConst {
	double box[3];
	double shift[3];      // grid origin shift (0 for the reference-parity kernels)
	double cell_w[3];     // box / n_side
	double inv_box[3];    // 1 / box (host-rounded)
	double inv_w[3];      // n_side * inv_box
	int n_side[3];        // even cell grid (global)
	int n_cells;          // cells held by this engine: nxl * n_side[1] * n_side[2]
	int N;
	// x-slab decomposition: this engine holds the x layers x_lo .. x_lo + nxl - 1 (mod n_side[0]) of the global
	// grid; with slab != 0 the first and the last of them are halo layers owned by the neighbouring engines
	int nxl;              // layers held (n_side[0] when slab == 0)
	int x_lo;             // global index of local layer 0
	int slab;
	int own_lo, own_hi;   // local layers [own_lo, own_hi) are owned (moved, counted) by this engine
	int pad1;
	double sqr_rcut;      // (1 + delta)^2                               system.c:86
	double cosmax;        // KF_cosmax
	double cos2_lo, cos2_hi; // cosmax^2 (1 -+ 1e-11): guard band of the division-free patch test
	double inv_b2;        // 1 / base[2]^2
	int fast_patch;       // the division-free patch test is valid (cosmax > 0)
	int pad0;
	double base[12];      // 4 body-frame patch directions                system.c:20-31
	double exp_table[kMaxExp]; // exp(-k / T) from the host's libm        MC.c:332
	double disp_max, theta_max;
	uint32_t seed_lo, seed_hi;
	uint32_t sweep_lo, sweep_hi8; // sweep index low word, (high word << 8)
};

// Device-resident counters.
struct Counters {
	long long tries, accepted, confined, overlapped;
	unsigned long long w_tries, w_accepted; // fixed point, 2^-24
	long long dbonds;       // running change of the bond count from sweeps
	long long pair_checks;
	long long bonds;        // K3 result
	int overlap;            // K3 overlap flag
	int pad;
};

struct Vec3 {
	double x, y, z;
};

// x-slab decomposition: the thread that commits an accepted move of a particle in the engine's first (last) owned
// layer also stores the new state into the left (right) neighbour's halo copy of that layer -- plain stores into the
// neighbouring GPU's memory over NVLink (or into another engine's arrays on the same GPU).  Both sides hold the layer
// in the same order (cells in y, z order, ids ascending inside a cell), so the slot offset inside the layer is shared.
struct HaloMirror {
	double4 *pos_l;       // left neighbour's position array; nullptr: no mirror (single engine)
	double *ori_l;
	uint4 *posf_l;
	double4 *pos_r;       // right neighbour's arrays (its left halo starts at slot 0)
	double *ori_r;
	uint4 *posf_r;
	const int *off_l_ptr; // where the left neighbour published the first slot of its right halo; nullptr: off_l holds it
	int off_l;
	int ls1, ls2, ls3;    // this engine's first owned layer is [ls1, ls2), its last owned layer starts at ls3
	int pad;
};

// ---------------------------------------------------------------------------------------------
// exact helpers (no contraction)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dot3_rn(double ax, double ay, double az, double bx, double by, double bz) {
	// utils.h:15  ((a0*b0 + a1*b1) + a2*b2)
	return __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
}

// patch[i] = M . base[i]                                                   utils.h:17-21
__device__ __forceinline__ void patches_from(const double *base, const double *m, double *patches) {
#pragma unroll
	for(int i = 0; i < 4; i++) {
		const double bx = base[3 * i], by = base[3 * i + 1], bz = base[3 * i + 2];
		patches[3 * i] = dot3_rn(m[0], m[1], m[2], bx, by, bz);
		patches[3 * i + 1] = dot3_rn(m[3], m[4], m[5], bx, by, bz);
		patches[3 * i + 2] = dot3_rn(m[6], m[7], m[8], bx, by, bz);
	}
}

// Cell of a position on the engine's own (even, shifted) grid.  Any grid with cells of side >= r_cut
// finds the same pairs as the reference's (cells.c:43-49), so the index itself needs no parity with
// the reference; it only has to be identical in the cell build, the confinement test and the CPU
// restatement of the sweep (oracle/pp_oracle.c: grid_cell).
__device__ __forceinline__ int cell_coord(double r, double shift, double inv_box, int n_side) {
	const double t = __dmul_rn(__dsub_rn(r, shift), inv_box);
	const double f = __dsub_rn(t, floor(t));
	const int c = (int) __dmul_rn(f, (double) n_side);
	return c < n_side ? c : n_side - 1;
}

// Linear index in this engine's cell array; ix is the LOCAL layer.  -1: the position lies in a layer this
// engine does not hold (only possible with a slab decomposition).
__device__ __forceinline__ int cell_index(const SimConst &sc, double x, double y, double z, int &ix, int &iy, int &iz) {
	ix = cell_coord(x, sc.shift[0], sc.inv_box[0], sc.n_side[0]) - sc.x_lo;
	ix += (ix < 0) ? sc.n_side[0] : 0;
	iy = cell_coord(y, sc.shift[1], sc.inv_box[1], sc.n_side[1]);
	iz = cell_coord(z, sc.shift[2], sc.inv_box[2], sc.n_side[2]);
	if(ix >= sc.nxl) return -1;
	return (ix * sc.n_side[1] + iy) * sc.n_side[2] + iz;
}

// index of a local cell in the GLOBAL grid: the key of its random streams, identical on every decomposition
__device__ __forceinline__ int global_cell(const SimConst &sc, int ix, int iy, int iz) {
	int gx = ix + sc.x_lo;
	gx -= (gx >= sc.n_side[0]) ? sc.n_side[0] : 0;
	return (gx * sc.n_side[1] + iy) * sc.n_side[2] + iz;
}

// Fixed-point shadow of a position: every coordinate as a 32-bit fraction of the box side, frac(r / box) * 2^32, and
// the particle's slot in .w.  The number wraps with the periodic image, so the difference of two shadows read as a
// signed integer IS the minimum-image separation in units of box / 2^32 -- no cell offsets, no image search, and the
// same for every cell grid.  Only the conservative FP32 pre-screens of the sweep kernel and of the bond count read it
// (truncation + FP32 error of a squared separation < 1e-5 for boxes up to a few thousand diameters, screen bands
// 2e-3, tests/test_shadow_bounds.py); every decision is taken in FP64.
__device__ __forceinline__ unsigned box_fraction(double r, double inv_box) {
	const double t = __dmul_rn(r, inv_box);
	const double f = __dsub_rn(t, floor(t));                  // [0, 1]
	return __double2uint_rz(__dmul_rn(f, 4294967296.0));      // saturates at 2^32 - 1
}
__device__ __forceinline__ uint4 box_shadow(const SimConst &sc, double x, double y, double z, int slot) {
	return make_uint4(box_fraction(x, sc.inv_box[0]), box_fraction(y, sc.inv_box[1]), box_fraction(z, sc.inv_box[2]), (unsigned) slot);
}
// the coordinates of a shadow change, the slot in .w stays (a 8-byte and a 4-byte store: .w is never touched, so a
// copy of the layer held by another engine keeps its own slot numbers)
__device__ __forceinline__ void store_shadow_xyz(uint4 *dst, const uint4 v) {
	*reinterpret_cast<uint2 *>(dst) = make_uint2(v.x, v.y);
	reinterpret_cast<unsigned *>(dst)[2] = v.z;
}
// shadow separation -> FP32 separation: (float) (int) (q - p) * (box / 2^32)
__device__ __forceinline__ float shadow_sep(unsigned q, unsigned p, float unit) { return (float) (int) (q - p) * unit; }

// May one of the four patches of the particle with (FP32) matrix mf point at a neighbour at separation
// (x, y, z), |.|^2 = d2 ?  d . (M b_i) = (M^T d) . b_i with b_i = (+-1, +-1, +-1) * base_z (system.c:23-31);
// only the largest projection matters.  false means "certainly not" (see the band in the caller).
__device__ __forceinline__ bool may_point(const float *mf, float x, float y, float z, float d2, float point_thr) {
	const float ux = fmaf(mf[6], z, fmaf(mf[3], y, mf[0] * x));
	const float uy = fmaf(mf[7], z, fmaf(mf[4], y, mf[1] * x));
	const float uz = fmaf(mf[8], z, fmaf(mf[5], y, mf[2] * x));
	const float pm = fmaxf(fmaxf(uz - ux - uy, ux - uy - uz), fmaxf(ux + uy + uz, uy - ux - uz));
	return point_thr < 0.f || (pm > 0.f && pm * pm >= point_thr * d2);
}

// rint(d / box) as the reference computes it (MC.c:234-236), without the division in the common
// case: rint(d * inv_box) can only differ from rint(d / box) when the quotient is within a couple
// of ulps of a half-integer, and then the IEEE division is used.
__device__ __forceinline__ double image_count(double d, double box, double inv_box) {
	const double q = __dmul_rn(d, inv_box);
	const double k = rint(q);
	if(__builtin_expect(0.5 - fabs(__dsub_rn(q, k)) < 1e-7, 0)) return rint(__ddiv_rn(d, box));
	return k;
}

// Minimum-image separation and its square, bit-identical to MC.c:233-238.
__device__ __forceinline__ double min_image(const SimConst &sc, const Vec3 &p, double qx, double qy, double qz, Vec3 &d) {
	d.x = __dsub_rn(qx, p.x);
	d.y = __dsub_rn(qy, p.y);
	d.z = __dsub_rn(qz, p.z);
	d.x = __dsub_rn(d.x, __dmul_rn(sc.box[0], image_count(d.x, sc.box[0], sc.inv_box[0])));
	d.y = __dsub_rn(d.y, __dmul_rn(sc.box[1], image_count(d.y, sc.box[1], sc.inv_box[1])));
	d.z = __dsub_rn(d.z, __dmul_rn(sc.box[2], image_count(d.z, sc.box[2], sc.inv_box[2])));
	return dot3_rn(d.x, d.y, d.z, d.x, d.y, d.z);
}
__device__ __forceinline__ double min_image(const SimConst &sc, const Vec3 &p, const double4 &q, Vec3 &d) {
	return min_image(sc, p, q.x, q.y, q.z, d);
}

// Orientation test for a pair already known to be inside the bonding shell, exactly as MC.c:241-262
// (normalise by IEEE divisions, left-to-right dot products, strict comparisons).
// ori_q points at q's 3x3 matrix in global memory; it is only read if one of p's patches
// points at q (the reference only enters its inner loop in that case).
template<int S = 1>
static __device__ __noinline__ bool patches_bond_exact(const SimConst &sc, Vec3 d, double d2, const double *pp, const double *ori_q) {
	const double norm = __dsqrt_rn(d2);
	d.x = __ddiv_rn(d.x, norm);
	d.y = __ddiv_rn(d.y, norm);
	d.z = __ddiv_rn(d.z, norm);
	bool p_points = false;
#pragma unroll
	for(int i = 0; i < 4; i++) p_points |= (dot3_rn(d.x, d.y, d.z, pp[(3 * i) * S], pp[(3 * i + 1) * S], pp[(3 * i + 2) * S]) > sc.cosmax);
	if(!p_points) return false;
	double mq[9];
#pragma unroll
	for(int k = 0; k < 9; k++) mq[k] = ori_q[k];
	bool q_points = false;
#pragma unroll
	for(int j = 0; j < 4; j++) {
		const double bx = sc.base[3 * j], by = sc.base[3 * j + 1], bz = sc.base[3 * j + 2];
		const double qx = dot3_rn(mq[0], mq[1], mq[2], bx, by, bz);
		const double qy = dot3_rn(mq[3], mq[4], mq[5], bx, by, bz);
		const double qz = dot3_rn(mq[6], mq[7], mq[8], bx, by, bz);
		q_points |= (-dot3_rn(d.x, d.y, d.z, qx, qy, qz) > sc.cosmax);
	}
	return q_points;
}

// Same decision as patches_bond_exact without square root or division: for cosmax > 0,
//     (d . patch) / |d| > cosmax   <=>   d . patch > 0  and  (d . patch)^2 > cosmax^2 d^2.
// The squared cosine is compared through a relative guard band of 1e-11 (cos2_lo / cos2_hi, ten
// thousand times the rounding error of the shortcut); a value inside the band sends the pair to
// the exact routine, so the returned decision is always the reference's.
template<int S = 1>
__device__ __forceinline__ bool patches_bond(const SimConst &sc, const Vec3 &d, double d2, const double *pp, const double *ori_q) {
	if(!sc.fast_patch) return patches_bond_exact<S>(sc, d, d2, pp, ori_q);
	const double lo = __dmul_rn(sc.cos2_lo, d2), hi = __dmul_rn(sc.cos2_hi, d2);
	bool pass = false, border = false;
#pragma unroll
	for(int i = 0; i < 4; i++) {
		const double s = dot3_rn(d.x, d.y, d.z, pp[(3 * i) * S], pp[(3 * i + 1) * S], pp[(3 * i + 2) * S]);
		const double s2 = __dmul_rn(s, s);
		pass |= (s > 0. && s2 > hi);
		border |= (s > 0. && s2 >= lo && s2 <= hi);
	}
	if(border) return patches_bond_exact<S>(sc, d, d2, pp, ori_q);
	if(!pass) return false;
	// -n . (M_q b_j) = -(M_q^T n) . b_j ; the body-frame patches are (+-1, +-1, +-1) * base[2]  (system.c:23-31)
	const double vx = dot3_rn(ori_q[0], ori_q[3], ori_q[6], d.x, d.y, d.z);
	const double vy = dot3_rn(ori_q[1], ori_q[4], ori_q[7], d.x, d.y, d.z);
	const double vz = dot3_rn(ori_q[2], ori_q[5], ori_q[8], d.x, d.y, d.z);
	const double w0 = __dsub_rn(__dadd_rn(vx, vy), vz); // -(-1, -1, +1) . v
	const double w1 = __dsub_rn(__dadd_rn(vy, vz), vx); // -(+1, -1, -1) . v
	const double w2 = -__dadd_rn(__dadd_rn(vx, vy), vz); // -(+1, +1, +1) . v
	const double w3 = __dsub_rn(__dadd_rn(vx, vz), vy); // -(-1, +1, -1) . v
	const double wmax = fmax(fmax(w0, w1), fmax(w2, w3));
	if(!(wmax > 0.)) return false;
	const double lo_q = __dmul_rn(lo, sc.inv_b2), hi_q = __dmul_rn(hi, sc.inv_b2); // thresholds for w^2 (cos = base[2] w / |d|)
	const double t0 = __dmul_rn(w0, w0), t1 = __dmul_rn(w1, w1), t2 = __dmul_rn(w2, w2), t3 = __dmul_rn(w3, w3);
	const bool qpass = (w0 > 0. && t0 > hi_q) || (w1 > 0. && t1 > hi_q) || (w2 > 0. && t2 > hi_q) || (w3 > 0. && t3 > hi_q);
	const bool qborder = (w0 > 0. && t0 >= lo_q && t0 <= hi_q) || (w1 > 0. && t1 >= lo_q && t1 <= hi_q) || (w2 > 0. && t2 >= lo_q && t2 <= hi_q) ||
			(w3 > 0. && t3 >= lo_q && t3 <= hi_q);
	if(qborder) return patches_bond_exact<S>(sc, d, d2, pp, ori_q);
	return qpass;
}

// The same test when the moved particle's patches have not been formed: d . (M b_i) = (M^T d) . b_i,
// so its side of the test needs only the 3x3 matrix mp (registers).  Guarded like patches_bond; the
// exact routine rebuilds the patch vectors with the reference's arithmetic when it is needed.
__device__ __forceinline__ bool patches_bond_m(const SimConst &sc, const Vec3 &d, double d2, const double *mp, const double *ori_q) {
	if(sc.fast_patch) {
		const double lo = __dmul_rn(__dmul_rn(sc.cos2_lo, d2), sc.inv_b2), hi = __dmul_rn(__dmul_rn(sc.cos2_hi, d2), sc.inv_b2);
		const double ux = dot3_rn(mp[0], mp[3], mp[6], d.x, d.y, d.z);
		const double uy = dot3_rn(mp[1], mp[4], mp[7], d.x, d.y, d.z);
		const double uz = dot3_rn(mp[2], mp[5], mp[8], d.x, d.y, d.z);
		const double p0 = __dsub_rn(uz, __dadd_rn(ux, uy));  // (-1, -1, +1) . u
		const double p1 = __dsub_rn(ux, __dadd_rn(uy, uz));  // (+1, -1, -1) . u
		const double p2 = __dadd_rn(__dadd_rn(ux, uy), uz);  // (+1, +1, +1) . u
		const double p3 = __dsub_rn(uy, __dadd_rn(ux, uz));  // (-1, +1, -1) . u
		const double s0 = __dmul_rn(p0, p0), s1 = __dmul_rn(p1, p1), s2 = __dmul_rn(p2, p2), s3 = __dmul_rn(p3, p3);
		const bool ppass = (p0 > 0. && s0 > hi) || (p1 > 0. && s1 > hi) || (p2 > 0. && s2 > hi) || (p3 > 0. && s3 > hi);
		const bool pborder = (p0 > 0. && s0 >= lo && s0 <= hi) || (p1 > 0. && s1 >= lo && s1 <= hi) || (p2 > 0. && s2 >= lo && s2 <= hi) ||
				(p3 > 0. && s3 >= lo && s3 <= hi);
		if(!pborder) {
			if(!ppass) return false;
			const double vx = dot3_rn(ori_q[0], ori_q[3], ori_q[6], d.x, d.y, d.z);
			const double vy = dot3_rn(ori_q[1], ori_q[4], ori_q[7], d.x, d.y, d.z);
			const double vz = dot3_rn(ori_q[2], ori_q[5], ori_q[8], d.x, d.y, d.z);
			const double w0 = __dsub_rn(__dadd_rn(vx, vy), vz), w1 = __dsub_rn(__dadd_rn(vy, vz), vx);
			const double w2 = -__dadd_rn(__dadd_rn(vx, vy), vz), w3 = __dsub_rn(__dadd_rn(vx, vz), vy);
			const double t0 = __dmul_rn(w0, w0), t1 = __dmul_rn(w1, w1), t2 = __dmul_rn(w2, w2), t3 = __dmul_rn(w3, w3);
			const bool qpass = (w0 > 0. && t0 > hi) || (w1 > 0. && t1 > hi) || (w2 > 0. && t2 > hi) || (w3 > 0. && t3 > hi);
			const bool qborder = (w0 > 0. && t0 >= lo && t0 <= hi) || (w1 > 0. && t1 >= lo && t1 <= hi) || (w2 > 0. && t2 >= lo && t2 <= hi) ||
					(w3 > 0. && t3 >= lo && t3 <= hi);
			if(!qborder) return qpass;
		}
	}
	double pp[12];
	patches_from(sc.base, mp, pp);
	return patches_bond_exact<1>(sc, d, d2, pp, ori_q);
}

// Full pair test: -1 overlap, 0 none, 1 bond                                 MC.c:232-266
__device__ __forceinline__ int pair_test(const SimConst &sc, const Vec3 &p, const double *pp, const double4 &q, const double *__restrict__ ori_q) {
	Vec3 d;
	const double d2 = min_image(sc, p, q, d);
	if(d2 < 1.) return -1;
	if(d2 < sc.sqr_rcut) return patches_bond(sc, d, d2, pp, ori_q) ? 1 : 0;
	return 0;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
	for(int round = 0; round < 10; round++) {
		const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		const uint32_t n0 = hi1 ^ c1 ^ k0;
		const uint32_t n2 = hi0 ^ c3 ^ k1;
		c0 = n0;
		c1 = lo1;
		c2 = n2;
		c3 = lo0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u32_to_unit(uint32_t x) {
	return __dmul_rn(__dadd_rn((double) x, 0.5), 1.0 / 4294967296.0);
}

// ---------------------------------------------------------------------------------------------
// 27-cell neighbourhood of a home cell as a prefix-summed table in shared memory.
// An axis with only two cells has offsets {0, +1} (the -1 neighbour is the same cell).
// ---------------------------------------------------------------------------------------------
struct Neighbourhood {
	int first[27]; // first particle slot of neighbour cell l
	int incl[27];  // inclusive prefix of particle counts
};

template<int G>
__device__ __forceinline__ int build_neighbourhood(const cg::thread_block_tile<G> &tile, const SimConst &sc, const int *__restrict__ cell_start,
		int ix, int iy, int iz, Neighbourhood &nb) {
	int running = 0;
	for(int base = 0; base < 27; base += G) {
		const int l = base + (int) tile.thread_rank();
		int cnt = 0, first = 0;
		if(l < 27) {
			const int dx = l / 9 - 1, dy = (l / 3) % 3 - 1, dz = l % 3 - 1;
			bool dup = (!sc.slab && sc.nxl == 2 && dx < 0) || (sc.n_side[1] == 2 && dy < 0) || (sc.n_side[2] == 2 && dz < 0);
			int cx = ix + dx, cy = iy + dy, cz = iz + dz;
			if(sc.slab) dup |= (cx < 0 || cx >= sc.nxl); // beyond the halo: never needed by an owned cell
			if(!dup) {
				cx += (cx < 0) ? sc.nxl : 0;
				cx -= (cx >= sc.nxl) ? sc.nxl : 0;
				cy += (cy < 0) ? sc.n_side[1] : 0;
				cy -= (cy >= sc.n_side[1]) ? sc.n_side[1] : 0;
				cz += (cz < 0) ? sc.n_side[2] : 0;
				cz -= (cz >= sc.n_side[2]) ? sc.n_side[2] : 0;
				const int c = (cx * sc.n_side[1] + cy) * sc.n_side[2] + cz;
				first = cell_start[c];
				cnt = cell_start[c + 1] - first;
			}
		}
		int scan = cnt;
#pragma unroll
		for(int o = 1; o < G; o <<= 1) {
			const int v = tile.shfl_up(scan, o);
			if((int) tile.thread_rank() >= o) scan += v;
		}
		if(l < 27) {
			nb.first[l] = first;
			nb.incl[l] = running + scan;
		}
		running += tile.shfl(scan, G - 1);
	}
	tile.sync();
	return running;
}

// flat neighbour index j -> particle slot
__device__ __forceinline__ int neighbour_slot(const Neighbourhood &nb, int j) {
	int lo = 0, hi = 26; // first l with incl[l] > j
#pragma unroll
	for(int it = 0; it < 5; it++) {
		const int mid = (lo + hi) >> 1;
		if(nb.incl[mid] > j) hi = mid;
		else lo = mid + 1;
	}
	const int before = (lo > 0) ? nb.incl[lo - 1] : 0;
	return nb.first[lo] + (j - before);
}

__device__ __forceinline__ double4 ld_pos(const double4 *__restrict__ pos, int slot) {
	return pos[slot];
}

// the halo copy of an accepted move (see HaloMirror); off_l: first slot of the left neighbour's right halo
__device__ __forceinline__ void mirror_state(const HaloMirror &hm, int off_l, int sp, const Vec3 &p, const uint4 shadow, const double *m) {
	if(hm.pos_l == nullptr) return;
	double4 *tp;
	double *to;
	uint4 *tf;
	int d;
	if(sp < hm.ls2) {
		tp = hm.pos_l;
		to = hm.ori_l;
		tf = hm.posf_l;
		d = off_l + (sp - hm.ls1);
	}
	else if(sp >= hm.ls3) {
		tp = hm.pos_r;
		to = hm.ori_r;
		tf = hm.posf_r;
		d = sp - hm.ls3;
	}
	else return;
	// eight stores over NVLink instead of fourteen: 16-byte words wherever the alignment allows
	double *q = reinterpret_cast<double *>(tp + d); // x, y, z; the id in .w stays
	*reinterpret_cast<double2 *>(q) = make_double2(p.x, p.y);
	q[2] = p.z;
	tf[d] = make_uint4(shadow.x, shadow.y, shadow.z, (unsigned) d); // (.w: the slot on the neighbour's side, which is d)
	double *o = to + 9ll * d; // 72-byte records: 16-byte aligned for even d
	if((d & 1) == 0) {
#pragma unroll
		for(int k = 0; k < 8; k += 2) *reinterpret_cast<double2 *>(o + k) = make_double2(m[k], m[k + 1]);
		o[8] = m[8];
	}
	else {
		o[0] = m[0];
#pragma unroll
		for(int k = 1; k < 9; k += 2) *reinterpret_cast<double2 *>(o + k) = make_double2(m[k], m[k + 1]);
	}
}

__device__ __forceinline__ int mirror_offset(const HaloMirror &hm) {
	return (hm.pos_l != nullptr && hm.off_l_ptr != nullptr) ? __ldcg(hm.off_l_ptr) : hm.off_l;
}

} // namespace pp
