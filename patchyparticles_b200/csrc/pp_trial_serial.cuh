// pp_trial_serial.cuh -- one roto-translation trial executed by a single thread, entirely in FP64 and
// straight from global memory (MC.c:311-340 with the engine's confinement rule).  It is the
// fallback of both sweep kernels for cells whose neighbourhood does not fit their staging area
// (very dense packings, boxes with only two cells along an axis); same decisions as the staged
// paths, only slower.
#pragma once

#include "pp_kernels.cuh"

namespace pp {

struct Tally {
	unsigned int tries, acc, conf, ovl;
	int dbonds;
	unsigned long long wt, wa, checks;
};

__device__ __forceinline__ int wrap_cell(int c, int n) {
	c += (c < 0) ? n : 0;
	c -= (c >= n) ? n : 0;
	return c;
}

static __device__ __noinline__ void trial_serial(const SimConst &sc, const int *__restrict__ cell_start, double4 *pos, double *ori, uint4 *posf,
		const Proposal *pr, int cell, int ix, int iy, int iz, int sp, Tally &tl, const HaloMirror &hm, int hm_off) {
	const int ny = sc.n_side[1], nz = sc.n_side[2];
	const double4 p4 = pos[sp];
	const Vec3 p_old = {p4.x, p4.y, p4.z};
	const Vec3 p_new = {__dadd_rn(p4.x, pr->dx), __dadd_rn(p4.y, pr->dy), __dadd_rn(p4.z, pr->dz)};
	int cx, cy, cz;
	if(cell_index(sc, p_new.x, p_new.y, p_new.z, cx, cy, cz) != cell) { // would leave its cell
		tl.conf++;
		return;
	}
	const unsigned long long Wfix = (unsigned long long) pr->wfix;
	tl.wt += Wfix;
	double m[9], mn[9];
	for(int k = 0; k < 9; k++) m[k] = ori[9ll * sp + k];
	for(int i = 0; i < 3; i++)
		for(int j = 0; j < 3; j++) mn[3 * i + j] = dot3_rn(m[3 * i], m[3 * i + 1], m[3 * i + 2], pr->R[j], pr->R[3 + j], pr->R[6 + j]);
	int b_old = 0, b_new = 0, n_seen = 0;
	bool ovl = false;
	for(int ddx = -1; ddx <= 1; ddx++) {
		int ccx = ix + ddx;
		if(sc.slab) {
			if(ccx < 0 || ccx >= sc.nxl) continue; // beyond the halo
		}
		else {
			if(sc.nxl == 2 && ddx < 0) continue; // two cells along x: -1 and +1 are the same neighbour
			ccx = wrap_cell(ccx, sc.nxl);
		}
		for(int ddy = -1; ddy <= 1; ddy++) {
			if(ny == 2 && ddy < 0) continue;
			for(int ddz = -1; ddz <= 1; ddz++) {
				if(nz == 2 && ddz < 0) continue;
				const int c = (ccx * ny + wrap_cell(iy + ddy, ny)) * nz + wrap_cell(iz + ddz, nz);
				for(int s = cell_start[c]; s < cell_start[c + 1]; s++) {
					if(s == sp) continue;
					n_seen++;
					const double4 q4 = pos[s];
					Vec3 d;
					double d2 = min_image(sc, p_old, q4, d);
					if(d2 >= 1. && d2 < sc.sqr_rcut) b_old += patches_bond_m(sc, d, d2, m, ori + 9ll * s) ? 1 : 0;
					d2 = min_image(sc, p_new, q4, d);
					if(d2 < 1.) ovl = true;
					else if(d2 < sc.sqr_rcut) b_new += patches_bond_m(sc, d, d2, mn, ori + 9ll * s) ? 1 : 0;
				}
			}
		}
	}
	tl.checks += 2ull * (unsigned long long) n_seen;
	if(ovl) {
		tl.ovl++;
		return;
	}
	const int dE = b_old - b_new;
	if(dE < 0 || pr->u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
		double *p = reinterpret_cast<double *>(pos + sp);
		p[0] = p_new.x;
		p[1] = p_new.y;
		p[2] = p_new.z;
		const uint4 moved = box_shadow(sc, p_new.x, p_new.y, p_new.z, sp);
		store_shadow_xyz(posf + sp, moved);
		for(int k = 0; k < 9; k++) ori[9ll * sp + k] = mn[k];
		mirror_state(hm, hm_off, sp, p_new, moved, mn);
		tl.dbonds += b_new - b_old;
		tl.acc++;
		tl.wa += Wfix;
	}
}

// The same trial with the G lanes of a group sharing the 27 cells (cell c goes to lane c mod G): what a crowded cell whose
// neighbourhood exceeds the staging area runs, instead of leaving the whole walk to one lane.  Every lane calls it; the owner
// lane keeps the books and commits, the group leaves together.
template<int G>
static __device__ __noinline__ void trial_group(const SimConst &sc, const int *__restrict__ cell_start, double4 *pos, double *ori, uint4 *posf,
		const Proposal *pr, int cell, int ix, int iy, int iz, int sp, Tally &tl, const HaloMirror &hm, int hm_off, int gl, unsigned gmask, bool is_owner) {
	const int ny = sc.n_side[1], nz = sc.n_side[2];
	const double4 p4 = pos[sp];
	const Vec3 p_old = {p4.x, p4.y, p4.z};
	const Vec3 p_new = {__dadd_rn(p4.x, pr->dx), __dadd_rn(p4.y, pr->dy), __dadd_rn(p4.z, pr->dz)};
	int cx, cy, cz;
	if(cell_index(sc, p_new.x, p_new.y, p_new.z, cx, cy, cz) != cell) { // would leave its cell
		if(is_owner) tl.conf++;
		return;
	}
	const unsigned long long Wfix = (unsigned long long) pr->wfix;
	if(is_owner) tl.wt += Wfix;
	double m[9], mn[9];
	for(int k = 0; k < 9; k++) m[k] = ori[9ll * sp + k];
	for(int i = 0; i < 3; i++)
		for(int j = 0; j < 3; j++) mn[3 * i + j] = dot3_rn(m[3 * i], m[3 * i + 1], m[3 * i + 2], pr->R[j], pr->R[3 + j], pr->R[6 + j]);
	int b_old = 0, b_new = 0, n_seen = 0;
	bool ovl = false;
	for(int c27 = gl; c27 < 27; c27 += G) {
		const int ddx = c27 / 9 - 1, ddy = (c27 / 3) % 3 - 1, ddz = c27 % 3 - 1;
		int ccx = ix + ddx;
		if(sc.slab) {
			if(ccx < 0 || ccx >= sc.nxl) continue; // beyond the halo
		}
		else {
			if(sc.nxl == 2 && ddx < 0) continue; // two cells along x: -1 and +1 are the same neighbour
			ccx = wrap_cell(ccx, sc.nxl);
		}
		if((ny == 2 && ddy < 0) || (nz == 2 && ddz < 0)) continue;
		const int c = (ccx * ny + wrap_cell(iy + ddy, ny)) * nz + wrap_cell(iz + ddz, nz);
		for(int s = cell_start[c]; s < cell_start[c + 1]; s++) {
			if(s == sp) continue;
			n_seen++;
			const double4 q4 = pos[s];
			Vec3 d;
			double d2 = min_image(sc, p_old, q4, d);
			if(d2 >= 1. && d2 < sc.sqr_rcut) b_old += patches_bond_m(sc, d, d2, m, ori + 9ll * s) ? 1 : 0;
			d2 = min_image(sc, p_new, q4, d);
			if(d2 < 1.) ovl = true;
			else if(d2 < sc.sqr_rcut) b_new += patches_bond_m(sc, d, d2, mn, ori + 9ll * s) ? 1 : 0;
		}
	}
	unsigned packed = (unsigned) b_old | ((unsigned) b_new << 8) | ((unsigned) n_seen << 16) | (ovl ? 1u << 31 : 0u); // (n_seen < 2^15)
#pragma unroll
	for(int o = G / 2; o > 0; o >>= 1) {
		const unsigned other = __shfl_xor_sync(gmask, packed, o);
		packed = ((packed & 0x7fffffffu) + (other & 0x7fffffffu)) | ((packed | other) & 0x80000000u);
	}
	b_old = (int) (packed & 0xffu);
	b_new = (int) ((packed >> 8) & 0xffu);
	ovl = (packed >> 31) != 0u;
	if(is_owner) {
		tl.checks += 2ull * (unsigned long long) ((packed >> 16) & 0x7fffu);
		if(ovl) tl.ovl++;
		else {
			const int dE = b_old - b_new;
			if(dE < 0 || pr->u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
				double *p = reinterpret_cast<double *>(pos + sp);
				p[0] = p_new.x;
				p[1] = p_new.y;
				p[2] = p_new.z;
				const uint4 moved = box_shadow(sc, p_new.x, p_new.y, p_new.z, sp);
				store_shadow_xyz(posf + sp, moved);
				for(int k = 0; k < 9; k++) ori[9ll * sp + k] = mn[k];
				mirror_state(hm, hm_off, sp, p_new, moved, mn);
				tl.dbonds += b_new - b_old;
				tl.acc++;
				tl.wa += Wfix;
			}
		}
	}
	__syncwarp(gmask); // the next trial of the cell reads what the owner has just written
}

} // namespace pp