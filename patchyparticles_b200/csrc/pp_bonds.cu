// pp_bonds.cu -- K3 (full bond count), per-particle bond counts, K4 (rescale), K5 (proposal
// evaluator), K6 (re-orthonormalise) and the upload/download packers.
#include "pp_kernels.cuh"

namespace pp {

constexpr int kBondThreads = 128;

// ---------------------------------------------------------------------------------------------
// K3: number of bonded pairs + overlap flag.  Replaces the total-energy loops of main.c:55-61,
// MC.c:488-495, MC.c:556-560 and MC.c:769-776 (0.5 * sum_i MC_energy(i)).  One tile of G lanes
// per cell; every unordered pair is tested once (slot_q > slot_p), which is exact because the
// reference's pair test is antisymmetric in (p, q) operation by operation.  Lane-local integer
// counts are reduced per block in shared memory, then one atomic per block.
// ---------------------------------------------------------------------------------------------
template<int G>
__global__ void __launch_bounds__(kBondThreads) k_total_bonds(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start,
		const double4 *__restrict__ pos, const double *__restrict__ ori, Counters *__restrict__ counters) {
	constexpr int kTiles = kBondThreads / G;
	__shared__ Neighbourhood s_nb[kTiles];
	__shared__ int s_bonds, s_overlap;
	cg::thread_block block = cg::this_thread_block();
	cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
	const int tile_in_block = threadIdx.x / G;
	const int lane = tile.thread_rank();
	if(threadIdx.x == 0) {
		s_bonds = 0;
		s_overlap = 0;
	}
	__syncthreads();

	int bonds = 0, overlap = 0;
	const int nyz = sc.n_side[1] * sc.n_side[2];
	for(int cell = blockIdx.x * kTiles + tile_in_block; cell < sc.n_cells; cell += gridDim.x * kTiles) {
		const int first = cell_start[cell], last = cell_start[cell + 1];
		const int ix = cell / nyz, iy = (cell - ix * nyz) / sc.n_side[2], iz = cell - ix * nyz - iy * sc.n_side[2];
		// only owned layers; a pair across a slab boundary is counted by the engine on its left, because the
		// right halo comes after every owned slot and the left halo before (slot_q > slot_p below)
		if(first == last || ix < sc.own_lo || ix >= sc.own_hi) continue;
		Neighbourhood &nb = s_nb[tile_in_block];
		tile.sync();
		const int total = build_neighbourhood<G>(tile, sc, cell_start, ix, iy, iz, nb);
		for(int sp = first; sp < last; sp++) {
			const double4 p4 = pos[sp];
			const Vec3 p = {p4.x, p4.y, p4.z};
			double mp[9], pp[12];
#pragma unroll
			for(int k = 0; k < 9; k++) mp[k] = ori[9ll * sp + k];
			patches_from(sc.base, mp, pp);
			for(int j = lane; j < total; j += G) {
				const int sq = neighbour_slot(nb, j);
				if(sq <= sp) continue;
				const int v = pair_test(sc, p, pp, pos[sq], ori + 9ll * sq);
				bonds += (v == 1);
				overlap |= (v == -1);
			}
		}
	}
	bonds = cg::reduce(tile, bonds, cg::plus<int>());
	overlap = cg::reduce(tile, overlap, cg::bit_or<int>());
	if(lane == 0) {
		if(bonds) atomicAdd(&s_bonds, bonds);
		if(overlap) atomicOr(&s_overlap, 1);
	}
	__syncthreads();
	if(threadIdx.x == 0) {
		if(s_bonds) atomicAdd((unsigned long long *) &counters->bonds, (unsigned long long) s_bonds);
		if(s_overlap) atomicOr(&counters->overlap, 1);
	}
}

__global__ void k_zero_bonds(Counters *counters) {
	counters->bonds = 0;
	counters->overlap = 0;
}

// ---------------------------------------------------------------------------------------------
// K3, streaming version (the default): one thread per (particle, column of the half shell).  Every unordered pair
// is met once: the particle's own cell (slot_q > slot_p) and the 13 neighbour cells with (dx, dy, dz) > (0, 0, 0)
// in lexicographic order, i.e. five columns -- (0, 0): the rest of the cell and the cell above it in z;
// (0, +1), (+1, -1), (+1, 0), (+1, +1): three cells each.  Every candidate is first screened in FP32 from the
// fixed-point shadows (box fractions; the same conservative screen as the sweep: squared distance below r_cut^2 + 2e-3,
// shadow arithmetic good to 1e-5), so the FP64 position of a particle is fetched only for pairs that may be in
// range and an orientation matrix only for a pair inside the bonding shell (p's side first; q's matrix only if
// one of p's patches points at q, as the reference enters its inner loop, MC.c:249-255).  The global overlap
// flag is polled by every thread: after the first overlap the rest returns at once, like the reference's
// `while(!overlap_found)` loops (MC.c:488-495, 556-560).  Needs at least four cells per axis (the half shell
// would meet a cell twice on an axis with two); smaller grids use the tile kernel above.
// A slab engine counts the pairs of its owned layers among themselves and with the right halo (dx = +1), never
// with the left halo (dx = -1 is not in the half shell): every cross-boundary pair is counted once, on its left.
// ---------------------------------------------------------------------------------------------
constexpr int kBondCellThreads = 256;

// FP64 pair test of (sp, sq) in the reference's operation order (MC.c:233-262) -> bond count, overlap
static __device__ __noinline__ void bond_pair_exact(const SimConst &sc, const double4 *__restrict__ pos, const double *__restrict__ ori, int sp, int sq,
		int &bonds, int &overlap) {
	const double4 p4 = pos[sp];
	const Vec3 p = {p4.x, p4.y, p4.z};
	Vec3 d;
	const double d2 = min_image(sc, p, pos[sq], d);
	if(d2 < 1.) overlap = 1;
	else if(d2 < sc.sqr_rcut) {
		double m[9];
#pragma unroll
		for(int k = 0; k < 9; k++) m[k] = ori[9ll * sp + k];
		bonds += patches_bond_m(sc, d, d2, m, ori + 9ll * sq) ? 1 : 0;
	}
}

__global__ void __launch_bounds__(kBondCellThreads) k_total_bonds_cells(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start,
		const int *__restrict__ cell_sorted, const double4 *__restrict__ pos, const double *__restrict__ ori, const uint4 *__restrict__ posf,
		Counters *counters) {
	// pairs that pass the FP32 screen are queued per warp and tested in FP64 by all its lanes together: the exact test is
	// a few hundred instructions, and taken where it is found it would run with one or two lanes of the warp active
	constexpr int kQueue = 96;
	__shared__ int2 s_queue[kBondCellThreads / 32][kQueue];
	__shared__ int s_queued[kBondCellThreads / 32];
	__shared__ int s_bonds, s_overlap;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if(threadIdx.x == 0) {
		s_bonds = 0;
		s_overlap = 0;
	}
	if(lane == 0) s_queued[warp] = 0;
	__syncthreads();
	const int ny = sc.n_side[1], nz = sc.n_side[2], nyz = ny * nz;
	int bonds = 0, overlap = 0;
	const long long t = (long long) blockIdx.x * kBondCellThreads + threadIdx.x;
	const int sp = (int) (t / 5), col = (int) (t - 5ll * sp);
	if(sp < sc.N && *((volatile int *) &counters->overlap) == 0) {
		const int cell = cell_sorted[sp];
		const int ix = cell / nyz, iy = (cell - ix * nyz) / nz, iz = cell - ix * nyz - iy * nz;
		if(cell < sc.n_cells && ix >= sc.own_lo && ix < sc.own_hi) {
			const int dx = col >= 2, dy = col == 0 ? 0 : (col == 1 ? 1 : col - 3);
			int cx = ix + dx;
			if(!sc.slab) cx -= (cx >= sc.nxl) ? sc.nxl : 0; // (slab: an owned layer always has its +x neighbour in the held range)
			int cy = iy + dy;
			cy += (cy < 0) ? ny : 0;
			cy -= (cy >= ny) ? ny : 0;
			const int *cs = cell_start + (cx * ny + cy) * nz;
			const int zm = iz > 0 ? iz - 1 : nz - 1, zp = iz < nz - 1 ? iz + 1 : 0;
			// three segments (dz = -1, 0, +1); column 0 has no dz = -1 cell and starts its own cell after the particle itself
			int sb[3], se[3];
			sb[0] = col ? cs[zm] : 0;
			se[0] = col ? cs[zm + 1] : 0;
			sb[1] = col ? cs[iz] : sp + 1;
			se[1] = cs[iz + 1];
			sb[2] = cs[zp];
			se[2] = cs[zp + 1];
			// neighbour minus own shadow, read as signed integers: the minimum-image separation in units of box / 2^32
			const uint4 sh = posf[sp];
			const float ux = (float) (sc.box[0] / 4294967296.), uy = (float) (sc.box[1] / 4294967296.), uz = (float) (sc.box[2] / 4294967296.);
			const float thr = (float) sc.sqr_rcut + 2e-3f, core_hi = 1.f + 2e-3f;
			// orientation screen as in the sweep: a pair clear of the hard core can only bond if one of p's patches points at
			// q; p's matrix is fetched (once per thread, in FP32) when the first pair in range turns up
			const float point_thr = sc.fast_patch ? (float) (sc.cos2_lo * sc.inv_b2) * (1.f - 2e-3f) : -1.f;
			float mf[9];
			bool have_m = false;
#pragma unroll
			for(int g = 0; g < 3; g++) {
				for(int sq = sb[g]; sq < se[g]; sq++) {
					const uint4 qh = posf[sq];
					const float ax = shadow_sep(qh.x, sh.x, ux), ay = shadow_sep(qh.y, sh.y, uy), az = shadow_sep(qh.z, sh.z, uz);
					const float d2 = fmaf(az, az, fmaf(ay, ay, ax * ax));
					if(d2 >= thr) continue;
					if(d2 >= core_hi && point_thr >= 0.f) {
						if(!have_m) {
#pragma unroll
							for(int k = 0; k < 9; k++) mf[k] = (float) ori[9ll * sp + k];
							have_m = true;
						}
						if(!may_point(mf, ax, ay, az, d2, point_thr)) continue;
					}
					{
						const int slot = atomicAdd(&s_queued[warp], 1);
						if(slot < kQueue) s_queue[warp][slot] = make_int2(sp, sq);
						else bond_pair_exact(sc, pos, ori, sp, sq, bonds, overlap); // (dense packings: more pairs than the queue holds)
					}
				}
			}
		}
	}
	__syncwarp();
	{
		const int n = min(s_queued[warp], kQueue);
		for(int i = lane; i < n; i += 32) bond_pair_exact(sc, pos, ori, s_queue[warp][i].x, s_queue[warp][i].y, bonds, overlap);
	}
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		bonds += __shfl_xor_sync(0xffffffffu, bonds, o);
		overlap |= __shfl_xor_sync(0xffffffffu, overlap, o);
	}
	if((threadIdx.x & 31) == 0) {
		if(bonds) atomicAdd(&s_bonds, bonds);
		if(overlap) atomicOr(&s_overlap, 1);
	}
	__syncthreads();
	if(threadIdx.x == 0) {
		if(s_bonds) atomicAdd((unsigned long long *) &counters->bonds, (unsigned long long) s_bonds);
		if(s_overlap) atomicOr(&counters->overlap, 1);
	}
}

int launch_total_bonds(const SimConst &sc, const int *cell_start, const int *cell_sorted, const double4 *pos, const double *ori, const uint4 *posf,
		Counters *counters, cudaStream_t st) {
	k_zero_bonds<<<1, 1, 0, st>>>(counters);
	const bool wide = (sc.slab || sc.nxl >= 4) && sc.n_side[1] >= 4 && sc.n_side[2] >= 4;
	if(wide && posf != nullptr && sc.N > 0) {
		const long long threads = 5ll * sc.N;
		k_total_bonds_cells<<<(unsigned) ((threads + kBondCellThreads - 1) / kBondCellThreads), kBondCellThreads, 0, st>>>(sc, cell_start, cell_sorted, pos,
				ori, posf, counters);
		return 2;
	}
	constexpr int G = 8;
	const int tiles_per_block = kBondThreads / G;
	int blocks = (sc.n_cells + tiles_per_block - 1) / tiles_per_block;
	const int cap = 148 * 16 * 4;
	if(blocks > cap) blocks = cap;
	if(blocks < 1) blocks = 1;
	k_total_bonds<G><<<blocks, kBondThreads, 0, st>>>(sc, cell_start, pos, ori, counters);
	return 2;
}

// ---------------------------------------------------------------------------------------------
// Per-particle bonded-neighbour counts and, optionally, the ids of the bonded neighbours (the two lines
// output_print_bonds writes per particle, output.c:213-268).
// ---------------------------------------------------------------------------------------------
template<int G>
__global__ void __launch_bounds__(kBondThreads) k_bond_counts(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start,
		const double4 *__restrict__ pos, const double *__restrict__ ori, const int *__restrict__ pid, int *__restrict__ counts,
		int *__restrict__ neighbours, int max_nb) {
	constexpr int kTiles = kBondThreads / G;
	__shared__ Neighbourhood s_nb[kTiles];
	cg::thread_block block = cg::this_thread_block();
	cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
	const int tile_in_block = threadIdx.x / G;
	const int lane = tile.thread_rank();
	const int nyz = sc.n_side[1] * sc.n_side[2];
	for(int cell = blockIdx.x * kTiles + tile_in_block; cell < sc.n_cells; cell += gridDim.x * kTiles) {
		const int first = cell_start[cell], last = cell_start[cell + 1];
		const int ix = cell / nyz, iy = (cell - ix * nyz) / sc.n_side[2], iz = cell - ix * nyz - iy * sc.n_side[2];
		// only owned layers; a pair across a slab boundary is counted by the engine on its left, because the
		// right halo comes after every owned slot and the left halo before (slot_q > slot_p below)
		if(first == last || ix < sc.own_lo || ix >= sc.own_hi) continue;
		Neighbourhood &nb = s_nb[tile_in_block];
		tile.sync();
		const int total = build_neighbourhood<G>(tile, sc, cell_start, ix, iy, iz, nb);
		for(int sp = first; sp < last; sp++) {
			const double4 p4 = pos[sp];
			const Vec3 p = {p4.x, p4.y, p4.z};
			double mp[9], pp[12];
#pragma unroll
			for(int k = 0; k < 9; k++) mp[k] = ori[9ll * sp + k];
			patches_from(sc.base, mp, pp);
			int bonds = 0;
			const int id_p = pid[sp];
			for(int j0 = 0; j0 < total; j0 += G) { // whole tile in step: the neighbour ids are written in neighbourhood order
				const int j = j0 + lane;
				int sq = -1;
				bool bonded = false;
				if(j < total) {
					sq = neighbour_slot(nb, j);
					bonded = (sq != sp) && (pair_test(sc, p, pp, pos[sq], ori + 9ll * sq) == 1);
				}
				const unsigned hit = tile.ballot(bonded);
				if(neighbours != nullptr && bonded) {
					const int at = bonds + __popc(hit & ((1u << lane) - 1u));
					if(at < max_nb) neighbours[(long long) id_p * max_nb + at] = pid[sq];
				}
				bonds += __popc(hit);
			}
			if(lane == 0) counts[id_p] = bonds;
		}
	}
}

int launch_bond_counts(const SimConst &sc, const int *cell_start, const double4 *pos, const double *ori, const int *pid, int *counts, int *neighbours,
		int max_nb, cudaStream_t st) {
	constexpr int G = 8;
	const int tiles_per_block = kBondThreads / G;
	int blocks = (sc.n_cells + tiles_per_block - 1) / tiles_per_block;
	const int cap = 148 * 16 * 4;
	if(blocks > cap) blocks = cap;
	if(blocks < 1) blocks = 1;
	k_bond_counts<G><<<blocks, kBondThreads, 0, st>>>(sc, cell_start, pos, ori, pid, counts, neighbours, max_nb);
	return 1;
}

// ---------------------------------------------------------------------------------------------
// K4: rescale positions (MC.c:474-478, 540-545).  Out of place so that a rejected box move keeps
// the old coordinates bit for bit (the reference divides back, MC.c:505-508, 567-572).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rescale(int N, const double4 *__restrict__ src, double4 *__restrict__ dst, double fx, double fy, double fz,
		int divide) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	double4 p = src[i];
	if(divide) { // undo of a rejected move the way the reference does it (MC.c:505-508)
		p.x = __ddiv_rn(p.x, fx);
		p.y = __ddiv_rn(p.y, fy);
		p.z = __ddiv_rn(p.z, fz);
	}
	else {
		p.x = __dmul_rn(p.x, fx);
		p.y = __dmul_rn(p.y, fy);
		p.z = __dmul_rn(p.z, fz);
	}
	dst[i] = p;
}

int launch_rescale(int N, const double4 *src, double4 *dst, double fx, double fy, double fz, int divide, cudaStream_t st) {
	if(N == 0) return 0;
	k_rescale<<<(N + 255) / 256, 256, 0, st>>>(N, src, dst, fx, fy, fz, divide);
	return 1;
}

// ---------------------------------------------------------------------------------------------
// K5: MC.c:328-335 for a list of proposals against the frozen configuration.  One tile per
// proposal.  The old energy is taken around the cell of the current position and the new one
// around the cell of the proposed position, as MC_energy does (MC.c:277).
// ---------------------------------------------------------------------------------------------
template<int G>
__device__ __forceinline__ void probe_bonds(const cg::thread_block_tile<G> &tile, const SimConst &sc, const int *__restrict__ cell_start,
		const double4 *__restrict__ pos, const double *__restrict__ ori, Neighbourhood &nb, const Vec3 &p, const double *pp, int self_slot,
		int &bonds, int &overlap) {
	int ix, iy, iz;
	cell_index(sc, p.x, p.y, p.z, ix, iy, iz);
	tile.sync();
	const int total = build_neighbourhood<G>(tile, sc, cell_start, ix, iy, iz, nb);
	int b = 0, o = 0;
	for(int j = tile.thread_rank(); j < total; j += G) {
		const int sq = neighbour_slot(nb, j);
		if(sq == self_slot) continue;
		const int v = pair_test(sc, p, pp, pos[sq], ori + 9ll * sq);
		b += (v == 1);
		o |= (v == -1);
	}
	bonds = cg::reduce(tile, b, cg::plus<int>());
	overlap = cg::reduce(tile, o, cg::bit_or<int>());
}

template<int G>
__global__ void __launch_bounds__(kBondThreads) k_proposals(const __grid_constant__ SimConst sc, int n, const ProposalIO io,
		const int *__restrict__ cell_start, const double4 *__restrict__ pos, const double *__restrict__ ori, const int *__restrict__ slot_of) {
	constexpr int kTiles = kBondThreads / G;
	__shared__ Neighbourhood s_nb[kTiles];
	cg::thread_block block = cg::this_thread_block();
	cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
	const int tile_in_block = threadIdx.x / G;
	Neighbourhood &nb = s_nb[tile_in_block];
	for(int k = blockIdx.x * kTiles + tile_in_block; k < n; k += gridDim.x * kTiles) {
		const int sp = slot_of[io.idx[k]];
		const double4 p4 = pos[sp];
		const Vec3 p_old = {p4.x, p4.y, p4.z};
		const Vec3 p_new = {__dadd_rn(p4.x, io.disp[3 * k]), __dadd_rn(p4.y, io.disp[3 * k + 1]), __dadd_rn(p4.z, io.disp[3 * k + 2])};
		double m[9], pp_old[12], pp_new[12];
#pragma unroll
		for(int c = 0; c < 9; c++) m[c] = ori[9ll * sp + c];
		patches_from(sc.base, m, pp_old);
#pragma unroll
		for(int c = 0; c < 9; c++) m[c] = io.orient[9ll * k + c];
		patches_from(sc.base, m, pp_new);

		int b_old, o_old, b_new, o_new;
		probe_bonds<G>(tile, sc, cell_start, pos, ori, nb, p_old, pp_old, sp, b_old, o_old);
		probe_bonds<G>(tile, sc, cell_start, pos, ori, nb, p_new, pp_new, sp, b_new, o_new);
		if(o_new) b_new = 0; // the reference returns E = 0 at the first overlap (MC.c:296-299)
		if(tile.thread_rank() == 0) {
			const int dE = b_old - b_new; // energy = -bonds
			int acc = 0;
			if(!o_new) acc = (dE < 0) || (io.u[k] < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]);
			int a, b, c, a2, b2, c2;
			const int c_old = cell_index(sc, p_old.x, p_old.y, p_old.z, a, b, c);
			const int c_new = cell_index(sc, p_new.x, p_new.y, p_new.z, a2, b2, c2);
			io.dE[k] = dE;
			io.overlap[k] = o_new;
			io.accept[k] = acc;
			io.cell_changed[k] = (c_old != c_new);
		}
	}
}

int launch_proposals(const SimConst &sc, int n, const ProposalIO &io, const int *cell_start, const double4 *pos, const double *ori, const int *slot_of,
		cudaStream_t st) {
	if(n == 0) return 0;
	constexpr int G = 8;
	const int tiles_per_block = kBondThreads / G;
	int blocks = (n + tiles_per_block - 1) / tiles_per_block;
	const int cap = 148 * 16 * 4;
	if(blocks > cap) blocks = cap;
	k_proposals<G><<<blocks, kBondThreads, 0, st>>>(sc, n, io, cell_start, pos, ori, slot_of);
	return 1;
}

__global__ void k_commit(const __grid_constant__ SimConst sc, int slot, const double *__restrict__ disp, const double *__restrict__ orient, double4 *pos,
		double *ori, uint4 *posf) {
	if(threadIdx.x == 0) {
		double4 p = pos[slot];
		p.x = __dadd_rn(p.x, disp[0]);
		p.y = __dadd_rn(p.y, disp[1]);
		p.z = __dadd_rn(p.z, disp[2]);
		pos[slot] = p;
		store_shadow_xyz(posf + slot, box_shadow(sc, p.x, p.y, p.z, slot)); // the shadow follows
	}
	if(threadIdx.x < 9) ori[9ll * slot + threadIdx.x] = orient[threadIdx.x];
}

int launch_commit(const SimConst &sc, int slot, const double *disp3, const double *orient9, double4 *pos, double *ori, uint4 *posf, cudaStream_t st) {
	k_commit<<<1, 32, 0, st>>>(sc, slot, disp3, orient9, pos, ori, posf);
	return 1;
}

// ---------------------------------------------------------------------------------------------
// K6: Gram-Schmidt on the rows of every orientation matrix (utils.c:91-113 as used by
// MC_check_energy, MC.c:774).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_orthonormalise(int N, double *__restrict__ ori) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	double m[9];
#pragma unroll
	for(int k = 0; k < 9; k++) m[k] = ori[9ll * i + k];
	double *v1 = m, *v2 = m + 3, *v3 = m + 6;
	const double v1n2 = dot3_rn(v1[0], v1[1], v1[2], v1[0], v1[1], v1[2]);
	const double b1 = __ddiv_rn(dot3_rn(v2[0], v2[1], v2[2], v1[0], v1[1], v1[2]), v1n2);
#pragma unroll
	for(int k = 0; k < 3; k++) v2[k] = __dsub_rn(v2[k], __dmul_rn(b1, v1[k]));
	const double v3v1 = dot3_rn(v3[0], v3[1], v3[2], v1[0], v1[1], v1[2]);
	const double v3v2 = dot3_rn(v3[0], v3[1], v3[2], v2[0], v2[1], v2[2]);
	const double v2n2 = dot3_rn(v2[0], v2[1], v2[2], v2[0], v2[1], v2[2]);
	const double c1 = __ddiv_rn(v3v1, v1n2), c2 = __ddiv_rn(v3v2, v2n2);
#pragma unroll
	for(int k = 0; k < 3; k++) v3[k] = __dsub_rn(v3[k], __dadd_rn(__dmul_rn(c1, v1[k]), __dmul_rn(c2, v2[k])));
#pragma unroll
	for(int r = 0; r < 3; r++) {
		double *v = m + 3 * r;
		const double n = __dsqrt_rn(dot3_rn(v[0], v[1], v[2], v[0], v[1], v[2]));
		v[0] = __ddiv_rn(v[0], n);
		v[1] = __ddiv_rn(v[1], n);
		v[2] = __ddiv_rn(v[2], n);
	}
#pragma unroll
	for(int k = 0; k < 9; k++) ori[9ll * i + k] = m[k];
}

int launch_orthonormalise(int N, double *ori, cudaStream_t st) {
	if(N == 0) return 0;
	k_orthonormalise<<<(N + 255) / 256, 256, 0, st>>>(N, ori);
	return 1;
}

// ---------------------------------------------------------------------------------------------
// upload / download: caller order <-> slot order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_state(int N, const double *__restrict__ r, const double *__restrict__ orient, const int *__restrict__ ids,
		double4 *__restrict__ pos, double *__restrict__ ori, int *__restrict__ pid) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const int id = ids ? ids[i] : i; // a slab engine receives a subset of the system with its global ids
	pos[i] = make_double4(r[3ll * i], r[3ll * i + 1], r[3ll * i + 2], (double) id);
	pid[i] = id;
#pragma unroll
	for(int k = 0; k < 9; k++) ori[9ll * i + k] = orient[9ll * i + k];
}

__global__ void __launch_bounds__(256) k_unpack_state(int N, const double4 *__restrict__ pos, const double *__restrict__ ori, const int *__restrict__ pid,
		double *__restrict__ r, double *__restrict__ orient) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= N) return;
	const int i = pid[s];
	const double4 p = pos[s];
	r[3ll * i] = p.x;
	r[3ll * i + 1] = p.y;
	r[3ll * i + 2] = p.z;
#pragma unroll
	for(int k = 0; k < 9; k++) orient[9ll * i + k] = ori[9ll * s + k];
}

// A slab engine takes its own particles out of the whole configuration in device-readable host memory: layer of every
// particle with the engine's cell arithmetic (unshifted grid), per-layer histogram, and -- when pos is given -- the
// selected particles gathered into the device arrays in arrival order (the cell build orders them by id).
// hist: [0] selected, [1] overflow flag, [2 .. 2 + nx) particles per global layer.
__global__ void __launch_bounds__(256) k_select_pack(int N, const double *__restrict__ r, const double *__restrict__ orient, double inv_box, int nx, int lo,
		int held, double4 *__restrict__ pos, double *__restrict__ ori, int *__restrict__ pid, int cap, int *__restrict__ hist) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const double x = r[3ll * i];
	const int layer = cell_coord(x, 0., inv_box, nx);
	atomicAdd(hist + 2 + layer, 1);
	int rel = layer - lo;
	rel += rel < 0 ? nx : 0;
	if(rel >= held) return;
	const int slot = atomicAdd(hist, 1);
	if(pos == nullptr) return;
	if(slot >= cap) {
		hist[1] = 1;
		return;
	}
	pos[slot] = make_double4(x, r[3ll * i + 1], r[3ll * i + 2], (double) i);
	pid[slot] = i;
#pragma unroll
	for(int k = 0; k < 9; k++) ori[9ll * slot + k] = orient[9ll * i + k];
}

int launch_select_pack(int N, const double *r, const double *orient, double inv_box, int nx, int lo, int held, double4 *pos, double *ori, int *pid, int cap,
		int *hist, cudaStream_t st) {
	if(N == 0) return 0;
	k_select_pack<<<(N + 255) / 256, 256, 0, st>>>(N, r, orient, inv_box, nx, lo, held, pos, ori, pid, cap, hist);
	return 1;
}

int launch_pack_state(int N, const double *r, const double *orient, const int *ids, double4 *pos, double *ori, int *pid, cudaStream_t st) {
	if(N == 0) return 0;
	k_pack_state<<<(N + 255) / 256, 256, 0, st>>>(N, r, orient, ids, pos, ori, pid);
	return 1;
}

int launch_unpack_state(int N, const double4 *pos, const double *ori, const int *pid, double *r, double *orient, cudaStream_t st) {
	if(N == 0) return 0;
	k_unpack_state<<<(N + 255) / 256, 256, 0, st>>>(N, pos, ori, pid, r, orient);
	return 1;
}

} // namespace pp
