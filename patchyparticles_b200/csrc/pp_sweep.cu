// pp_sweep.cu -- K2: the 8-colour checkerboard sweep.
//
// Replaces the inner loop of do_NVT / do_NPT (MC.c:19-29, 60-68) together with
// MC_move_rototranslate, MC_energy, MC_would_interact, MC_rototraslate_particle and
// MC_rollback_particle (MC.c:146-340).
//
// Scheme (DESIGN.md "checkerboard sweep"): the cell grid has an even number of cells of side
// >= r_cut per axis and a random origin shift.  Cells of one colour (parity of (ix, iy, iz)) are
// at least one cell apart, so particles of two active cells never interact as long as a move may
// not take a particle out of its cell; such proposals are rejected before any energy is
// evaluated.  An active cell holding n particles makes n trials, each on a member drawn uniformly
// with replacement, so a sweep is exactly N attempts like a reference step.
//
// Two kernels:
//   k_make_proposals  one thread per trial of the whole sweep: Philox4x32-10 keyed by the seed,
//                     counter (cell, trial, sweep, block) -> picked member, displacement, Marsaglia
//                     axis, Rodrigues matrix, Metropolis uniform (MC.c:312-324).  Fully parallel,
//                     off the critical path of the passes.
//   k_sweep_pass      one tile of G lanes per non-empty cell of the current colour, taken from
//                     the work list (cells bucketed by occupancy, so the tiles of a warp run the
//                     same number of trials).  The (deduplicated) 27-cell neighbourhood is staged
//                     once in shared memory; every trial then (1) shares the 3x3 product and the
//                     patch vectors of the moved particle across the lanes, (2) lets the lanes
//                     split the staged neighbours for the FP64 distance tests of the old and the
//                     new position, compacting the pairs inside the bonding shell into a candidate
//                     list, (3) runs the orientation tests on the candidates, (4) reduces bond
//                     counts / overlap across the tile, applies the Metropolis rule and writes the
//                     accepted state back to HBM and to the staged copy.
// A trajectory depends only on (seed, sweep, cell, trial): not on G, the block size, or the GPU.
#include "pp_kernels.cuh"

namespace pp {

constexpr int kStageCap = 64; // staged neighbours per cell; denser neighbourhoods read HBM/L2 directly

// ---------------------------------------------------------------------------------------------
// K2a: proposals
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(Proposal) == 7 * 16, "k_make_proposals copies its records out as seven 16-byte words each");
__global__ void __launch_bounds__(128) k_make_proposals(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start,
		const int *__restrict__ cell_sorted, Proposal *__restrict__ out) {
	// the 112-byte records of a block leave through shared memory: 16-byte stores of consecutive lanes to consecutive addresses
	// instead of seven stores of stride 112 per thread
	__shared__ __align__(16) Proposal s_out[128];
	__shared__ unsigned char s_mine[128];
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	bool mine = i < sc.N;
	int lcell = 0, lx = 0, ly = 0, lz = 0;
	if(mine) {
		lcell = cell_sorted[i];
		const int nyz = sc.n_side[1] * sc.n_side[2];
		lx = lcell / nyz;
		ly = (lcell - lx * nyz) / sc.n_side[2];
		lz = lcell - lx * nyz - ly * sc.n_side[2];
		mine = lx >= sc.own_lo && lx < sc.own_hi; // halo particle: moved by its owner (its record is never read)
	}
	if(mine) {
	const int first = cell_start[lcell];
	const int n_here = cell_start[lcell + 1] - first;
	const int t = i - first;
	const int cell = global_cell(sc, lx, ly, lz); // random streams are keyed by the global cell
	const uint4 x = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | 0u, sc.seed_lo, sc.seed_hi);
	const uint4 y = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | 1u, sc.seed_lo, sc.seed_hi);
	Proposal pr;
	pr.pick = (int) __umulhi(x.x, (uint32_t) n_here);
	pr.dx = __dmul_rn(__dsub_rn(u32_to_unit(x.y), 0.5), sc.disp_max);
	pr.dy = __dmul_rn(__dsub_rn(u32_to_unit(x.z), 0.5), sc.disp_max);
	pr.dz = __dmul_rn(__dsub_rn(u32_to_unit(x.w), 0.5), sc.disp_max);
	{ // weight in the acceptance estimator: 1 / P(this displacement stays inside a uniformly shifted cell)
		const double keep = __dmul_rn(__dmul_rn(__dsub_rn(1., __dmul_rn(fabs(pr.dx), sc.inv_w[0])), __dsub_rn(1., __dmul_rn(fabs(pr.dy), sc.inv_w[1]))),
				__dsub_rn(1., __dmul_rn(fabs(pr.dz), sc.inv_w[2])));
		// a displacement comparable to the cell width (keep -> 0) must not wrap the 2^-24 fixed point: weights are capped at 128
		pr.wfix = (unsigned int) (unsigned long long) __dmul_rn(__ddiv_rn(1., fmax(keep, 0.0078125)), 16777216.0);
	}
	const double theta = __dmul_rn(u32_to_unit(y.x), sc.theta_max);
	pr.u_acc = __dmul_rn((double) ((((unsigned long long) y.y) << 21) | (((unsigned long long) y.z) >> 11)), 1.0 / 9007199254740992.0);
	double ran1, ran2, ransq;
	for(uint32_t blk = 2;; blk++) { // Marsaglia pairs (utils.c:120-124)
		const uint4 z = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | blk, sc.seed_lo, sc.seed_hi);
		ran1 = __dsub_rn(1., __dmul_rn(2., u32_to_unit(z.x)));
		ran2 = __dsub_rn(1., __dmul_rn(2., u32_to_unit(z.y)));
		ransq = __dadd_rn(__dmul_rn(ran1, ran1), __dmul_rn(ran2, ran2));
		if(ransq < 1.) break;
		ran1 = __dsub_rn(1., __dmul_rn(2., u32_to_unit(z.z)));
		ran2 = __dsub_rn(1., __dmul_rn(2., u32_to_unit(z.w)));
		ransq = __dadd_rn(__dmul_rn(ran1, ran1), __dmul_rn(ran2, ran2));
		if(ransq < 1.) break;
	}
	// axis (utils.c:127-131), Rodrigues matrix (utils.c:186-202)
	const double ranh = __dmul_rn(2., __dsqrt_rn(__dsub_rn(1., ransq)));
	const double ex = __dmul_rn(ran1, ranh), ey = __dmul_rn(ran2, ranh), ez = __dsub_rn(1., __dmul_rn(2., ransq));
	double st, ct;
	sincos(theta, &st, &ct);
	const double uct = __dsub_rn(1., ct);
	const double xx = __dmul_rn(ex, ex), yy = __dmul_rn(ey, ey), zz = __dmul_rn(ez, ez);
	const double xy = __dmul_rn(__dmul_rn(ex, ey), uct), xz = __dmul_rn(__dmul_rn(ex, ez), uct), yz = __dmul_rn(__dmul_rn(ey, ez), uct);
	const double xs = __dmul_rn(ex, st), ys = __dmul_rn(ey, st), zs = __dmul_rn(ez, st);
	pr.R[0] = __dadd_rn(xx, __dmul_rn(__dsub_rn(1., xx), ct));
	pr.R[1] = __dsub_rn(xy, zs);
	pr.R[2] = __dadd_rn(xz, ys);
	pr.R[3] = __dadd_rn(xy, zs);
	pr.R[4] = __dadd_rn(yy, __dmul_rn(__dsub_rn(1., yy), ct));
	pr.R[5] = __dsub_rn(yz, xs);
	pr.R[6] = __dsub_rn(xz, ys);
	pr.R[7] = __dadd_rn(yz, xs);
	pr.R[8] = __dadd_rn(zz, __dmul_rn(__dsub_rn(1., zz), ct));
	s_out[threadIdx.x] = pr;
	}
	s_mine[threadIdx.x] = mine;
	__syncthreads();
	{
		const long long base = (long long) blockIdx.x * blockDim.x; // first record of the block
		const int n_rec = (int) min((long long) blockDim.x, (long long) sc.N - base);
		const uint4 *src = reinterpret_cast<const uint4 *>(s_out);
		uint4 *dst = reinterpret_cast<uint4 *>(out + base);
		for(int k = threadIdx.x; k < 7 * n_rec; k += blockDim.x) // (sizeof(Proposal) = 7 x 16)
			if(s_mine[k / 7]) dst[k] = src[k]; // (a halo particle's record is never read: not written either, it would only take room in L2)
	}
}

int launch_proposals_for_sweep(const SimConst &sc, const CellBuffers &cb, Proposal *proposals, cudaStream_t st) {
	if(sc.N == 0) return 0;
	k_make_proposals<<<(sc.N + 127) / 128, 128, 0, st>>>(sc, cb.cell_start, cb.cell_sorted, proposals);
	return 1;
}

// ---------------------------------------------------------------------------------------------
// K2b: colour pass
// ---------------------------------------------------------------------------------------------
struct BlockCounters {
	unsigned int tries, accepted, confined, overlapped;
	int dbonds;
	unsigned long long w_tries, w_accepted, pair_checks;
};

template<int G>
struct TileSmem {
	Neighbourhood nb;
	double x[kStageCap], y[kStageCap], z[kStageCap];
	int slot[kStageCap];
	double prop[14]; // R[9], dx, dy, dz, u_acc, (pick, wfix)
	double m_old[9], m_new[9], pp_old[12], pp_new[12];
	unsigned short cand[2 * kStageCap]; // staged index | (new ? 0x8000 : 0)
};

template<int G>
__host__ __device__ constexpr int sweep_threads() {
	return G == 4 ? 32 : 64;
}

template<int G>
__global__ void __launch_bounds__(sweep_threads<G>(), 512 / sweep_threads<G>()) k_sweep_pass(const __grid_constant__ SimConst sc, const int *__restrict__ range,
		const int *__restrict__ cell_start, const int *__restrict__ work, const Proposal *__restrict__ proposals,
		double4 *pos, double *ori, uint4 *posf, Counters *__restrict__ counters, const __grid_constant__ HaloMirror hm) {
	constexpr int kThreads = sweep_threads<G>();
	constexpr int kTiles = kThreads / G;
	__shared__ TileSmem<G> s_tile[kTiles];
	__shared__ BlockCounters s_cnt;
	cg::thread_block block = cg::this_thread_block();
	cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
	const int lane = tile.thread_rank();
	TileSmem<G> &ts = s_tile[threadIdx.x / G];
	if(threadIdx.x == 0) s_cnt = BlockCounters{0, 0, 0, 0, 0, 0, 0, 0};
	__syncthreads();

	const int w_begin = range[0], w_end = range[1];
	const int nyz = sc.n_side[1] * sc.n_side[2];
	const int hm_off = mirror_offset(hm);
	// per-tile tallies (kept by every lane, flushed by lane 0)
	unsigned int a_tries = 0, a_acc = 0, a_conf = 0, a_ovl = 0;
	unsigned long long a_wt = 0, a_wa = 0, a_checks = 0;
	int a_dbonds = 0;

	for(int wi = w_begin + blockIdx.x * kTiles + (int) (threadIdx.x / G); wi < w_end; wi += gridDim.x * kTiles) {
		const int cell = work[wi];
		const int ix = cell / nyz, iy = (cell - ix * nyz) / sc.n_side[2], iz = cell - ix * nyz - iy * sc.n_side[2];
		const int first = cell_start[cell];
		const int n_here = cell_start[cell + 1] - first;
		tile.sync();
		const int total = build_neighbourhood<G>(tile, sc, cell_start, ix, iy, iz, ts.nb);
		const bool staged = total <= kStageCap;
		if(staged) {
			for(int j = lane; j < total; j += G) {
				const int sq = neighbour_slot(ts.nb, j);
				const double4 q4 = pos[sq];
				ts.x[j] = q4.x;
				ts.y[j] = q4.y;
				ts.z[j] = q4.z;
				ts.slot[j] = sq;
			}
		}
		const int home_base = ts.nb.incl[12]; // the home cell is entry 13 of the 27-cell table

		for(int t = 0; t < n_here; t++) {
			tile.sync();
			{ // fetch the proposal of this trial (112 contiguous bytes)
				const double *src = reinterpret_cast<const double *>(proposals + first + t);
				for(int k = lane; k < 14; k += G) ts.prop[k] = src[k];
			}
			tile.sync();
			const double *R = ts.prop;
			const double dx = ts.prop[9], dy = ts.prop[10], dz = ts.prop[11], u_acc = ts.prop[12];
			const int pick = reinterpret_cast<const int *>(ts.prop + 13)[0];
			const unsigned long long Wfix = (unsigned long long) reinterpret_cast<const unsigned int *>(ts.prop + 13)[1];
			a_tries++;
			const int sp = first + pick;
			const int j_self = home_base + pick;
			Vec3 p_old;
			if(staged) p_old = {ts.x[j_self], ts.y[j_self], ts.z[j_self]};
			else {
				const double4 p4 = pos[sp];
				p_old = {p4.x, p4.y, p4.z};
			}
			const Vec3 p_new = {__dadd_rn(p_old.x, dx), __dadd_rn(p_old.y, dy), __dadd_rn(p_old.z, dz)};
			int cx, cy, cz;
			if(cell_index(sc, p_new.x, p_new.y, p_new.z, cx, cy, cz) != cell) { // would leave its cell
				a_conf++;
				continue;
			}
			a_wt += Wfix;

			// ---- distance tests of the old and new position (MC.c:233-241) --------------------
			for(int k = lane; k < 9; k += G) ts.m_old[k] = ori[9ll * sp + k]; // in flight during the distance tests
			int o_new = 0, n_cand = 0;
			for(int j0 = 0; j0 < total; j0 += G) {
				const int j = j0 + lane;
				bool in_old = false, in_new = false;
				if(j < total) {
					int sq = 0;
					double qx, qy, qz;
					bool self;
					if(staged) {
						self = (j == j_self);
						qx = ts.x[j];
						qy = ts.y[j];
						qz = ts.z[j];
					}
					else {
						sq = neighbour_slot(ts.nb, j);
						self = (sq == sp);
						const double4 q4 = pos[sq];
						qx = q4.x;
						qy = q4.y;
						qz = q4.z;
					}
					if(!self) {
						Vec3 d;
						const double d2_old = min_image(sc, p_old, qx, qy, qz, d);
						const double d2_new = min_image(sc, p_new, qx, qy, qz, d);
						if(d2_new < 1.) o_new = 1;
						in_old = (d2_old >= 1. && d2_old < sc.sqr_rcut);
						in_new = (d2_new >= 1. && d2_new < sc.sqr_rcut);
					}
				}
				if(staged) { // compact the in-shell pairs into the candidate list
					const unsigned m_old_ = tile.ballot(in_old), m_new_ = tile.ballot(in_new);
					const unsigned below = (1u << lane) - 1u;
					if(in_old) ts.cand[n_cand + __popc(m_old_ & below)] = (unsigned short) j;
					n_cand += __popc(m_old_);
					if(in_new) ts.cand[n_cand + __popc(m_new_ & below)] = (unsigned short) (j | 0x8000);
					n_cand += __popc(m_new_);
				}
			}
			a_checks += 2ull * (unsigned long long) (total - 1);
			if(tile.any(o_new)) { // hard-core overlap: rejected whatever the bonds (MC.c:296-299, 332)
				a_ovl++;
				continue;
			}

			// ---- new orientation and both sets of patches, one element per lane -----------
			tile.sync();
			for(int k = lane; k < 9; k += G) {
				const int i = k / 3, j = k - 3 * i; // new = old . R (utils.c:172-184)
				ts.m_new[k] = dot3_rn(ts.m_old[3 * i], ts.m_old[3 * i + 1], ts.m_old[3 * i + 2], R[j], R[3 + j], R[6 + j]);
			}
			tile.sync();
			for(int k = lane; k < 24; k += G) { // patch[i][c] = M[c][:] . base[i][:]  (MC.c:225-227)
				const int kk = k < 12 ? k : k - 12;
				const int i = kk / 3, c = kk - 3 * i;
				const double *m = k < 12 ? ts.m_old : ts.m_new;
				const double v = dot3_rn(m[3 * c], m[3 * c + 1], m[3 * c + 2], sc.base[3 * i], sc.base[3 * i + 1], sc.base[3 * i + 2]);
				if(k < 12) ts.pp_old[kk] = v;
				else ts.pp_new[kk] = v;
			}
			tile.sync();

			// ---- orientation tests on the pairs inside the bonding shell (MC.c:241-262) ---------
			int b_old = 0, b_new = 0;
			if(staged) {
				for(int c = lane; c < n_cand; c += G) {
					const unsigned e = ts.cand[c];
					const int j = (int) (e & 0x7fffu);
					const bool is_new = (e & 0x8000u) != 0;
					Vec3 d;
					const double d2 = min_image(sc, is_new ? p_new : p_old, ts.x[j], ts.y[j], ts.z[j], d);
					const bool bond = patches_bond(sc, d, d2, is_new ? ts.pp_new : ts.pp_old, ori + 9ll * ts.slot[j]);
					b_old += (bond && !is_new);
					b_new += (bond && is_new);
				}
			}
			else {
				for(int j = lane; j < total; j += G) {
					const int sq = neighbour_slot(ts.nb, j);
					if(sq == sp) continue;
					const double4 q4 = pos[sq];
					Vec3 d_old, d_new;
					const double d2_old = min_image(sc, p_old, q4.x, q4.y, q4.z, d_old);
					const double d2_new = min_image(sc, p_new, q4.x, q4.y, q4.z, d_new);
					if(d2_old >= 1. && d2_old < sc.sqr_rcut) b_old += patches_bond(sc, d_old, d2_old, ts.pp_old, ori + 9ll * sq) ? 1 : 0;
					if(d2_new >= 1. && d2_new < sc.sqr_rcut) b_new += patches_bond(sc, d_new, d2_new, ts.pp_new, ori + 9ll * sq) ? 1 : 0;
				}
			}
			const unsigned packed = cg::reduce(tile, (unsigned) b_old | ((unsigned) b_new << 16), cg::plus<unsigned>());
			b_old = (int) (packed & 0xffffu);
			b_new = (int) (packed >> 16);

			// ---- Metropolis (MC.c:332-339) ------------------------------------------------------
			const int dE = b_old - b_new; // energy = -bonds
			if(dE < 0 || u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
				if(lane == 0) {
					double *p = reinterpret_cast<double *>(pos + sp); // x, y, z; the id in .w stays
					p[0] = p_new.x;
					p[1] = p_new.y;
					p[2] = p_new.z;
					const uint4 moved = box_shadow(sc, p_new.x, p_new.y, p_new.z, sp);
					store_shadow_xyz(posf + sp, moved);
					mirror_state(hm, hm_off, sp, p_new, moved, ts.m_new);
					if(staged) {
						ts.x[j_self] = p_new.x;
						ts.y[j_self] = p_new.y;
						ts.z[j_self] = p_new.z;
					}
				}
				for(int k = lane; k < 9; k += G) ori[9ll * sp + k] = ts.m_new[k];
				a_dbonds += b_new - b_old;
				a_acc++;
				a_wa += Wfix;
			}
		}
	}

	if(lane == 0 && a_tries) {
		atomicAdd(&s_cnt.tries, a_tries);
		atomicAdd(&s_cnt.accepted, a_acc);
		atomicAdd(&s_cnt.confined, a_conf);
		atomicAdd(&s_cnt.overlapped, a_ovl);
		atomicAdd(&s_cnt.dbonds, a_dbonds);
		atomicAdd(&s_cnt.w_tries, a_wt);
		atomicAdd(&s_cnt.w_accepted, a_wa);
		atomicAdd(&s_cnt.pair_checks, a_checks);
	}
	__syncthreads();
	if(threadIdx.x == 0 && s_cnt.tries) {
		atomicAdd((unsigned long long *) &counters->tries, (unsigned long long) s_cnt.tries);
		atomicAdd((unsigned long long *) &counters->accepted, (unsigned long long) s_cnt.accepted);
		atomicAdd((unsigned long long *) &counters->confined, (unsigned long long) s_cnt.confined);
		atomicAdd((unsigned long long *) &counters->overlapped, (unsigned long long) s_cnt.overlapped);
		atomicAdd(&counters->w_tries, s_cnt.w_tries);
		atomicAdd(&counters->w_accepted, s_cnt.w_accepted);
		atomicAdd((unsigned long long *) &counters->pair_checks, s_cnt.pair_checks);
		atomicAdd((unsigned long long *) &counters->dbonds, (unsigned long long) (long long) s_cnt.dbonds);
	}
}

template<int G>
static int launch_pass_g(const SimConst &sc, const int *range, const CellBuffers &cb, const Proposal *proposals, double4 *pos, double *ori, Counters *counters,
		const HaloMirror &hm, cudaStream_t st) {
	constexpr int kThreads = sweep_threads<G>();
	// the number of non-empty cells of this colour lives on the device (work_start); size the grid for
	// the worst case (every active cell occupied, at most one cell per particle) and let surplus tiles exit
	long long n_active = (long long) (sc.nxl / 2 + 1) * (sc.n_side[1] / 2) * (sc.n_side[2] / 2);
	if(n_active > sc.N) n_active = sc.N;
	const int tiles_per_block = kThreads / G;
	long long blocks = (n_active + tiles_per_block - 1) / tiles_per_block;
	const long long cap = 148ll * 8 * 3; // three waves of resident blocks; tiles stride over the occupancy-sorted list
	if(blocks > cap) blocks = cap;
	if(blocks < 1) blocks = 1;
	k_sweep_pass<G><<<(unsigned) blocks, kThreads, 0, st>>>(sc, range, cb.cell_start, cb.work, proposals, pos, ori, cb.posf, counters, hm);
	return 1;
}

int launch_sweep_pass(const SimConst &sc, const int *range, int lanes, const CellBuffers &cb, const Proposal *proposals, double4 *pos, double *ori,
		Counters *counters, const HaloMirror &hm, cudaStream_t st) {
	switch(lanes) {
	case 4:
		return launch_pass_g<4>(sc, range, cb, proposals, pos, ori, counters, hm, st);
	case 8:
		return launch_pass_g<8>(sc, range, cb, proposals, pos, ori, counters, hm, st);
	case 16:
		return launch_pass_g<16>(sc, range, cb, proposals, pos, ori, counters, hm, st);
	default:
		return launch_pass_g<32>(sc, range, cb, proposals, pos, ori, counters, hm, st);
	}
}

} // namespace pp
