// pp_sweep.cu -- K2: one colour pass of the 8-colour checkerboard sweep.
//
// Replaces the inner loop of do_NVT / do_NPT (MC.c:19-29, 60-68) together with
// MC_move_rototranslate, MC_energy, MC_would_interact, MC_rototraslate_particle and
// MC_rollback_particle (MC.c:146-340).
//
// Scheme (DESIGN.md "checkerboard sweep"): the cell grid has an even number of cells of side
// >= r_cut per axis and a per-sweep random origin shift.  Cells of one colour (parity of
// (ix, iy, iz)) are at least one cell apart, so particles of two active cells never interact as
// long as a move may not take a particle out of its cell; such proposals are rejected before any
// energy is evaluated.  An active cell holding n particles makes n trials, each on a member drawn
// uniformly with replacement, so a sweep is exactly N attempts like a reference step.
//
// One tile of G lanes per active cell: the lanes split the particles of the (deduplicated)
// 27-cell neighbourhood, evaluate the old and the new pair state in FP64 with the reference's
// operation order, and reduce bond counts / overlap with tile shuffles.  Random numbers come
// from Philox4x32-10 keyed by the seed with counter (cell, trial, sweep, block), so a trajectory
// does not depend on how cells are mapped to tiles, blocks or GPUs.
#include "pp_kernels.cuh"

namespace pp {

constexpr int kSweepThreads = 128;

struct BlockCounters {
	unsigned long long tries, accepted, confined, overlapped, w_tries, w_accepted, pair_checks;
	long long dbonds;
};

template<int G>
__global__ void __launch_bounds__(kSweepThreads) k_sweep_pass(const __grid_constant__ SimConst sc, int colour, const int *__restrict__ cell_start,
		double4 *pos, double *ori, Counters *__restrict__ counters) {
	constexpr int kTiles = kSweepThreads / G;
	__shared__ Neighbourhood s_nb[kTiles];
	__shared__ BlockCounters s_cnt;
	cg::thread_block block = cg::this_thread_block();
	cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
	const int tile_in_block = threadIdx.x / G;
	const int lane = tile.thread_rank();
	if(threadIdx.x == 0) s_cnt = BlockCounters{0, 0, 0, 0, 0, 0, 0, 0};
	__syncthreads();

	const int ox = colour & 1, oy = (colour >> 1) & 1, oz = (colour >> 2) & 1;
	const int hx = sc.n_side[0] >> 1, hy = sc.n_side[1] >> 1, hz = sc.n_side[2] >> 1;
	const int n_active = hx * hy * hz;

	// lane-0 accumulators of this tile
	unsigned long long a_tries = 0, a_acc = 0, a_conf = 0, a_ovl = 0, a_wt = 0, a_wa = 0, a_checks = 0;
	long long a_dbonds = 0;

	for(int a = blockIdx.x * kTiles + tile_in_block; a < n_active; a += gridDim.x * kTiles) {
		const int ax = a / (hy * hz), ay = (a - ax * hy * hz) / hz, az = a - ax * hy * hz - ay * hz;
		const int ix = 2 * ax + ox, iy = 2 * ay + oy, iz = 2 * az + oz;
		const int cell = (ix * sc.n_side[1] + iy) * sc.n_side[2] + iz;
		const int first = cell_start[cell];
		const int n_here = cell_start[cell + 1] - first;
		if(n_here == 0) continue;
		Neighbourhood &nb = s_nb[tile_in_block];
		tile.sync();
		const int total = build_neighbourhood<G>(tile, sc, cell_start, ix, iy, iz, nb);

		for(int t = 0; t < n_here; t++) {
			// ---- proposal (MC.c:312-324) -------------------------------------------------
			const uint4 x = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | 0u, sc.seed_lo, sc.seed_hi);
			const uint4 y = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | 1u, sc.seed_lo, sc.seed_hi);
			const int sp = first + (int) __umulhi(x.x, (uint32_t) n_here);
			const double dx = __dmul_rn(__dsub_rn(u32_to_unit(x.y), 0.5), sc.disp_max);
			const double dy = __dmul_rn(__dsub_rn(u32_to_unit(x.z), 0.5), sc.disp_max);
			const double dz = __dmul_rn(__dsub_rn(u32_to_unit(x.w), 0.5), sc.disp_max);
			const double theta = __dmul_rn(u32_to_unit(y.x), sc.theta_max);
			const double u_acc = __dmul_rn((double) ((((unsigned long long) y.y) << 21) | (((unsigned long long) y.z) >> 11)), 1.0 / 9007199254740992.0);
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
			a_tries++;

			const double4 p4 = pos[sp];
			const Vec3 p_old = {p4.x, p4.y, p4.z};
			const Vec3 p_new = {__dadd_rn(p4.x, dx), __dadd_rn(p4.y, dy), __dadd_rn(p4.z, dz)};
			int nx_, ny_, nz_;
			if(cell_index(sc, p_new.x, p_new.y, p_new.z, nx_, ny_, nz_) != cell) { // would leave its cell
				a_conf++;
				continue;
			}
			const double W = __ddiv_rn(1.,
					__dmul_rn(__dmul_rn(__dsub_rn(1., __ddiv_rn(fabs(dx), sc.cell_w[0])), __dsub_rn(1., __ddiv_rn(fabs(dy), sc.cell_w[1]))),
							__dsub_rn(1., __ddiv_rn(fabs(dz), sc.cell_w[2]))));
			const unsigned long long Wfix = (unsigned long long) __dmul_rn(W, 16777216.0);
			a_wt += Wfix;

			// axis (utils.c:127-131), Rodrigues matrix (utils.c:186-202), new = old . R (utils.c:172-184)
			const double ranh = __dmul_rn(2., __dsqrt_rn(__dsub_rn(1., ransq)));
			const double ex = __dmul_rn(ran1, ranh), ey = __dmul_rn(ran2, ranh), ez = __dsub_rn(1., __dmul_rn(2., ransq));
			double st, ct;
			sincos(theta, &st, &ct);
			const double uct = __dsub_rn(1., ct);
			double R[9];
			{
				const double xx = __dmul_rn(ex, ex), yy = __dmul_rn(ey, ey), zz = __dmul_rn(ez, ez);
				const double xy = __dmul_rn(__dmul_rn(ex, ey), uct), xz = __dmul_rn(__dmul_rn(ex, ez), uct), yz = __dmul_rn(__dmul_rn(ey, ez), uct);
				const double xs = __dmul_rn(ex, st), ys = __dmul_rn(ey, st), zs = __dmul_rn(ez, st);
				R[0] = __dadd_rn(xx, __dmul_rn(__dsub_rn(1., xx), ct));
				R[1] = __dsub_rn(xy, zs);
				R[2] = __dadd_rn(xz, ys);
				R[3] = __dadd_rn(xy, zs);
				R[4] = __dadd_rn(yy, __dmul_rn(__dsub_rn(1., yy), ct));
				R[5] = __dsub_rn(yz, xs);
				R[6] = __dsub_rn(xz, ys);
				R[7] = __dadd_rn(yz, xs);
				R[8] = __dadd_rn(zz, __dmul_rn(__dsub_rn(1., zz), ct));
			}
			double m_old[9], m_new[9], pp_old[12], pp_new[12];
#pragma unroll
			for(int k = 0; k < 9; k++) m_old[k] = ori[9ll * sp + k];
#pragma unroll
			for(int i = 0; i < 3; i++)
#pragma unroll
				for(int j = 0; j < 3; j++) m_new[3 * i + j] = dot3_rn(m_old[3 * i], m_old[3 * i + 1], m_old[3 * i + 2], R[j], R[3 + j], R[6 + j]);
			patches_from(sc.base, m_old, pp_old);
			patches_from(sc.base, m_new, pp_new);

			// ---- old / new pair state over the neighbourhood (MC.c:328-330) -----------------
			int b_old = 0, b_new = 0, o_new = 0;
			for(int j = lane; j < total; j += G) {
				const int sq = neighbour_slot(nb, j);
				if(sq == sp) continue;
				const double4 q4 = pos[sq];
				Vec3 d_old, d_new;
				const double d2_old = min_image(sc, p_old, q4, d_old);
				const double d2_new = min_image(sc, p_new, q4, d_new);
				if(d2_new < 1.) o_new = 1;
				if(d2_old >= 1. && d2_old < sc.sqr_rcut) b_old += patches_bond(sc, d_old, d2_old, pp_old, ori + 9ll * sq) ? 1 : 0;
				if(d2_new >= 1. && d2_new < sc.sqr_rcut) b_new += patches_bond(sc, d_new, d2_new, pp_new, ori + 9ll * sq) ? 1 : 0;
			}
			a_checks += 2ull * (unsigned long long) (total - 1);
			b_old = cg::reduce(tile, b_old, cg::plus<int>());
			b_new = cg::reduce(tile, b_new, cg::plus<int>());
			o_new = cg::reduce(tile, o_new, cg::bit_or<int>());

			// ---- Metropolis (MC.c:332-339) ----------------------------------------------------
			if(o_new) {
				a_ovl++;
				continue;
			}
			const int dE = b_old - b_new; // energy = -bonds
			if(dE < 0 || u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
				if(lane == 0) pos[sp] = make_double4(p_new.x, p_new.y, p_new.z, p4.w);
				for(int k = lane; k < 9; k += G) ori[9ll * sp + k] = m_new[k];
				a_dbonds += b_new - b_old;
				a_acc++;
				a_wa += Wfix;
				tile.sync(); // make the new state visible to the tile's next trial
			}
		}
	}

	if(lane == 0) {
		atomicAdd(&s_cnt.tries, a_tries);
		atomicAdd(&s_cnt.accepted, a_acc);
		atomicAdd(&s_cnt.confined, a_conf);
		atomicAdd(&s_cnt.overlapped, a_ovl);
		atomicAdd(&s_cnt.w_tries, a_wt);
		atomicAdd(&s_cnt.w_accepted, a_wa);
		atomicAdd(&s_cnt.pair_checks, a_checks);
		atomicAdd((unsigned long long *) &s_cnt.dbonds, (unsigned long long) a_dbonds);
	}
	__syncthreads();
	if(threadIdx.x == 0 && s_cnt.tries) {
		atomicAdd((unsigned long long *) &counters->tries, s_cnt.tries);
		atomicAdd((unsigned long long *) &counters->accepted, s_cnt.accepted);
		atomicAdd((unsigned long long *) &counters->confined, s_cnt.confined);
		atomicAdd((unsigned long long *) &counters->overlapped, s_cnt.overlapped);
		atomicAdd(&counters->w_tries, s_cnt.w_tries);
		atomicAdd(&counters->w_accepted, s_cnt.w_accepted);
		atomicAdd((unsigned long long *) &counters->pair_checks, s_cnt.pair_checks);
		atomicAdd((unsigned long long *) &counters->dbonds, (unsigned long long) s_cnt.dbonds);
	}
}

template<int G>
static int launch_pass_g(const SimConst &sc, int colour, const int *cell_start, double4 *pos, double *ori, Counters *counters, cudaStream_t st) {
	const int n_active = (sc.n_side[0] / 2) * (sc.n_side[1] / 2) * (sc.n_side[2] / 2);
	const int tiles_per_block = kSweepThreads / G;
	int blocks = (n_active + tiles_per_block - 1) / tiles_per_block;
	const int cap = 148 * 16 * 8;
	if(blocks > cap) blocks = cap;
	if(blocks < 1) blocks = 1;
	k_sweep_pass<G><<<blocks, kSweepThreads, 0, st>>>(sc, colour, cell_start, pos, ori, counters);
	return 1;
}

int launch_sweep_pass(const SimConst &sc, int colour, int lanes, const int *cell_start, double4 *pos, double *ori, Counters *counters, cudaStream_t st) {
	switch(lanes) {
	case 4:
		return launch_pass_g<4>(sc, colour, cell_start, pos, ori, counters, st);
	case 8:
		return launch_pass_g<8>(sc, colour, cell_start, pos, ori, counters, st);
	case 16:
		return launch_pass_g<16>(sc, colour, cell_start, pos, ori, counters, st);
	default:
		return launch_pass_g<32>(sc, colour, cell_start, pos, ori, counters, st);
	}
}

} // namespace pp
