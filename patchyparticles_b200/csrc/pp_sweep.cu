// pp_sweep.cu -- K2: one colour pass of the 8-colour checkerboard sweep.
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
// One tile of G lanes per active cell:
//   1. the (deduplicated) 27-cell neighbourhood is staged once in shared memory (positions + slots);
//   2. the proposals of up to G trials are generated in parallel, one per lane (Philox4x32-10 keyed
//      by the seed, counter (cell, trial, sweep, block); Marsaglia axis; Rodrigues matrix) and
//      parked in shared memory;
//   3. the trials run in order: the lanes share the 3x3 products and the patch vectors of the moved
//      particle (one element per lane, parked in shared memory), split the staged neighbours,
//      evaluate the old and new pair state in FP64 with the reference's operation order, reduce
//      bond counts / overlap across the tile, apply the Metropolis rule and write the accepted
//      state back to HBM and to the staged copy.
// A trajectory depends only on (seed, sweep, cell, trial): not on G, the block size, or the GPU.
#include "pp_kernels.cuh"

namespace pp {

constexpr int kStageCap = 64; // staged neighbours per cell; denser neighbourhoods read HBM/L2 directly

struct BlockCounters {
	unsigned long long tries, accepted, confined, overlapped, w_tries, w_accepted, pair_checks;
	long long dbonds;
};

struct Proposal {
	double R[9];
	double dx, dy, dz;
	double u_acc;
	int pick;
	int pad;
};

template<int G>
struct TileSmem {
	Neighbourhood nb;
	double x[kStageCap], y[kStageCap], z[kStageCap];
	int slot[kStageCap];
	double m_old[9], m_new[9], pp_old[12], pp_new[12];
	Proposal prop[G];
};

// proposal of trial t of a cell (MC.c:312-324, utils.c:115-131, 186-202); every number is a pure
// function of (seed, sweep, cell, t)
__device__ __forceinline__ void make_proposal(const SimConst &sc, int cell, int t, int n_here, Proposal &pr) {
	const uint4 x = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | 0u, sc.seed_lo, sc.seed_hi);
	const uint4 y = philox4x32((uint32_t) cell, (uint32_t) t, sc.sweep_lo, sc.sweep_hi8 | 1u, sc.seed_lo, sc.seed_hi);
	pr.pick = (int) __umulhi(x.x, (uint32_t) n_here);
	pr.dx = __dmul_rn(__dsub_rn(u32_to_unit(x.y), 0.5), sc.disp_max);
	pr.dy = __dmul_rn(__dsub_rn(u32_to_unit(x.z), 0.5), sc.disp_max);
	pr.dz = __dmul_rn(__dsub_rn(u32_to_unit(x.w), 0.5), sc.disp_max);
	const double theta = __dmul_rn(u32_to_unit(y.x), sc.theta_max);
	pr.u_acc = __dmul_rn((double) ((((unsigned long long) y.y) << 21) | (((unsigned long long) y.z) >> 11)), 1.0 / 9007199254740992.0);
	double ran1, ran2, ransq;
	for(uint32_t blk = 2;; blk++) { // Marsaglia pairs
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
}

// weight of a non-confined trial in the acceptance estimator: 1 / P(this displacement stays in a cell)
__device__ __forceinline__ unsigned long long trial_weight(const SimConst &sc, double dx, double dy, double dz) {
	const double keep = __dmul_rn(__dmul_rn(__dsub_rn(1., __dmul_rn(fabs(dx), sc.inv_w[0])), __dsub_rn(1., __dmul_rn(fabs(dy), sc.inv_w[1]))),
			__dsub_rn(1., __dmul_rn(fabs(dz), sc.inv_w[2])));
	return (unsigned long long) __dmul_rn(__ddiv_rn(1., keep), 16777216.0);
}

template<int G>
__host__ __device__ constexpr int sweep_threads() {
	return G == 4 ? 32 : 64;
}

template<int G>
__global__ void __launch_bounds__(sweep_threads<G>(), 512 / sweep_threads<G>()) k_sweep_pass(const __grid_constant__ SimConst sc, int colour, const int *__restrict__ cell_start,
		double4 *pos, double *ori, Counters *__restrict__ counters) {
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

	const int ox = colour & 1, oy = (colour >> 1) & 1, oz = (colour >> 2) & 1;
	const int hx = sc.n_side[0] >> 1, hy = sc.n_side[1] >> 1, hz = sc.n_side[2] >> 1;
	const int n_active = hx * hy * hz;
	// per-tile tallies (kept by every lane, flushed by lane 0)
	unsigned int a_tries = 0, a_acc = 0, a_conf = 0, a_ovl = 0;
	unsigned long long a_wt = 0, a_wa = 0, a_checks = 0;
	int a_dbonds = 0;

	for(int a = blockIdx.x * kTiles + (int) (threadIdx.x / G); a < n_active; a += gridDim.x * kTiles) {
		const int ax = a / (hy * hz), ay = (a - ax * hy * hz) / hz, az = a - ax * hy * hz - ay * hz;
		const int ix = 2 * ax + ox, iy = 2 * ay + oy, iz = 2 * az + oz;
		const int cell = (ix * sc.n_side[1] + iy) * sc.n_side[2] + iz;
		const int first = cell_start[cell];
		const int n_here = cell_start[cell + 1] - first;
		if(n_here == 0) continue;
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

		for(int t0 = 0; t0 < n_here; t0 += G) {
			tile.sync();
			if(t0 + lane < n_here) make_proposal(sc, cell, t0 + lane, n_here, ts.prop[lane]);
			tile.sync();
			const int n_batch = min(G, n_here - t0);
			for(int tt = 0; tt < n_batch; tt++) {
				const Proposal &pr = ts.prop[tt];
				a_tries++;
				const int sp = first + pr.pick;
				const int j_self = home_base + pr.pick;
				Vec3 p_old;
				if(staged) p_old = {ts.x[j_self], ts.y[j_self], ts.z[j_self]};
				else {
					const double4 p4 = pos[sp];
					p_old = {p4.x, p4.y, p4.z};
				}
				const Vec3 p_new = {__dadd_rn(p_old.x, pr.dx), __dadd_rn(p_old.y, pr.dy), __dadd_rn(p_old.z, pr.dz)};
				int cx, cy, cz;
				if(cell_index(sc, p_new.x, p_new.y, p_new.z, cx, cy, cz) != cell) { // would leave its cell
					a_conf++;
					continue;
				}
				const unsigned long long Wfix = trial_weight(sc, pr.dx, pr.dy, pr.dz);
				a_wt += Wfix;

				// ---- new orientation and both sets of patches, one element per lane -----------
				for(int k = lane; k < 9; k += G) ts.m_old[k] = ori[9ll * sp + k];
				tile.sync();
				for(int k = lane; k < 9; k += G) {
					const int i = k / 3, j = k - 3 * i; // new = old . R (utils.c:172-184)
					ts.m_new[k] = dot3_rn(ts.m_old[3 * i], ts.m_old[3 * i + 1], ts.m_old[3 * i + 2], pr.R[j], pr.R[3 + j], pr.R[6 + j]);
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

				// ---- old / new pair state over the neighbourhood (MC.c:328-330) -------------
				int b_old = 0, b_new = 0, o_new = 0;
				for(int j = lane; j < total; j += G) {
					int sq;
					double qx, qy, qz;
					if(staged) {
						if(j == j_self) continue;
						sq = ts.slot[j];
						qx = ts.x[j];
						qy = ts.y[j];
						qz = ts.z[j];
					}
					else {
						sq = neighbour_slot(ts.nb, j);
						if(sq == sp) continue;
						const double4 q4 = pos[sq];
						qx = q4.x;
						qy = q4.y;
						qz = q4.z;
					}
					Vec3 d_old, d_new;
					const double d2_old = min_image(sc, p_old, qx, qy, qz, d_old);
					const double d2_new = min_image(sc, p_new, qx, qy, qz, d_new);
					if(d2_new < 1.) o_new = 1;
					if(d2_old >= 1. && d2_old < sc.sqr_rcut) b_old += patches_bond(sc, d_old, d2_old, ts.pp_old, ori + 9ll * sq) ? 1 : 0;
					if(d2_new >= 1. && d2_new < sc.sqr_rcut) b_new += patches_bond(sc, d_new, d2_new, ts.pp_new, ori + 9ll * sq) ? 1 : 0;
				}
				a_checks += 2ull * (unsigned long long) (total - 1);
				// one packed reduction: bonds old | bonds new | overlap
				const unsigned packed = cg::reduce(tile, (unsigned) b_old | ((unsigned) b_new << 10) | ((unsigned) o_new << 20), cg::plus<unsigned>());
				b_old = (int) (packed & 1023u);
				b_new = (int) ((packed >> 10) & 1023u);
				o_new = (int) (packed >> 20);

				// ---- Metropolis (MC.c:332-339) ------------------------------------------------
				if(o_new) {
					a_ovl++;
					continue;
				}
				const int dE = b_old - b_new; // energy = -bonds
				if(dE < 0 || pr.u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
					if(lane == 0) {
						double *p = reinterpret_cast<double *>(pos + sp); // x, y, z; the id in .w stays
						p[0] = p_new.x;
						p[1] = p_new.y;
						p[2] = p_new.z;
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
					tile.sync(); // make the new state visible to the tile's next trial
				}
			}
		}
	}

	if(lane == 0 && a_tries) {
		atomicAdd(&s_cnt.tries, (unsigned long long) a_tries);
		atomicAdd(&s_cnt.accepted, (unsigned long long) a_acc);
		atomicAdd(&s_cnt.confined, (unsigned long long) a_conf);
		atomicAdd(&s_cnt.overlapped, (unsigned long long) a_ovl);
		atomicAdd(&s_cnt.w_tries, a_wt);
		atomicAdd(&s_cnt.w_accepted, a_wa);
		atomicAdd(&s_cnt.pair_checks, a_checks);
		atomicAdd((unsigned long long *) &s_cnt.dbonds, (unsigned long long) (long long) a_dbonds);
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
	constexpr int kThreads = sweep_threads<G>();
	const int n_active = (sc.n_side[0] / 2) * (sc.n_side[1] / 2) * (sc.n_side[2] / 2);
	const int tiles_per_block = kThreads / G;
	int blocks = (n_active + tiles_per_block - 1) / tiles_per_block;
	const int cap = 148 * 32 * 8;
	if(blocks > cap) blocks = cap;
	if(blocks < 1) blocks = 1;
	k_sweep_pass<G><<<blocks, kThreads, 0, st>>>(sc, colour, cell_start, pos, ori, counters);
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
