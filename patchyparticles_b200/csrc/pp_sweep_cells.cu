// pp_sweep_cells.cu -- K2b, G lanes (1, 2, 4 or 8) per non-empty cell of the current colour.
//
// Same algorithm, same random streams and bit-identical trajectories as the tile kernel in
// pp_sweep.cu (both are checked against the serial restatement in oracle/pp_oracle.c); what changes
// is the mapping onto the machine and the arithmetic of the screens.
//   G = 1  a thread walks its cell's trials on its own, so nothing is computed redundantly, shuffled or
//          reduced.  The price is a serial chain per cell, so the cell build routes the sparsely populated
//          cells here (at most M members, a neighbourhood that fits the staging area: >95 % of the cells of
//          a fluid) ...
//   G = 8  ... and the crowded ones to the 8-lane configuration, whose lanes split the nine columns of the
//          neighbourhood and which also stages the cell's proposals and the FP64 state of its own particles
//          (shortest chain per trial).  One launch per colour pass runs the configurations side by side.
// The work list hands consecutive groups cells of equal occupancy, which keeps the lanes of a warp in the same
// trial.  No TMA: the staging is a gather of 16-byte items from nine short runs per cell, not a tile; cp.async
// (LDGSTS) is the fire-and-forget primitive that fits it.
//
//   per cell   the shadows (pp_common.cuh: box fractions of x, y, z as 32-bit integers + the slot) of the 27-cell
//              neighbourhood are brought into shared memory with cp.async.  A column of three cells along z is one
//              contiguous run of slots (two at the box edge), so staging is nine run copies -- all in flight at once,
//              one wait -- and the staged entry needs no tag: shadow differences are minimum-image separations;
//   per trial  1. FP32 distance screen of the staged neighbours for the old and the new position, four
//                 entries per iteration (the list is padded with far-away entries): a new position closer than
//                 1 - 2e-3 (squared) to a neighbour is rejected on the spot, a pair is a candidate when its
//                 squared distance is below r_cut^2 + 2e-3 (the shadow arithmetic is good to 1e-5, so no pair the
//                 FP64 test would put inside the shell, or flag as overlapping, is ever dropped);
//              2. FP32 orientation screen of the candidates: most pairs in the shell cannot bond because no patch of the
//                 moved particle points at them (own matrix only); for the few that are left, the same screen from the
//                 neighbour's matrix -- a bond needs a patch of each pointing at the other (MC.c:249-259);
//              3. FP64, reference operation order, on what is left: overlap of the new position
//                 (MC.c:240), shell test (MC.c:241), orientation tests (MC.c:249-261);
//              4. Metropolis rule (MC.c:332), write-back of position, orientation, shadow.
#include "pp_trial_serial.cuh"

#include <cuda_pipeline.h>

namespace pp {

// one warp per block: 12 blocks per SM at 168 registers and 18 KB of shared memory.  (Two warps per block, six blocks per SM, is the
// same at every size but 256k, where the finer blocks gain 13 %: a block's slot is free again when ITS cells are done.)
constexpr int kCellThreads = 32;
constexpr int kCellBlocksPerSM = 384 / kCellThreads;
constexpr float kScreenBand = 2e-3f;         // safety band of the FP32 screens (error of the shadow arithmetic < 1e-5)
constexpr unsigned kNearOld = 0x40u, kNearNew = 0x80u, kIdxMask = 0x3fu;

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct CellBlockCounters {
	unsigned int tries, accepted, confined, overlapped;
	int dbonds;
	unsigned long long w_tries, w_accepted, pair_checks;
};

// staging capacity of one lane (neighbours per lane, a multiple of 4); the cell build routes a sparsely populated cell
// whose neighbourhood exceeds lane_cap(1) to the 8-lane configuration, a cell that does not fit there takes the serial path
__host__ __device__ constexpr int lane_cap(int G) { return G == 1 ? kCellsNbCap : (G == 2 ? 24 : (G == 4 ? 20 : 16)); }
static_assert(kCellsNbCap % 4 == 0 && kCellsNbCap <= 64, "candidate entries index the staging area with 6 bits; the screen loop takes four entries at a time");

// Latency configuration (G >= 4): the proposals of the cell's trials and the FP64 state of its own particles are
// staged too, so that a trial of a crowded cell does not wait for global memory between its steps.
__host__ __device__ constexpr int max_own(int G) { return G >= 8 ? 8 : 4; }
template<int kMaxOwn>
struct OwnStage {
	Proposal pr[kMaxOwn];
	double4 pos[kMaxOwn];
	double ori[kMaxOwn * 9];
};

// shared memory of one block (kCellThreads threads) running G lanes per cell
template<int G>
struct CellSmem {
	uint4 stage[lane_cap(G)][kCellThreads];        // [entry][thread]: shadow of a staged neighbour
	unsigned char cand[lane_cap(G)][kCellThreads];
	CellBlockCounters cnt;
	OwnStage<max_own(G)> own[G >= 4 ? kCellThreads / G : 1];
};

// G lanes per cell.  Lane l stages and screens the columns c = l, l + G, ... of the 3 x 3 columns of the
// neighbourhood and runs the FP64 tests of its own candidates; the lanes meet three times per trial (overlap
// flag after the distance screen, bond counts after the FP64 tests, via shuffles inside the group).  G = 1 is
// the throughput configuration; G > 1 shortens the serial chain of a cell when a pass is too small to
// fill the machine.
template<int G>
__device__ __forceinline__ void sweep_cells_body(CellSmem<G> &sm, const SimConst &sc, const int w_begin, const int w_end, const int block, const int n_blocks,
		const int *__restrict__ cell_start, const int *__restrict__ work, const Proposal *__restrict__ proposals, double4 *pos, double *ori, uint4 *posf,
		Counters *__restrict__ counters, const HaloMirror &hm) {
	constexpr int kCap = lane_cap(G);
	constexpr bool kOwn = G >= 4;
	constexpr int kMaxOwn = max_own(G);
	constexpr int kGroups = kCellThreads / G;
	constexpr int kCols = (9 + G - 1) / G;  // columns of the neighbourhood per lane
	constexpr int kOwner = 4 % G; // the lane that stages the home column (c = 4) holds the moved particle's entry and does the book-keeping
	uint4 (&s_stage)[kCap][kCellThreads] = sm.stage;
	unsigned char (&s_cand)[kCap][kCellThreads] = sm.cand;
	CellBlockCounters &s_cnt = sm.cnt;
	OwnStage<kMaxOwn> *s_own = sm.own;
	const int tid = threadIdx.x;
	if((long long) block * kGroups >= w_end - w_begin) return; // the grid is sized for the worst case: nothing left for this block
	const int hm_off = mirror_offset(hm);
	const int gl = tid & (G - 1); // lane inside the group
	OwnStage<kMaxOwn> &os = s_own[kOwn ? tid / G : 0];
	const unsigned gmask = (G == 1) ? 0u : (((1u << G) - 1u) << ((tid & 31) & ~(G - 1)));
	const int owner_lane = ((tid & 31) & ~(G - 1)) + kOwner; // warp lane of the group's owner
	const bool is_owner = gl == kOwner;
	if(tid == 0) s_cnt = CellBlockCounters{0, 0, 0, 0, 0, 0, 0, 0};
	__syncthreads();

	const int ny = sc.n_side[1], nz = sc.n_side[2], nyz = ny * nz;
	const float ux = (float) (sc.box[0] / 4294967296.), uy = (float) (sc.box[1] / 4294967296.), uz = (float) (sc.box[2] / 4294967296.); // shadow units
	// box fraction (x) of the cell origins: layer gx starts at shift / box + gx / nx
	const float frac_w = 4294967296.f / (float) sc.n_side[0];
	const unsigned shift_fx = box_fraction(sc.shift[0], sc.inv_box[0]);
	const float thr = (float) sc.sqr_rcut + kScreenBand;
	const float core_lo = 1.f - kScreenBand, core_hi = 1.f + kScreenBand; // certain overlap below, certain clearance above
	// (M^T d . b_i)^2 > cosmax^2 d^2 / base_z^2 with a relative band; -1 disables the screen (every pair in range goes to FP64)
	const float point_thr = sc.fast_patch ? (float) (sc.cos2_lo * sc.inv_b2) * (1.f - kScreenBand) : -1.f;
	Tally tl = {0, 0, 0, 0, 0, 0, 0, 0};

	for(int wi = w_begin + block * kGroups + tid / G; wi < w_end; wi += n_blocks * kGroups) {
		const int cell = work[wi];
		const int ix = cell / nyz, iy = (cell - ix * nyz) / nz, iz = cell - ix * nyz - iy * nz;
		// a column of the neighbourhood (three cells along z) is one run of slots, [a0, a1), plus at the two ends of the
		// box the wrapped cell zw; the loads of all the lane's columns are issued together
		const int z_lo = iz > 0 ? iz - 1 : 0, z_hi = iz < nz - 1 ? iz + 1 : nz - 1;
		const bool edge = iz == 0 || iz == nz - 1;
		const int zw = iz == 0 ? nz - 1 : 0;
		int a0[kCols], a1[kCols];
#pragma unroll
		for(int k = 0; k < kCols; k++) {
			const int c = gl + k * G;
			a0[k] = a1[k] = 0;
			if(c < 9) {
				int cx = ix + c / 3 - 1;
				if(!sc.slab) cx = wrap_cell(cx, sc.nxl); // (slab: an owned layer always has both x neighbours in the held range)
				const int *cs = cell_start + (cx * ny + wrap_cell(iy + c % 3 - 1, ny)) * nz;
				a0[k] = cs[z_lo];
				a1[k] = cs[z_hi + 1];
			}
		}
		const int first = cell_start[cell];
		const int n_here = cell_start[cell + 1] - first;
		const bool own_staged = kOwn && n_here <= kMaxOwn;
		if(own_staged) { // the cell's proposals and its particles' FP64 state: staged with the neighbourhood
			__syncwarp(gmask); // the group is done with the previous cell
			for(int i = gl; i < 7 * n_here; i += G)
				__pipeline_memcpy_async(reinterpret_cast<char *>(os.pr) + 16 * i, reinterpret_cast<const char *>(proposals + first) + 16 * i, 16);
			for(int i = gl; i < 2 * n_here; i += G)
				__pipeline_memcpy_async(reinterpret_cast<char *>(os.pos) + 16 * i, reinterpret_cast<const char *>(pos + first) + 16 * i, 16);
			for(int i = gl; i < 9 * n_here; i += G) __pipeline_memcpy_async(os.ori + i, ori + 9ll * first + i, 8);
		}
		else
			for(int t = gl; t < n_here; t += G) { // on their way to L2 while the neighbourhood is staged
				prefetch_l2(proposals + first + t);
				prefetch_l2(pos + first + t);
				prefetch_l2(ori + 9ll * (first + t));
				prefetch_l2(ori + 9ll * (first + t) + 8);
			}

		// ---- stage this lane's columns: run copies, all in flight, one wait ----------------------------------
		int total = 0, self_base = 0;
#pragma unroll
		for(int k = 0; k < kCols; k++) {
			const int c = gl + k * G;
			if(c == 4) self_base = total + (first - a0[k]);
			for(int s = a0[k]; s < a1[k]; s++, total++)
				if(total < kCap) {
					__pipeline_memcpy_async(&s_stage[total][tid], posf + s, 16);
					if(kOwn) { // FP64 state of a possible bonding partner: towards L2
						prefetch_l2(pos + s);
						prefetch_l2(ori + 9ll * s);
						prefetch_l2(ori + 9ll * s + 8);
					}
				}
		}
		if(edge) { // first / last cell along z: the periodic neighbour sits at the other end of the column
#pragma unroll 1
			for(int c = gl; c < 9; c += G) {
				int cx = ix + c / 3 - 1;
				if(!sc.slab) cx = wrap_cell(cx, sc.nxl);
				const int *cs = cell_start + (cx * ny + wrap_cell(iy + c % 3 - 1, ny)) * nz;
				const int e = cs[zw + 1];
				for(int s = cs[zw]; s < e; s++, total++)
					if(total < kCap) __pipeline_memcpy_async(&s_stage[total][tid], posf + s, 16);
			}
		}
		__pipeline_commit();
		bool staged = total <= kCap;
		if(G > 1) {
#pragma unroll
			for(int o = G / 2; o > 0; o >>= 1) staged &= (bool) __shfl_xor_sync(gmask, (int) staged, o);
		}
		// a far-away entry (half a box along x from the cell's centre): pads the list to a multiple of four and stands in
		// for the moved particle itself
		int gx = ix + sc.x_lo;
		gx -= (gx >= sc.n_side[0]) ? sc.n_side[0] : 0;
		const uint4 far = make_uint4(shift_fx + (unsigned) (((float) gx + 0.5f) * frac_w) + 0x80000000u, 0u, 0u, 0xffffffffu);
		const int total4 = (total + 3) & ~3;
		if(staged)
			for(int j = total; j < total4; j++) s_stage[j][tid] = far;
		__pipeline_wait_prior(0);
		if(kOwn) __syncwarp(gmask); // every lane's copies have landed

		for(int t = 0; t < n_here; t++) {
			const Proposal *pr = own_staged ? &os.pr[t] : proposals + first + t;
			const int pick = pr->pick;
			const int sp = first + pick;
			const double4 *own_pos = own_staged ? &os.pos[pick] : pos + sp;
			const double *own_ori = own_staged ? os.ori + 9 * pick : ori + 9ll * sp;
			if(is_owner) tl.tries++;
			if(!staged) { // neighbourhood larger than the staging area (dense packing): FP64 walk of the 27 cells, shared by the group's lanes
				if(G == 1) trial_serial(sc, cell_start, pos, ori, posf, proposals + first + t, cell, ix, iy, iz, sp, tl, hm, hm_off);
				else trial_group<G>(sc, cell_start, pos, ori, posf, proposals + first + t, cell, ix, iy, iz, sp, tl, hm, hm_off, gl, gmask, is_owner);
				continue;
			}
			const double dx = pr->dx, dy = pr->dy, dz = pr->dz;
			const double4 p4 = *own_pos;
			const Vec3 p_old = {p4.x, p4.y, p4.z};
			const Vec3 p_new = {__dadd_rn(p4.x, dx), __dadd_rn(p4.y, dy), __dadd_rn(p4.z, dz)};
			int cx, cy, cz;
			if(cell_index(sc, p_new.x, p_new.y, p_new.z, cx, cy, cz) != cell) { // would leave its cell
				if(is_owner) tl.conf++;
				continue;
			}
			const unsigned long long Wfix = (unsigned long long) pr->wfix;
			if(is_owner) tl.wt += Wfix;
			const int j_self = self_base + pick; // meaningful on the owner lane only
			uint4 self = s_stage[is_owner ? j_self : 0][tid];
			if(G > 1) {
				self.x = __shfl_sync(gmask, self.x, owner_lane);
				self.y = __shfl_sync(gmask, self.y, owner_lane);
				self.z = __shfl_sync(gmask, self.z, owner_lane);
			}
			const float fdx = (float) dx, fdy = (float) dy, fdz = (float) dz;

			// ---- 1. FP32 distance screen over the staged neighbourhood ---------------------
			if(is_owner) s_stage[j_self][tid] = far; // the moved particle itself: out of everybody's range
			int n_cand = 0;
			float d2n_min = 1e30f;
			for(int j0 = 0; j0 < total4; j0 += 4) {
#pragma unroll
				for(int u = 0; u < 4; u++) {
					const uint4 q = s_stage[j0 + u][tid];
					const float ax = shadow_sep(q.x, self.x, ux), ay = shadow_sep(q.y, self.y, uy), az = shadow_sep(q.z, self.z, uz);
					const float bx = ax - fdx, by = ay - fdy, bz = az - fdz;
					const float d2o = fmaf(az, az, fmaf(ay, ay, ax * ax));
					const float d2n = fmaf(bz, bz, fmaf(by, by, bx * bx));
					d2n_min = fminf(d2n_min, d2n);
					if(fminf(d2o, d2n) < thr) s_cand[n_cand++][tid] = (unsigned char) (j0 + u); // n_cand <= total <= kCap
				}
			}
			tl.checks += 2ull * (unsigned long long) (total - (is_owner ? 1 : 0));
			bool ovl = d2n_min < core_lo;
			if(G > 1) ovl = __any_sync(gmask, ovl);
			if(ovl) { // overlapping beyond any FP32 doubt: rejected whatever the bonds (MC.c:296-299, 332)
				if(is_owner) {
					s_stage[j_self][tid] = self;
					tl.ovl++;
				}
				continue;
			}

			// ---- 2. FP32 orientation screen of the pairs in range -------------------------------
			// A bond needs one of the moved particle's patches to point at the neighbour (MC.c:249-255).
			// With cosmax = 0.98 that is true for ~4 % of the pairs in the shell; the others are dismissed
			// here from the FP32 separation and the FP32 copy of the particle's own matrix, through a
			// relative band on the squared cosine.  What survives -- pointing or doubtful patches,
			// separations within the band of the hard core -- goes to FP64.
			double m[9], mn[9];
#pragma unroll
			for(int k = 0; k < 9; k++) m[k] = own_ori[k];
#pragma unroll
			for(int i = 0; i < 3; i++)
#pragma unroll
				for(int j = 0; j < 3; j++) mn[3 * i + j] = dot3_rn(m[3 * i], m[3 * i + 1], m[3 * i + 2], pr->R[j], pr->R[3 + j], pr->R[6 + j]);
			int n_exact = 0;
			{
				float mf[9], mnf[9];
#pragma unroll
				for(int k = 0; k < 9; k++) {
					mf[k] = (float) m[k];
					mnf[k] = (float) mn[k];
				}
				for(int c0 = 0; c0 < n_cand; c0++) {
					const unsigned j = s_cand[c0][tid];
					const uint4 q = s_stage[j][tid];
					const float ax = shadow_sep(q.x, self.x, ux), ay = shadow_sep(q.y, self.y, uy), az = shadow_sep(q.z, self.z, uz);
					const float bx = ax - fdx, by = ay - fdy, bz = az - fdz;
					const float d2o = fmaf(az, az, fmaf(ay, ay, ax * ax));
					const float d2n = fmaf(bz, bz, fmaf(by, by, bx * bx));
					const bool core_new = d2n < core_hi; // (d2n < thr as well) the hard core is decided in FP64 whatever the patches do
					bool pt_new = d2n < thr && !core_new && may_point(mnf, bx, by, bz, d2n, point_thr);
					bool pt_old = d2o < thr && may_point(mf, ax, ay, az, d2o, point_thr);
					if((pt_new || pt_old) && point_thr >= 0.f) {
						// ... and one of the neighbour's patches to point back (MC.c:256-259): the same screen from the neighbour's
						// matrix, which the FP64 test would fetch anyway; 96 % of what came this far ends here
						float mq[9];
						const double *oq = ori + 9ll * (int) q.w;
#pragma unroll
						for(int k = 0; k < 9; k++) mq[k] = (float) oq[k];
						pt_new = pt_new && may_point(mq, -bx, -by, -bz, d2n, point_thr);
						pt_old = pt_old && may_point(mq, -ax, -ay, -az, d2o, point_thr);
					}
					const unsigned keep = ((core_new || pt_new) ? kNearNew : 0u) | (pt_old ? kNearOld : 0u);
					if(keep) s_cand[n_exact++][tid] = (unsigned char) (j | keep); // in place: n_exact <= c0
				}
			}

			// ---- 3. FP64, reference operation order, on what is left ---------------------------
			int b_old = 0, b_new = 0;
#pragma unroll 1
			for(int c0 = 0; c0 < n_exact && !ovl; c0++) {
				const unsigned e = s_cand[c0][tid];
				const int sq = (int) s_stage[e & kIdxMask][tid].w; // the shadow carries its slot
				const double4 q4 = pos[sq];
				if(e & kNearNew) {
					Vec3 d;
					const double d2 = min_image(sc, p_new, q4, d);
					if(d2 < 1.) ovl = true;
					else if(d2 < sc.sqr_rcut) b_new += patches_bond_m(sc, d, d2, mn, ori + 9ll * sq) ? 1 : 0;
				}
				if(e & kNearOld) {
					Vec3 d;
					const double d2 = min_image(sc, p_old, q4, d);
					if(d2 >= 1. && d2 < sc.sqr_rcut) b_old += patches_bond_m(sc, d, d2, m, ori + 9ll * sq) ? 1 : 0;
				}
			}
			if(G > 1) { // the group's verdict: any overlap, total bonds of the old and of the new state
				unsigned packed = (unsigned) b_old | ((unsigned) b_new << 8) | (ovl ? 1u << 16 : 0u);
#pragma unroll
				for(int o = G / 2; o > 0; o >>= 1) packed += __shfl_xor_sync(gmask, packed, o);
				b_old = (int) (packed & 0xffu);
				b_new = (int) ((packed >> 8) & 0xffu);
				ovl = (packed >> 16) != 0u;
			}
			if(ovl) { // hard-core overlap: rejected whatever the bonds (MC.c:296-299, 332)
				if(is_owner) {
					s_stage[j_self][tid] = self;
					tl.ovl++;
				}
				continue;
			}

			// ---- 4. Metropolis (MC.c:332-339) ------------------------------------------------
			const int dE = b_old - b_new; // energy = -bonds
			if(kOwn) __syncwarp(gmask); // every lane has read the particle's staged state (the shuffles above order execution, not memory)
			if(is_owner) {
				if(dE < 0 || pr->u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
					double *p = reinterpret_cast<double *>(pos + sp); // x, y, z; the id in .w stays
					p[0] = p_new.x;
					p[1] = p_new.y;
					p[2] = p_new.z;
					const uint4 moved = box_shadow(sc, p_new.x, p_new.y, p_new.z, sp);
					store_shadow_xyz(posf + sp, moved);
					s_stage[j_self][tid] = moved;
#pragma unroll
					for(int k = 0; k < 9; k++) ori[9ll * sp + k] = mn[k];
					mirror_state(hm, hm_off, sp, p_new, moved, mn);
					if(own_staged) {
						os.pos[pick] = make_double4(p_new.x, p_new.y, p_new.z, p4.w);
#pragma unroll
						for(int k = 0; k < 9; k++) os.ori[9 * pick + k] = mn[k];
					}
					tl.dbonds += b_new - b_old;
					tl.acc++;
					tl.wa += Wfix;
				}
				else s_stage[j_self][tid] = self;
			}
			if(G > 1) __syncwarp(gmask); // the other lanes read the moved particle's new state in the next trials
		}
	}

	// block tallies: warp shuffle reduce, then shared atomics, then one set of global atomics per block
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		tl.tries += __shfl_xor_sync(0xffffffffu, tl.tries, o);
		tl.acc += __shfl_xor_sync(0xffffffffu, tl.acc, o);
		tl.conf += __shfl_xor_sync(0xffffffffu, tl.conf, o);
		tl.ovl += __shfl_xor_sync(0xffffffffu, tl.ovl, o);
		tl.dbonds += __shfl_xor_sync(0xffffffffu, tl.dbonds, o);
		tl.wt += __shfl_xor_sync(0xffffffffu, tl.wt, o);
		tl.wa += __shfl_xor_sync(0xffffffffu, tl.wa, o);
		tl.checks += __shfl_xor_sync(0xffffffffu, tl.checks, o);
	}
	if((tid & 31) == 0 && tl.tries) {
		atomicAdd(&s_cnt.tries, tl.tries);
		atomicAdd(&s_cnt.accepted, tl.acc);
		atomicAdd(&s_cnt.confined, tl.conf);
		atomicAdd(&s_cnt.overlapped, tl.ovl);
		atomicAdd(&s_cnt.dbonds, tl.dbonds);
		atomicAdd(&s_cnt.w_tries, tl.wt);
		atomicAdd(&s_cnt.w_accepted, tl.wa);
		atomicAdd(&s_cnt.pair_checks, tl.checks);
	}
	__syncthreads();
	if(tid == 0 && s_cnt.tries) {
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

// one configuration alone over one range of the work list
template<int G>
__global__ void __launch_bounds__(kCellThreads, kCellBlocksPerSM) k_sweep_cells(const __grid_constant__ SimConst sc, const int *__restrict__ range,
		const int *__restrict__ cell_start, const int *__restrict__ work, const Proposal *__restrict__ proposals, double4 *pos, double *ori,
		uint4 *posf, Counters *__restrict__ counters, const __grid_constant__ HaloMirror hm) {
	__shared__ __align__(16) CellSmem<G> sm;
	sweep_cells_body<G>(sm, sc, range[0], range[1], (int) blockIdx.x, (int) gridDim.x, cell_start, work, proposals, pos, ori, posf, counters, hm);
}

// One launch per colour pass: the three ranges of the colour's work list (occupancy descending) run side by side in one
// grid -- the most crowded cells on 8 lanes per cell, the middle range on 4, the sparsely populated rest on one thread
// per cell.  The first blocks take the crowded cells (the longest serial chains start first); each configuration gets
// the blocks its range needs, capped by its share of the grid, and strides over its range.
union FusedSmem {
	CellSmem<1> c1;
	CellSmem<4> c4;
	CellSmem<8> c8;
	__device__ FusedSmem() {}
};

__global__ void __launch_bounds__(kCellThreads, kCellBlocksPerSM) k_sweep_fused(const __grid_constant__ SimConst sc, const int *__restrict__ ranges /* 6 ints */,
		const int *__restrict__ cell_start, const int *__restrict__ work, const Proposal *__restrict__ proposals, double4 *pos, double *ori,
		uint4 *posf, Counters *__restrict__ counters, const __grid_constant__ HaloMirror hm) {
	__shared__ __align__(16) unsigned char raw[sizeof(FusedSmem)];
	const int b8 = ranges[0], e8 = ranges[1], b4 = ranges[2], e4 = ranges[3], b1 = ranges[4], e1 = ranges[5];
	const int grid = (int) gridDim.x;
	constexpr int k8 = kCellThreads / 8, k4 = kCellThreads / 4, k1 = kCellThreads; // cells per block of each configuration
	int n8 = (e8 - b8 + k8 - 1) / k8, n4 = (e4 - b4 + k4 - 1) / k4, n1 = (e1 - b1 + k1 - 1) / k1;
	// shares of a grid that is smaller than the work: crowded cells keep at most a quarter each, the rest goes to the sparse range
	n8 = min(n8, max(grid / 4, 1));
	n4 = min(n4, max(grid / 4, 1));
	n1 = min(n1, max(grid - n8 - n4, 1));
	const int b = (int) blockIdx.x;
	if(b < n8) sweep_cells_body<8>(*reinterpret_cast<CellSmem<8> *>(raw), sc, b8, e8, b, n8, cell_start, work, proposals, pos, ori, posf, counters, hm);
	else if(b < n8 + n4)
		sweep_cells_body<4>(*reinterpret_cast<CellSmem<4> *>(raw), sc, b4, e4, b - n8, n4, cell_start, work, proposals, pos, ori, posf, counters, hm);
	else if(b < n8 + n4 + n1)
		sweep_cells_body<1>(*reinterpret_cast<CellSmem<1> *>(raw), sc, b1, e1, b - n8 - n4, n1, cell_start, work, proposals, pos, ori, posf, counters, hm);
}

template<int G>
static void launch_cells_g(const SimConst &sc, const int *range, long long max_items, const CellBuffers &cb, const Proposal *proposals, double4 *pos,
		double *ori, Counters *counters, const HaloMirror &hm, cudaStream_t st) {
	constexpr int groups = kCellThreads / G;
	long long blocks = (max_items + groups - 1) / groups;
	const long long cap = 148ll * 4 * 8; // a few waves of resident blocks; groups stride over the occupancy-sorted list
	if(blocks > cap) blocks = cap;
	k_sweep_cells<G><<<(unsigned) blocks, kCellThreads, 0, st>>>(sc, range, cb.cell_start, cb.work, proposals, pos, ori, cb.posf, counters, hm);
}

// range: device pointer to {first work item, one past the last work item}; lanes: 1, 2, 4 or 8 lanes per cell
int launch_sweep_cells(const SimConst &sc, const int *range, long long max_items, int lanes, const CellBuffers &cb, const Proposal *proposals,
		double4 *pos, double *ori, Counters *counters, const HaloMirror &hm, cudaStream_t st) {
	if(max_items < 1) max_items = 1;
	switch(lanes) {
	case 8: launch_cells_g<8>(sc, range, max_items, cb, proposals, pos, ori, counters, hm, st); break;
	case 4: launch_cells_g<4>(sc, range, max_items, cb, proposals, pos, ori, counters, hm, st); break;
	case 2: launch_cells_g<2>(sc, range, max_items, cb, proposals, pos, ori, counters, hm, st); break;
	default: launch_cells_g<1>(sc, range, max_items, cb, proposals, pos, ori, counters, hm, st); break;
	}
	return 1;
}

// ranges: device pointer to the colour's six range bounds {8-lane begin, end, 4-lane begin, end, 1-lane begin, end}
int launch_sweep_fused(const SimConst &sc, const int *ranges, long long max_items, const CellBuffers &cb, const Proposal *proposals, double4 *pos,
		double *ori, Counters *counters, const HaloMirror &hm, cudaStream_t st) {
	if(max_items < 1) max_items = 1;
	// worst case: every active cell non-empty and sparse needs max_items / 64 blocks, every one crowded max_items / 8;
	// typical fluids need about max_items / 40.  The kernel strides when the grid is smaller than the work.
	long long blocks = (max_items + kCellThreads / 4 - 1) / (kCellThreads / 4) + 8;
	const long long cap = 148ll * kCellBlocksPerSM * 6;
	if(blocks > cap) blocks = cap;
	k_sweep_fused<<<(unsigned) blocks, kCellThreads, 0, st>>>(sc, ranges, cb.cell_start, cb.work, proposals, pos, ori, cb.posf, counters, hm);
	return 1;
}

} // namespace pp
