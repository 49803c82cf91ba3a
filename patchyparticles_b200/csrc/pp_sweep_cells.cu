// pp_sweep_cells.cu -- K2b, one THREAD per non-empty cell of the current colour.
//
// Same algorithm, same random streams and bit-identical trajectories as the tile kernel in
// pp_sweep.cu (both are checked against the serial restatement in oracle/pp_oracle.c); what changes
// is the mapping onto the machine.  A thread walks its cell's trials on its own, so nothing is
// computed redundantly, shuffled or reduced: ~9x fewer warp instructions per trial than the tile
// kernel.  The price is a long dependent chain per thread, so the kernel is organised around
// memory-level parallelism:
//
//   per cell   the FP32 shadow positions (float4: offset inside the cell + the cell's z index) of the
//              whole 27-cell neighbourhood are brought into shared memory with cp.async -- nine
//              contiguous z-runs, all copies in flight at once, one wait;
//   per trial  1. FP32 pre-screen of the staged neighbours for the old and the new position; a pair
//                 is a candidate when its FP32 squared distance is below r_cut^2 + 1e-3 (the FP32
//                 error is < 1e-5, so no pair the FP64 test would put inside the shell, or flag as
//                 overlapping, is ever dropped);
//              2. FP64, reference operation order, on the candidates only (their double positions
//                 are fetched four at a time): overlap of the new position (MC.c:240), shell test
//                 (MC.c:241), orientation tests (MC.c:249-261);
//              3. Metropolis rule (MC.c:332), write-back of position, orientation, FP32 shadow.
//
// The work list hands consecutive threads cells of equal occupancy, which keeps the lanes of a warp
// in the same trial.
#include "pp_trial_serial.cuh"

#include <cuda_pipeline.h>

namespace pp {

constexpr int kCellThreads = 64;
constexpr int kNbCap = 32;                   // staged neighbours per cell (mean 24 at rho = 0.6)
constexpr int kCandCap = 24;                 // candidates per trial
constexpr float kScreenBand = 1e-3f;         // FP32 safety band around r_cut^2
constexpr unsigned kNearOld = 0x40u, kNearNew = 0x80u, kIdxMask = 0x3fu;

struct CellBlockCounters {
	unsigned int tries, accepted, confined, overlapped;
	int dbonds;
	unsigned long long w_tries, w_accepted, pair_checks;
};

__global__ void __launch_bounds__(kCellThreads, 5) k_sweep_cells(const __grid_constant__ SimConst sc, int colour, const int *__restrict__ cell_start,
		const int *__restrict__ work, const int *__restrict__ work_start, const Proposal *__restrict__ proposals, double4 *pos, double *ori,
		float4 *posf, Counters *__restrict__ counters) {
	__shared__ __align__(16) float4 s_stage[kNbCap][kCellThreads]; // [entry][thread]: conflict-free 16-byte accesses
	__shared__ int s_slot[kNbCap][kCellThreads];
	__shared__ unsigned char s_cand[kCandCap][kCellThreads];
	__shared__ CellBlockCounters s_cnt;
	const int tid = threadIdx.x;
	if(tid == 0) s_cnt = CellBlockCounters{0, 0, 0, 0, 0, 0, 0, 0};
	__syncthreads();

	const int w_begin = work_start[colour], w_end = work_start[colour + 1];
	const int nx = sc.n_side[0], ny = sc.n_side[1], nz = sc.n_side[2];
	const float wx = (float) sc.cell_w[0], wy = (float) sc.cell_w[1], wz = (float) sc.cell_w[2];
	const float thr = (float) sc.sqr_rcut + kScreenBand;
	Tally tl = {0, 0, 0, 0, 0, 0, 0, 0};

	for(int wi = w_begin + blockIdx.x * kCellThreads + tid; wi < w_end; wi += gridDim.x * kCellThreads) {
		const int cell = work[wi];
		const int ix = cell / (ny * nz), iy = (cell - ix * ny * nz) / nz, iz = cell - ix * ny * nz - iy * nz;
		const int first = cell_start[cell];
		const int n_here = cell_start[cell + 1] - first;

		// ---- stage the neighbourhood: 9 z-runs (+ the wrapped end cell), all copies in flight ------
		int total = 0, self_base = 0;
		int ccount[9]; // staged entries per (dx, dy) column
		{
			int rb[9], re[9], rb2[9], re2[9];
			const int z0 = iz > 0 ? iz - 1 : 0, z1 = iz < nz - 1 ? iz + 1 : nz - 1;
#pragma unroll
			for(int c = 0; c < 9; c++) {
				const int col = (wrap_cell(ix + c / 3 - 1, nx) * ny + wrap_cell(iy + c % 3 - 1, ny)) * nz;
				rb[c] = cell_start[col + z0];
				re[c] = cell_start[col + z1 + 1];
				rb2[c] = re2[c] = 0;
				if(iz == 0) {
					rb2[c] = cell_start[col + nz - 1];
					re2[c] = cell_start[col + nz];
				}
				else if(iz == nz - 1) {
					rb2[c] = cell_start[col];
					re2[c] = cell_start[col + 1];
				}
			}
#pragma unroll
			for(int c = 0; c < 9; c++) {
				if(c == 4) self_base = total + (first - rb[4]); // entry of the home cell's first particle
				ccount[c] = (re[c] - rb[c]) + (re2[c] - rb2[c]);
				for(int s = rb[c]; s < re[c]; s++, total++)
					if(total < kNbCap) {
						__pipeline_memcpy_async(&s_stage[total][tid], posf + s, sizeof(float4));
						s_slot[total][tid] = s;
					}
				for(int s = rb2[c]; s < re2[c]; s++, total++)
					if(total < kNbCap) {
						__pipeline_memcpy_async(&s_stage[total][tid], posf + s, sizeof(float4));
						s_slot[total][tid] = s;
					}
			}
			__pipeline_commit();
		}
		const bool staged = total <= kNbCap;
		__pipeline_wait_prior(0);

		for(int t = 0; t < n_here; t++) {
			const Proposal *pr = proposals + first + t;
			const int pick = pr->pick;
			const int sp = first + pick;
			tl.tries++;
			if(!staged) {
				trial_serial(sc, cell_start, pos, ori, posf, pr, cell, ix, iy, iz, sp, tl);
				continue;
			}
			const double dx = pr->dx, dy = pr->dy, dz = pr->dz;
			const double4 p4 = pos[sp];
			const Vec3 p_old = {p4.x, p4.y, p4.z};
			const Vec3 p_new = {__dadd_rn(p4.x, dx), __dadd_rn(p4.y, dy), __dadd_rn(p4.z, dz)};
			int cx, cy, cz;
			if(cell_index(sc, p_new.x, p_new.y, p_new.z, cx, cy, cz) != cell) { // would leave its cell
				tl.conf++;
				continue;
			}
			const unsigned long long Wfix = (unsigned long long) pr->wfix;
			tl.wt += Wfix;
			const int j_self = self_base + pick;
			const float4 f_old = s_stage[j_self][tid];
			const float4 f_new = cell_relative(sc, p_new.x, p_new.y, p_new.z);

			// ---- 1. FP32 pre-screen over the staged neighbourhood -------------------------
			int n_cand = 0;
			{
				int j = 0;
#pragma unroll
				for(int c = 0; c < 9; c++) {
					// entries of column c: consecutive in the staging order; their x / y offsets are those of the column
					const float ox = (float) (c / 3 - 1) * wx, oy = (float) (c % 3 - 1) * wy;
					const float ox_old = ox - f_old.x, oy_old = oy - f_old.y, ox_new = ox - f_new.x, oy_new = oy - f_new.y;
					for(int k = 0; k < ccount[c]; k++, j++) {
						if(j == j_self) continue;
						const float4 q = s_stage[j][tid];
						int dzc = (int) q.w - iz;
						dzc -= (dzc > 1) ? nz : 0;
						dzc += (dzc < -1) ? nz : 0;
						const float qz = q.z + (float) dzc * wz;
						const float ax = q.x + ox_old, ay = q.y + oy_old, az = qz - f_old.z;
						const float bx = q.x + ox_new, by = q.y + oy_new, bz = qz - f_new.z;
						const float d2o = ax * ax + ay * ay + az * az;
						const float d2n = bx * bx + by * by + bz * bz;
						const unsigned flags = (d2o < thr ? kNearOld : 0u) | (d2n < thr ? kNearNew : 0u);
						if(flags) {
							if(n_cand < kCandCap) s_cand[n_cand][tid] = (unsigned char) ((unsigned) j | flags);
							n_cand++;
						}
					}
				}
			}
			tl.checks += 2ull * (unsigned long long) (total - 1);
			if(n_cand > kCandCap) { // cannot happen for hard spheres with r_cut < 1.5; kept for safety
				trial_serial(sc, cell_start, pos, ori, posf, pr, cell, ix, iy, iz, sp, tl);
				tl.wt -= Wfix;
				tl.checks -= 2ull * (unsigned long long) (total - 1);
				continue;
			}

			// ---- 2. FP64 on the candidates, four double positions in flight ------------------
			double m[9], mn[9];
#pragma unroll
			for(int k = 0; k < 9; k++) m[k] = ori[9ll * sp + k];
#pragma unroll
			for(int i = 0; i < 3; i++)
#pragma unroll
				for(int j = 0; j < 3; j++) mn[3 * i + j] = dot3_rn(m[3 * i], m[3 * i + 1], m[3 * i + 2], pr->R[j], pr->R[3 + j], pr->R[6 + j]);
			int b_old = 0, b_new = 0;
			bool ovl = false;
			for(int c0 = 0; c0 < n_cand && !ovl; c0 += 4) {
				double4 q4[4];
				unsigned e[4];
				int sq[4];
#pragma unroll
				for(int k = 0; k < 4; k++) {
					e[k] = (c0 + k < n_cand) ? s_cand[c0 + k][tid] : 0u;
					sq[k] = s_slot[e[k] & kIdxMask][tid];
					if(e[k]) q4[k] = pos[sq[k]];
				}
#pragma unroll
				for(int k = 0; k < 4; k++) {
					if(e[k] & kNearNew) {
						Vec3 d;
						const double d2 = min_image(sc, p_new, q4[k], d);
						if(d2 < 1.) ovl = true;
						else if(d2 < sc.sqr_rcut) b_new += patches_bond_m(sc, d, d2, mn, ori + 9ll * sq[k]) ? 1 : 0;
					}
					if(e[k] & kNearOld) {
						Vec3 d;
						const double d2 = min_image(sc, p_old, q4[k], d);
						if(d2 >= 1. && d2 < sc.sqr_rcut) b_old += patches_bond_m(sc, d, d2, m, ori + 9ll * sq[k]) ? 1 : 0;
					}
				}
			}
			if(ovl) { // hard-core overlap: rejected whatever the bonds (MC.c:296-299, 332)
				tl.ovl++;
				continue;
			}

			// ---- 3. Metropolis (MC.c:332-339) ------------------------------------------------
			const int dE = b_old - b_new; // energy = -bonds
			if(dE < 0 || pr->u_acc < sc.exp_table[dE < kMaxExp ? dE : kMaxExp - 1]) {
				double *p = reinterpret_cast<double *>(pos + sp); // x, y, z; the id in .w stays
				p[0] = p_new.x;
				p[1] = p_new.y;
				p[2] = p_new.z;
				posf[sp] = f_new;
				s_stage[j_self][tid] = f_new;
#pragma unroll
				for(int k = 0; k < 9; k++) ori[9ll * sp + k] = mn[k];
				tl.dbonds += b_new - b_old;
				tl.acc++;
				tl.wa += Wfix;
			}
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

int launch_sweep_cells(const SimConst &sc, int colour, const CellBuffers &cb, const Proposal *proposals, double4 *pos, double *ori, Counters *counters,
		cudaStream_t st) {
	long long n_active = (long long) (sc.n_side[0] / 2) * (sc.n_side[1] / 2) * (sc.n_side[2] / 2);
	if(n_active > sc.N) n_active = sc.N;
	long long blocks = (n_active + kCellThreads - 1) / kCellThreads;
	if(blocks < 1) blocks = 1;
	k_sweep_cells<<<(unsigned) blocks, kCellThreads, 0, st>>>(sc, colour, cb.cell_start, cb.work, cb.bins + kWorkStart, proposals, pos, ori, cb.posf,
			counters);
	return 1;
}

} // namespace pp
