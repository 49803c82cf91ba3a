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
	counters->pad = 0;
}

// ---------------------------------------------------------------------------------------------
// K3, streaming version (the default), two kernels.
//
// k_bonds_screen: one block per segment of a pencil of cells along z, one thread per particle.  Every unordered pair
// is met once: the particle's own cell (slot_q > slot_p) and the 13 neighbour cells with (dx, dy, dz) > (0, 0, 0) in
// lexicographic order, i.e. five columns -- (0, 0): the rest of the cell and the cell above it in z; (0, +1),
// (+1, -1), (+1, 0), (+1, +1): three cells each.  With the linear cell index (ix ny + iy) nz + iz a column is ONE
// contiguous run of slots, so the block brings the five runs it needs (its own z range plus one cell at either end; at
// the ends of the box a copy of the periodic neighbour cell takes that place, so the three cells a particle looks at
// are always contiguous in the staged run) into shared memory with coalesced 16-byte loads of the fixed-point
// shadows, together with the cell bounds of those runs: two global round trips per block, everything after that is
// served from shared memory.  A thread walks its particle's five ranges as one flat loop (the lanes of a warp then
// differ by the spread of the neighbourhood size, not by the spread of each column).  Every candidate is screened in
// FP32 from the shadows (box fractions; the same conservative screen as the sweep: squared distance below
// r_cut^2 + 2e-3, shadow arithmetic good to 1e-5) and, clear of the hard core, by orientation: a bond needs one of
// p's patches to point at q (MC.c:249-255), decided from p's matrix in FP32 through a relative band.  What survives
// -- a few per cent of the pairs in range -- is appended to a pair queue in global memory (per-block queue in shared
// memory, one global atomic per block).
// k_bonds_exact: one thread per queued pair, the reference's FP64 arithmetic (MC.c:233-262), integer reduction.
// The screen runs no FP64 arithmetic and touches no FP64 position: 48 registers, ten blocks per SM.
// The global overlap flag is polled by every block of both kernels: after the first overlap the rest returns at once,
// like the reference's `while(!overlap_found)` loops (MC.c:488-495, 556-560).
// Needs at least four cells per axis (the half shell would meet a cell twice on an axis with two); smaller grids
// use the tile kernel above, and so does a configuration whose queue overflows (signalled through Counters.pad).
// A run longer than the staging area (possible only in an overlapping configuration) is read from global memory
// instead, same arithmetic.
// A slab engine counts the pairs of its owned layers among themselves and with the right halo (dx = +1), never
// with the left halo (dx = -1 is not in the half shell): every cross-boundary pair is counted once, on its left.
// ---------------------------------------------------------------------------------------------
constexpr int kPencilMaxThreads = 128;
constexpr int kPencilMaxCells = 96;   // cells of a pencil per block
constexpr int kPencilMaxCap = 288;    // staged shadows per column (hard spheres: < 2 per cell of side r_cut)
constexpr int kPencilQueue = 192;     // queued pairs per block (beyond: straight to the global queue)

// per column of the half shell: where its staged run comes from -- [copy of cell nz - 1][cells z_lo .. z_hi - 1][copy of cell 0]
struct PencilColumn {
	int g0, n_main;
	int lw0, n_lw;
	int hw0, n_hw;
};

__device__ __forceinline__ int pencil_slot(const PencilColumn &pc, int j) { // staged index -> slot
	return j < pc.n_lw ? pc.lw0 + j : (j < pc.n_lw + pc.n_main ? pc.g0 + (j - pc.n_lw) : pc.hw0 + (j - pc.n_lw - pc.n_main));
}

__global__ void __launch_bounds__(kPencilMaxThreads, 10) k_bonds_screen(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start,
		const int *__restrict__ cell_sorted, const double *__restrict__ ori, const uint4 *__restrict__ posf, const Counters *counters, int2 *queue,
		const int queue_cap, const int seg_cells, const int n_seg, const int cap) {
	constexpr int kRow = kPencilMaxCells + 3;
	extern __shared__ __align__(16) unsigned char s_raw[];
	__shared__ PencilColumn s_col[5];
	__shared__ int s_nq, s_qbase;
	if(*((volatile const int *) &counters->overlap) != 0) return;
	const int tid = threadIdx.x;
	const int ny = sc.n_side[1], nz = sc.n_side[2];
	const int pencil = (int) blockIdx.x / n_seg, seg = (int) blockIdx.x - pencil * n_seg;
	const int ix = sc.own_lo + pencil / ny, iy = pencil - (pencil / ny) * ny;
	const int z0 = seg * seg_cells, z1 = min(z0 + seg_cells, nz); // own cells [z0, z1)
	const int nk = z1 - z0 + 2;                                    // window cell k <-> cell z0 - 1 + k (periodic)
	const int z_lo = max(z0 - 1, 0), z_hi = min(z1 + 1, nz);
	uint4 *s_stage = reinterpret_cast<uint4 *>(s_raw);                              // [5][cap]
	int2 *s_q = reinterpret_cast<int2 *>(s_raw + sizeof(uint4) * 5 * cap);          // [kPencilQueue]
	int *s_beg = reinterpret_cast<int *>(s_q + kPencilQueue);                       // [5][kRow]: staged index of the first member of window cell k (k = nk: the end)
	if(tid == 0) s_nq = 0;

	// ---- cell bounds of the five columns: one round trip (the loads of an iteration are independent, iterations unrolled) ----
	const int n_items = 5 * (nk + 1);
#pragma unroll 4
	for(int i = tid; i < n_items; i += blockDim.x) {
		const int c = i / (nk + 1), k = i - c * (nk + 1);
		const int dx = c >= 2, dy = c == 0 ? 0 : (c == 1 ? 1 : c - 3);
		int cx = ix + dx;
		if(!sc.slab) cx -= (cx >= sc.nxl) ? sc.nxl : 0; // (slab: an owned layer always has its +x neighbour in the held range)
		int cy = iy + dy;
		cy += (cy < 0) ? ny : 0;
		cy -= (cy >= ny) ? ny : 0;
		const int *cs = cell_start + (cx * ny + cy) * nz;
		const int z = z0 - 1 + k; // -1 .. nz + 1; k = nk is one past the last window cell
		// the run is staged as [copy of cell nz - 1 (z0 == 0)][cells z_lo .. z_hi - 1][copy of cell 0 (z1 == nz)]
		const int g0 = cs[z_lo], c_first = cs[0], c_second = cs[1], c_last = cs[nz - 1], c_end = cs[nz], c_z = cs[min(max(z, 0), nz)];
		const int n_lw = z0 == 0 ? c_end - c_last : 0, n_hw = z1 == nz ? c_second - c_first : 0;
		int at;
		if(z < 0) at = 0;
		else if(z <= nz) at = n_lw + (c_z - g0);
		else at = n_lw + (c_end - g0) + n_hw;
		s_beg[c * kRow + k] = at;
		if(k == 0) {
			PencilColumn pc;
			pc.g0 = g0;
			pc.n_main = cs[z_hi] - g0;
			pc.lw0 = c_last;
			pc.n_lw = n_lw;
			pc.hw0 = c_first;
			pc.n_hw = n_hw;
			s_col[c] = pc;
		}
	}
	__syncthreads();
	int tot[5];
	bool fits = true;
#pragma unroll
	for(int c = 0; c < 5; c++) {
		tot[c] = s_col[c].n_main + s_col[c].n_lw + s_col[c].n_hw;
		fits &= tot[c] <= cap;
	}
	const PencilColumn own = s_col[0];
	const int own_first = pencil_slot(own, s_beg[1]), n_own = s_beg[nk - 1] - s_beg[1];
	// this thread's particle (first round): its cell and matrix are on their way while the shadows are staged
	int cell_p = 0;
	double m_p[9];
	if(tid < n_own) {
		cell_p = cell_sorted[own_first + tid];
#pragma unroll
		for(int k = 0; k < 9; k++) m_p[k] = ori[9ll * (own_first + tid) + k];
	}

	// ---- the five runs of shadows, as one flat copy ---------------------------------------------------------------
	if(fits) {
		const int t1 = tot[0], t2 = t1 + tot[1], t3 = t2 + tot[2], t4 = t3 + tot[3], t5 = t4 + tot[4];
#pragma unroll 4
		for(int i = tid; i < t5; i += blockDim.x) {
			const int c = (i >= t1) + (i >= t2) + (i >= t3) + (i >= t4);
			const int j = i - (c == 0 ? 0 : (c == 1 ? t1 : (c == 2 ? t2 : (c == 3 ? t3 : t4))));
			s_stage[c * cap + j] = posf[pencil_slot(s_col[c], j)];
		}
	}
	__syncthreads();

	const float ux = (float) (sc.box[0] / 4294967296.), uy = (float) (sc.box[1] / 4294967296.), uz = (float) (sc.box[2] / 4294967296.);
	const float thr = (float) sc.sqr_rcut + 2e-3f, core_hi = 1.f + 2e-3f;
	const float point_thr = sc.fast_patch ? (float) (sc.cos2_lo * sc.inv_b2) * (1.f - 2e-3f) : -1.f;
	for(int base = 0; base < n_own; base += blockDim.x) {
		const int sp = own_first + base + tid;
		if(base + tid < n_own) {
			if(base > 0) { // (more particles than threads: a dense packing)
				cell_p = cell_sorted[sp];
#pragma unroll
				for(int k = 0; k < 9; k++) m_p[k] = ori[9ll * sp + k];
			}
			float mf[9];
#pragma unroll
			for(int k = 0; k < 9; k++) mf[k] = (float) m_p[k];
			const int k_own = cell_p % nz - z0 + 1;
			const int j_own = own.n_lw + (sp - own.g0); // staged index of the particle itself (column 0)
			const uint4 sh = fits ? s_stage[j_own] : posf[sp];
			// the five ranges as one loop: column 0 from the particle itself to the end of the cell above, the others three cells
			int c = 0, j = j_own + 1, e = s_beg[k_own + 2];
			while(true) {
				while(j >= e && c < 4) {
					c++;
					j = s_beg[c * kRow + k_own - 1];
					e = s_beg[c * kRow + k_own + 2];
				}
				if(j >= e) break;
				const uint4 qh = fits ? s_stage[c * cap + j] : posf[pencil_slot(s_col[c], j)];
				j++;
				const float ax = shadow_sep(qh.x, sh.x, ux), ay = shadow_sep(qh.y, sh.y, uy), az = shadow_sep(qh.z, sh.z, uz);
				const float d2 = fmaf(az, az, fmaf(ay, ay, ax * ax));
				if(d2 >= thr) continue;
				if(d2 >= core_hi && !may_point(mf, ax, ay, az, d2, point_thr)) continue;
				const int2 pair = make_int2(sp, (int) qh.w); // the shadow carries its slot
				const int slot = atomicAdd(&s_nq, 1);
				if(slot < kPencilQueue) s_q[slot] = pair;
				else { // (more pairs than the block's queue holds: straight to the global queue)
					const int g = atomicAdd(&queue[0].x, 1);
					if(g < queue_cap) queue[1 + g] = pair;
					else queue[0].y = 1;
				}
			}
		}
	}
	__syncthreads();
	const int n = min(s_nq, kPencilQueue);
	if(n == 0) return;
	if(tid == 0) s_qbase = atomicAdd(&queue[0].x, n);
	__syncthreads();
	const int gb = s_qbase;
	for(int i = tid; i < n; i += blockDim.x) {
		if(gb + i < queue_cap) queue[1 + gb + i] = s_q[i];
		else queue[0].y = 1;
	}
}

// FP64 pair test of the queued pairs in the reference's operation order (MC.c:233-262) -> bond count, overlap
__global__ void __launch_bounds__(128) k_bonds_exact(const __grid_constant__ SimConst sc, const double4 *__restrict__ pos, const double *__restrict__ ori,
		const int2 *__restrict__ queue, const int queue_cap, Counters *counters) {
	__shared__ int s_bonds, s_overlap;
	if(threadIdx.x == 0) {
		s_bonds = 0;
		s_overlap = 0;
	}
	__syncthreads();
	const int2 head = queue[0];
	if(head.y != 0 && blockIdx.x == 0 && threadIdx.x == 0) counters->pad = 1; // queue overflow: the host repeats the count with the tile kernel
	const int n = min(head.x, queue_cap);
	int bonds = 0, overlap = 0;
	for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		if(*((volatile int *) &counters->overlap) != 0) break;
		const int2 pr = queue[1 + i];
		// both positions and both matrices in one round trip (most queued pairs are inside the shell and need them all)
		const double4 p4 = pos[pr.x], q4 = pos[pr.y];
		double m[9], mq[9];
#pragma unroll
		for(int k = 0; k < 9; k++) {
			m[k] = ori[9ll * pr.x + k];
			mq[k] = ori[9ll * pr.y + k];
		}
		const Vec3 p = {p4.x, p4.y, p4.z};
		Vec3 d;
		const double d2 = min_image(sc, p, q4, d);
		if(d2 < 1.) overlap = 1;
		else if(d2 < sc.sqr_rcut) bonds += patches_bond_m(sc, d, d2, m, mq) ? 1 : 0;
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

__global__ void k_zero_bonds_queue(Counters *counters, int2 *queue) {
	counters->bonds = 0;
	counters->overlap = 0;
	counters->pad = 0;
	queue[0] = make_int2(0, 0);
}

// scratch: device buffer of scratch_bytes the pair queue may use (nullptr: tile kernel); force_tile: the exact tile kernel whatever the grid
int launch_total_bonds(const SimConst &sc, const int *cell_start, const int *cell_sorted, const double4 *pos, const double *ori, const uint4 *posf,
		Counters *counters, void *scratch, size_t scratch_bytes, bool force_tile, cudaStream_t st) {
	const bool wide = (sc.slab || sc.nxl >= 4) && sc.n_side[1] >= 4 && sc.n_side[2] >= 4;
	if(wide && !force_tile && posf != nullptr && scratch != nullptr && scratch_bytes >= 4096 && sc.N > 0) {
		int2 *queue = reinterpret_cast<int2 *>(scratch);
		const long long qcap = (long long) (scratch_bytes / sizeof(int2)) - 1;
		const int queue_cap = qcap > 0x3fffffffll ? 0x3fffffff : (int) qcap;
		k_zero_bonds_queue<<<1, 1, 0, st>>>(counters, queue);
		const int nz = sc.n_side[2];
		const int n_seg = (nz + kPencilMaxCells - 1) / kPencilMaxCells, seg_cells = (nz + n_seg - 1) / n_seg;
		const long long n_pencils = (long long) (sc.own_hi - sc.own_lo) * sc.n_side[1];
		// staging area per column: twice the mean population of the window at one particle per cell (a hard-sphere
		// packing holds fewer than two per cell of side r_cut; cells of a small box are wider, the window is shorter)
		const double occ = (double) sc.N / ((double) sc.n_cells > 0. ? (double) sc.n_cells : 1.);
		int cap = (int) ((seg_cells + 4) * (occ > 1. ? 2. * occ : 2.)) + 16;
		if(cap > kPencilMaxCap) cap = kPencilMaxCap;
		const double own = occ * seg_cells; // particles of a block
		// a grid that fits the machine at once is latency bound: the widest block brings the bounds and the shadows in the fewest rounds
		const bool small = n_pencils * n_seg * kPencilMaxThreads <= 148 * 2048; // (every block resident at once even at full width)
		const int threads = small ? kPencilMaxThreads : (own <= 26. ? 32 : (own <= 56. ? 64 : (own <= 84. ? 96 : kPencilMaxThreads)));
		const size_t smem = sizeof(uint4) * 5 * cap + sizeof(int2) * kPencilQueue + sizeof(int) * 5 * (kPencilMaxCells + 3);
		k_bonds_screen<<<(unsigned) (n_pencils * n_seg), threads, smem, st>>>(sc, cell_start, cell_sorted, ori, posf, counters, queue, queue_cap, seg_cells,
				n_seg, cap);
		long long est = (long long) sc.N / 2 + 128; // pairs expected in the queue at most in a fluid; the kernel strides over whatever is there
		int blocks = (int) ((est + 127) / 128 > 148 * 8 ? 148 * 8 : (est + 127) / 128);
		k_bonds_exact<<<blocks, 128, 0, st>>>(sc, pos, ori, queue, queue_cap, counters);
		return 3;
	}
	k_zero_bonds<<<1, 1, 0, st>>>(counters);
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
