// pp_cells.cu -- K1: on-GPU cell build as a counting sort (replaces cells_init + cells_fill,
// cells.c:17-63, and the incremental relinking of MC_change_cell, MC.c:180-206).
//
// count -> exclusive scan -> scatter source indices -> order each cell by particle id -> gather.
// The result is the particle state permuted into cell order (SoA: double4 position+id, 3x3 FP64
// orientation, int id) plus the CSR offsets cell_start[n_cells + 1].  Ordering each cell by
// particle id makes the layout (and therefore every trajectory) independent of atomic ordering.
//
// The build also emits the sweep's work lists: for each of the 8 checkerboard colours the non-empty
// cells, bucketed by occupancy in descending order.  Tiles that share a warp in the sweep kernel
// then run the same number of trials (no idle tiles), and the longest chains start first.
#include "pp_kernels.cuh"

namespace pp {

__global__ void __launch_bounds__(256) k_cell_count(const __grid_constant__ SimConst sc, const double4 *__restrict__ pos, int *__restrict__ cell_of,
		int *__restrict__ cell_count) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= sc.N) return;
	const double4 p = pos[i];
	int ix, iy, iz;
	int c = cell_index(sc, p.x, p.y, p.z, ix, iy, iz);
	if(c < 0) c = sc.n_cells; // not in a layer held here (slab decomposition): parked in the trash cell, dropped after the sort
	cell_of[i] = c;
	atomicAdd(cell_count + c, 1);
}

// ---- exclusive scan over the cell histogram -------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem /* 32 ints */, int &block_total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int incl = v;
#pragma unroll
	for(int o = 1; o < 32; o <<= 1) {
		const int t = __shfl_up_sync(0xffffffffu, incl, o);
		if(lane >= o) incl += t;
	}
	if(lane == 31) smem[warp] = incl;
	__syncthreads();
	if(warp == 0) {
		int w = (lane < (int) (blockDim.x >> 5)) ? smem[lane] : 0;
#pragma unroll
		for(int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, w, o);
			if(lane >= o) w += t;
		}
		smem[lane] = w;
	}
	__syncthreads();
	const int warp_offset = (warp > 0) ? smem[warp - 1] : 0;
	block_total = smem[(blockDim.x >> 5) - 1];
	__syncthreads();
	return warp_offset + incl - v;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_tiles(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ tile_sums, int n) {
	__shared__ int smem[32];
	const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
	int v[kScanItems], sum = 0;
#pragma unroll
	for(int k = 0; k < kScanItems; k++) {
		v[k] = (base + k < n) ? in[base + k] : 0;
		sum += v[k];
	}
	int total;
	int run = block_exclusive_scan(sum, smem, total);
#pragma unroll
	for(int k = 0; k < kScanItems; k++) {
		if(base + k < n) out[base + k] = run;
		run += v[k];
	}
	if(threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int *__restrict__ tile_sums, int n_tiles, int *__restrict__ grand_total) {
	__shared__ int smem[32];
	int carry = 0;
	for(int base = 0; base < n_tiles; base += 1024) {
		const int i = base + threadIdx.x;
		const int v = (i < n_tiles) ? tile_sums[i] : 0;
		int total;
		const int ex = block_exclusive_scan(v, smem, total);
		if(i < n_tiles) tile_sums[i] = carry + ex;
		carry += total;
	}
	if(threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_add(int *__restrict__ out, const int *__restrict__ tile_sums, int n) {
	const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
	const int add = tile_sums[blockIdx.x];
#pragma unroll
	for(int k = 0; k < kScanItems; k++)
		if(base + k < n) out[base + k] += add;
}

// ---- scatter / order / gather -----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cell_scatter(int N, const int *__restrict__ cell_of, const int *__restrict__ cell_start,
		int *__restrict__ cell_fill, int *__restrict__ perm) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= N) return;
	const int c = cell_of[i];
	perm[cell_start[c] + atomicAdd(cell_fill + c, 1)] = i;
}

// particles in the 27-cell neighbourhood of an owned cell (periodic wrap, or the slab's halo layers along x)
__device__ __forceinline__ int neighbourhood_total(const SimConst &sc, const int *__restrict__ cell_start, int ix, int iy, int iz) {
	const int ny = sc.n_side[1], nz = sc.n_side[2];
	int total = 0;
	for(int dx = -1; dx <= 1; dx++) {
		int cx = ix + dx;
		if(!sc.slab) {
			cx += (cx < 0) ? sc.nxl : 0;
			cx -= (cx >= sc.nxl) ? sc.nxl : 0;
		}
		for(int dy = -1; dy <= 1; dy++) {
			int cy = iy + dy;
			cy += (cy < 0) ? ny : 0;
			cy -= (cy >= ny) ? ny : 0;
			const int col = (cx * ny + cy) * nz;
			const int z0 = iz > 0 ? iz - 1 : 0, z1 = iz < nz - 1 ? iz + 1 : nz - 1;
			total += cell_start[col + z1 + 1] - cell_start[col + z0]; // the z-run is contiguous
			if(iz == 0) total += cell_start[col + nz] - cell_start[col + nz - 1];
			else if(iz == nz - 1) total += cell_start[col + 1] - cell_start[col];
		}
	}
	return total;
}

// bin of a non-empty cell in the work lists: colour-major, occupancy descending inside a colour.  A cell whose
// neighbourhood does not fit the staging area of the thread-per-cell kernel is filed with the most crowded ones,
// which always go to the tile kernel.
__device__ __forceinline__ int work_bin(const SimConst &sc, int ix, int iy, int iz, int n, bool dense, int nb_total) {
	const int colour = ((ix + sc.x_lo) & 1) | ((iy & 1) << 1) | ((iz & 1) << 2); // parity of the GLOBAL layer
	const int bucket = (dense || n >= kWorkBuckets) ? kWorkBuckets : n;            // 1 .. kWorkBuckets
	// staged entries rounded up to four, in fours (classes 1 .. kWorkSub - 1); class 0 collects the cells of the first and last z
	// layer, whose neighbourhood wraps around the box: the sweep kernel stages those with an extra loop, which then runs in a few
	// warps full of such cells instead of one lane of every other warp
	const bool edge = iz == 0 || iz == sc.n_side[2] - 1;
	const int cls = edge ? 0 : 1 + (nb_total > 0 ? min((nb_total - 1) >> 2, kWorkSub - 2) : 0);
	return (colour * kWorkBuckets + (kWorkBuckets - bucket)) * kWorkSub + (kWorkSub - 1 - cls); // big cells first, big neighbourhoods first
}

// per cell: order its particles by id (insertion sort; a handful of particles), choose and count its work bin.
// The bins are few and hot: they are counted per block in shared memory, one global atomic per bin and block.
__global__ void __launch_bounds__(256) k_cell_order(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start, int *__restrict__ perm,
		const int *__restrict__ pid_src, int *__restrict__ bin_count, int *__restrict__ cell_bin, int cells_max_occupancy, int cells_nb_cap) {
	__shared__ int s_bin[kWorkBins];
	for(int b = threadIdx.x; b < kWorkBins; b += blockDim.x) s_bin[b] = 0;
	__syncthreads();
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c < sc.n_cells) {
		const int b = cell_start[c], e = cell_start[c + 1];
		if(e > b) {
			const int nyz = sc.n_side[1] * sc.n_side[2];
			const int ix = c / nyz, iy = (c - ix * nyz) / sc.n_side[2], iz = c - ix * nyz - iy * sc.n_side[2];
			for(int a = b + 1; a < e; a++) {
				const int src = perm[a];
				const int key = pid_src[src];
				int k = a - 1;
				while(k >= b && pid_src[perm[k]] > key) {
					perm[k + 1] = perm[k];
					k--;
				}
				perm[k + 1] = src;
			}
			if(ix >= sc.own_lo && ix < sc.own_hi) { // halo cells are never moved here
				const int nb_total = (e - b <= cells_max_occupancy) ? neighbourhood_total(sc, cell_start, ix, iy, iz) : 0;
				const bool dense = nb_total > cells_nb_cap;
				const int bin = work_bin(sc, ix, iy, iz, e - b, dense, dense ? 0 : nb_total);
				cell_bin[c] = bin;
				atomicAdd(&s_bin[bin], 1);
			}
		}
	}
	__syncthreads();
	for(int b = threadIdx.x; b < kWorkBins; b += blockDim.x)
		if(s_bin[b]) atomicAdd(bin_count + b, s_bin[b]);
}

// one block of kWorkBins threads: exclusive scan of the bins
__global__ void __launch_bounds__(kWorkBins) k_bins_scan(const int *__restrict__ bin_count, int *__restrict__ bin_start, int *__restrict__ ranges,
		int cells_max_occupancy, int cells_mid_occupancy) {
	__shared__ int s[kWorkBins + 1];
	__shared__ int warp_sums[32];
	{
		int total;
		const int ex = block_exclusive_scan(bin_count[threadIdx.x], warp_sums, total);
		s[threadIdx.x] = ex;
		if(threadIdx.x == 0) s[kWorkBins] = total;
	}
	__syncthreads();
	for(int b = threadIdx.x; b <= kWorkBins; b += blockDim.x) bin_start[b] = s[b];
	if(threadIdx.x < 8) { // split each colour's list (occupancy descending): crowded cells, middle range, cells with <= cells_max_occupancy members
		const int c = threadIdx.x, m1 = cells_max_occupancy < kWorkBuckets ? cells_max_occupancy : kWorkBuckets;
		const int m2 = cells_mid_occupancy < m1 ? m1 : (cells_mid_occupancy < kWorkBuckets ? cells_mid_occupancy : kWorkBuckets);
		const int split2 = s[(c * kWorkBuckets + (kWorkBuckets - m2)) * kWorkSub], split1 = s[(c * kWorkBuckets + (kWorkBuckets - m1)) * kWorkSub];
		ranges[6 * c + 0] = s[c * kWorkBuckets * kWorkSub];
		ranges[6 * c + 1] = split2;
		ranges[6 * c + 2] = split2;
		ranges[6 * c + 3] = split1;
		ranges[6 * c + 4] = split1;
		ranges[6 * c + 5] = s[(c + 1) * kWorkBuckets * kWorkSub];
	}
}

// work lists: rank inside the block's share of a bin from shared memory, one global atomic per bin and block
__global__ void __launch_bounds__(256) k_work_scatter(const __grid_constant__ SimConst sc, const int *__restrict__ cell_start,
		const int *__restrict__ bin_start, int *__restrict__ bin_fill, const int *__restrict__ cell_bin, int *__restrict__ work) {
	__shared__ int s_cnt[kWorkBins], s_base[kWorkBins];
	for(int b = threadIdx.x; b < kWorkBins; b += blockDim.x) s_cnt[b] = 0;
	__syncthreads();
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	int bin = -1, rank = 0;
	if(c < sc.n_cells) {
		const int n = cell_start[c + 1] - cell_start[c];
		const int layer = c / (sc.n_side[1] * sc.n_side[2]);
		if(n > 0 && layer >= sc.own_lo && layer < sc.own_hi) {
			bin = cell_bin[c];
			rank = atomicAdd(&s_cnt[bin], 1);
		}
	}
	__syncthreads();
	for(int b = threadIdx.x; b < kWorkBins; b += blockDim.x)
		if(s_cnt[b]) s_base[b] = bin_start[b] + atomicAdd(bin_fill + b, s_cnt[b]);
	__syncthreads();
	if(bin >= 0) work[s_base[bin] + rank] = c;
}

__global__ void __launch_bounds__(256) k_cell_gather_pos(const __grid_constant__ SimConst sc, const int *__restrict__ perm,
		const int *__restrict__ cell_of, const double4 *__restrict__ pos_src, const int *__restrict__ pid_src, double4 *__restrict__ pos_dst,
		int *__restrict__ pid_dst, int *__restrict__ slot_of, int *__restrict__ cell_sorted, const int *__restrict__ cell_start,
		uint4 *__restrict__ posf) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= sc.N) return;
	const int src = perm[s];
	const double4 p = pos_src[src];
	pos_dst[s] = p;
	const int id = pid_src[src];
	pid_dst[s] = id;
	slot_of[id] = s;
	const int c = cell_of[src];
	cell_sorted[s] = c;
	posf[s] = box_shadow(sc, p.x, p.y, p.z, s); // shadow used by the pre-screens of the sweep and of the bond count
}

__global__ void __launch_bounds__(256) k_cell_gather_ori(long long n9, const int *__restrict__ perm, const double *__restrict__ ori_src,
		double *__restrict__ ori_dst) {
	const long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	if(t >= n9) return;
	const int s = (int) (t / 9), comp = (int) (t - 9ll * s);
	ori_dst[t] = ori_src[9ll * perm[s] + comp];
}

// ---------------------------------------------------------------------------------------------
int launch_cell_build(const SimConst &sc, const CellBuffers &cb, const StateView &src, const StateView &dst, int *slot_of, int cells_max_occupancy,
		int cells_mid_occupancy, cudaStream_t st) {
	const int N = sc.N;
	const int nc = sc.n_cells + 1; // + the trash cell that collects particles outside the held layers
	cudaMemsetAsync(cb.bins, 0, sizeof(int) * kBinInts, st);
	if(N == 0) {
		cudaMemsetAsync(cb.cell_start, 0, sizeof(int) * (size_t) (nc + 1), st);
		return 0;
	}
	cudaMemsetAsync(cb.cell_count, 0, sizeof(int) * (size_t) nc, st);
	cudaMemsetAsync(cb.cell_fill, 0, sizeof(int) * (size_t) nc, st);
	const int nb = (N + 255) / 256, ncb = (sc.n_cells + 255) / 256;
	k_cell_count<<<nb, 256, 0, st>>>(sc, src.pos, cb.cell_of, cb.cell_count);
	const int n_tiles = (nc + kScanTile - 1) / kScanTile;
	k_scan_tiles<<<n_tiles, kScanThreads, 0, st>>>(cb.cell_count, cb.cell_start, cb.tile_sums, nc);
	k_scan_sums<<<1, 1024, 0, st>>>(cb.tile_sums, n_tiles, cb.cell_start + nc);
	k_scan_add<<<n_tiles, kScanThreads, 0, st>>>(cb.cell_start, cb.tile_sums, nc);
	k_cell_scatter<<<nb, 256, 0, st>>>(N, cb.cell_of, cb.cell_start, cb.cell_fill, cb.perm);
	k_cell_order<<<ncb, 256, 0, st>>>(sc, cb.cell_start, cb.perm, src.pid, cb.bins + kBinCount, cb.cell_fill, cells_max_occupancy, kCellsNbCap);
	k_bins_scan<<<1, kWorkBins, 0, st>>>(cb.bins + kBinCount, cb.bins + kBinStart, cb.bins + kWorkRanges, cells_max_occupancy, cells_mid_occupancy);
	k_work_scatter<<<ncb, 256, 0, st>>>(sc, cb.cell_start, cb.bins + kBinStart, cb.bins + kBinFill, cb.cell_fill, cb.work);
	k_cell_gather_pos<<<nb, 256, 0, st>>>(sc, cb.perm, cb.cell_of, src.pos, src.pid, dst.pos, dst.pid, slot_of, cb.cell_sorted, cb.cell_start, cb.posf);
	const long long n9 = 9ll * N;
	k_cell_gather_ori<<<(unsigned) ((n9 + 255) / 256), 256, 0, st>>>(n9, cb.perm, src.ori, dst.ori);
	return 10;
}

} // namespace pp
