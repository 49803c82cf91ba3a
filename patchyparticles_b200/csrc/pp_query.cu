// pp_query.cu -- neighbour queries of the cluster moves, answered on the GPU from the
// cell-sorted state: the moves themselves (AVB, avb.c; VMMC, vmmc.c) stay on the host, what they ask of the configuration is
// the 27-cell pair loop of the hot path.
//   pp_patch_neighbours  _set_neighbours (avb.c:43-75): for a particle, the bonded neighbour sitting on each of its patches
//                        (MC_interact's onp, MC.c:249-261) and the number of bonded neighbours;
//   pp_range_neighbours  the geometric part of _populate_possible_links (vmmc.c:148-229): the particles within the
//                        interaction range (dist2 < kf_sqr_rcut, vmmc.c:178) of a particle.
// Both walk the neighbourhood in the reference's order -- cells (j, k, l) = x outer, z inner -- and the members of a cell by
// ascending index (the reference follows its linked lists; the order matters only where two neighbours share a patch, and
// there the last one in this order stays, as `neighbours[p_patch] = q` leaves it).  A batch of particles per call.
#include "pp_internal.h"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;
using namespace pp;
using namespace ppi;

namespace pp {

constexpr int kQueryThreads = 128;

// MC_would_interact (MC.c:232-266) with the patches it reports: 1 bond (onp: the first patch of p that points at q and finds a
// patch of q pointing back), 0 none, -1 overlap.  *in_range: dist2 < kf_sqr_rcut, overlaps included (vmmc.c:178).
__device__ __forceinline__ int pair_with_patches(const SimConst &sc, const Vec3 &p, const double *pp, const double4 &q, const double *__restrict__ ori_q,
		int &onp, int &onq, bool &in_range) {
	Vec3 d;
	const double d2 = min_image(sc, p, q, d);
	in_range = d2 < sc.sqr_rcut;
	if(d2 < 1.) return -1;
	if(!in_range) return 0;
	const double norm = __dsqrt_rn(d2);
	d.x = __ddiv_rn(d.x, norm);
	d.y = __ddiv_rn(d.y, norm);
	d.z = __ddiv_rn(d.z, norm);
	double mq[9], pq[12];
#pragma unroll
	for(int k = 0; k < 9; k++) mq[k] = ori_q[k];
	patches_from(sc.base, mq, pq);
	for(int i = 0; i < 4; i++) {
		const double p_cos = dot3_rn(d.x, d.y, d.z, pp[3 * i], pp[3 * i + 1], pp[3 * i + 2]);
		if(p_cos > sc.cosmax)
			for(int j = 0; j < 4; j++) {
				const double q_cos = -dot3_rn(d.x, d.y, d.z, pq[3 * j], pq[3 * j + 1], pq[3 * j + 2]);
				if(q_cos > sc.cosmax) {
					onp = i;
					onq = j;
					return 1;
				}
			}
	}
	return 0;
}

// one tile of G lanes per queried particle.  mode 0: out_list[k][4] = index of the bonded neighbour on each patch (-1: free),
// out_count[k] = bonded neighbours; mode 1: out_list[k][max_nb] = indices of the particles in range, out_count[k] = how many
template<int G>
__global__ void __launch_bounds__(kQueryThreads) k_neighbour_query(const __grid_constant__ SimConst sc, int n, const int *__restrict__ idx, int mode, int max_nb,
		const int *__restrict__ cell_start, const double4 *__restrict__ pos, const double *__restrict__ ori, const int *__restrict__ pid,
		const int *__restrict__ slot_of, int *__restrict__ out_count, int *__restrict__ out_list) {
	constexpr int kTiles = kQueryThreads / G;
	__shared__ Neighbourhood s_nb[kTiles];
	cg::thread_block block = cg::this_thread_block();
	cg::thread_block_tile<G> tile = cg::tiled_partition<G>(block);
	const int tile_in_block = threadIdx.x / G, lane = tile.thread_rank();
	Neighbourhood &nb = s_nb[tile_in_block];
	for(int k = blockIdx.x * kTiles + tile_in_block; k < n; k += gridDim.x * kTiles) {
		const int sp = slot_of[idx[k]];
		const double4 p4 = pos[sp];
		const Vec3 p = {p4.x, p4.y, p4.z};
		double m[9], pp[12];
#pragma unroll
		for(int c = 0; c < 9; c++) m[c] = ori[9ll * sp + c];
		patches_from(sc.base, m, pp);
		int ix, iy, iz;
		cell_index(sc, p.x, p.y, p.z, ix, iy, iz);
		tile.sync();
		const int total = build_neighbourhood<G>(tile, sc, cell_start, ix, iy, iz, nb);
		const int width = mode == 0 ? 4 : max_nb;
		for(int c = lane; c < width; c += G) out_list[(long long) k * width + c] = -1;
		tile.sync();
		int count = 0;
		for(int j0 = 0; j0 < total; j0 += G) { // the whole tile in step: results are placed in neighbourhood order
			const int j = j0 + lane;
			int sq = -1, onp = -1, onq = -1;
			bool hit = false;
			if(j < total) {
				sq = neighbour_slot(nb, j);
				if(sq != sp) {
					bool in_range;
					const int v = pair_with_patches(sc, p, pp, pos[sq], ori + 9ll * sq, onp, onq, in_range);
					hit = mode == 0 ? v == 1 : in_range;
				}
			}
			const unsigned hits = tile.ballot(hit);
			if(mode == 0) {
				for(int patch = 0; patch < 4; patch++) { // the last neighbour in this order keeps the patch (avb.c:64)
					const unsigned on_patch = tile.ballot(hit && onp == patch);
					if(on_patch != 0u && lane == 31 - __clz((int) on_patch)) out_list[(long long) k * 4 + patch] = pid[sq];
				}
			}
			else if(hit) {
				const int at = count + __popc(hits & ((1u << lane) - 1u));
				if(at < max_nb) out_list[(long long) k * max_nb + at] = pid[sq];
			}
			count += __popc(hits);
		}
		if(lane == 0) out_count[k] = count;
	}
}

} // namespace pp

namespace {

int neighbour_query(pp_handle *h, const char *who, int mode, int n, const int *idx, int max_nb, int *counts, int *list) {
	if(h == nullptr || idx == nullptr || counts == nullptr || list == nullptr) return 1;
	if(h->slab_world > 1) return fail(h, "%s runs on a single engine", who);
	if(n <= 0) return 0;
	if(mode == 1 && max_nb < 1) return fail(h, "%s: max_per_particle must be positive", who);
	for(int k = 0; k < n; k++)
		if(idx[k] < 0 || idx[k] >= h->N) return fail(h, "%s: particle index %d out of range", who, idx[k]);
	CUDA_OK(h, cudaSetDevice(h->device));
	if(ensure_cells_valid(h)) return 1;
	const int width = mode == 0 ? 4 : max_nb;
	int *d_idx = nullptr, *d_out = nullptr;
	CUDA_OK(h, cudaMalloc(&d_idx, sizeof(int) * (size_t) n));
	CUDA_OK(h, cudaMalloc(&d_out, sizeof(int) * (size_t) n * (size_t) (width + 1)));
	CUDA_OK(h, cudaMemcpyAsync(d_idx, idx, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, h->stream));
	const SimConst sc = make_const(h, h->box, h->shift);
	constexpr int G = 8;
	int blocks = (n + kQueryThreads / G - 1) / (kQueryThreads / G);
	if(blocks > 148 * 16) blocks = 148 * 16;
	prof_begin(h);
	k_neighbour_query<G><<<blocks, kQueryThreads, 0, h->stream>>>(sc, n, d_idx, mode, max_nb, h->cb.cell_start, h->pos[h->cp], h->ori[h->co], h->pid[h->co],
			h->slot_of, d_out, d_out + n);
	prof_end(h, PP_K_PROPOSALS, 1);
	cudaMemcpyAsync(counts, d_out, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
	cudaMemcpyAsync(list, d_out + n, sizeof(int) * (size_t) n * (size_t) width, cudaMemcpyDeviceToHost, h->stream);
	int rc = 0;
	if(cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, "%s: kernel failed: %s", who, cudaGetErrorString(cudaGetLastError()));
	cudaFree(d_idx);
	cudaFree(d_out);
	return rc;
}

} // namespace

extern "C" {

int pp_patch_neighbours(pp_handle *h, int n, const int *idx, int *neighbours, int *num_neighbours) {
	return neighbour_query(h, "pp_patch_neighbours", 0, n, idx, 4, num_neighbours, neighbours);
}

int pp_range_neighbours(pp_handle *h, int n, const int *idx, int max_per_particle, int *counts, int *neighbours) {
	return neighbour_query(h, "pp_range_neighbours", 1, n, idx, max_per_particle, counts, neighbours);
}

} // extern "C"
