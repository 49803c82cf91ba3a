// pp_api.cu -- the C ABI (include/pp_b200.h): handle, device memory, and the host-side scheduling
// that replaces do_NVT / do_NPT (MC.c:19-29, 60-68), MC_change_volume (MC.c:450-513),
// MC_change_Lx (MC.c:515-585) and MC_check_energy (MC.c:766-784).  No CPU fallback: every
// numerical result comes from the kernels in pp_cells.cu / pp_sweep.cu / pp_bonds.cu.
// The x-slab decomposition lives in pp_slab_api.cu, the single-thread driver of several engines in pp_group.cu.
#include "pp_internal.h"

using namespace pp;
using namespace ppi;

namespace {

// host copy of Philox4x32-10 for the per-sweep grid shift and the box-move schedule
void philox_host(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
	for(int round = 0; round < 10; round++) {
		const uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
		const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
		c0 = n0;
		c1 = (uint32_t) p1;
		c2 = n2;
		c3 = (uint32_t) p0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0;
	out[1] = c1;
	out[2] = c2;
	out[3] = c3;
}

constexpr uint32_t kTagShift = 0xFFFFFFFFu;   // counter word 0 of the grid-shift stream
constexpr uint32_t kTagBoxMove = 0xFFFFFFFEu; // counter word 0 of the box-move stream

double unit_from_u32(uint32_t x) {
	return ((double) x + 0.5) * (1.0 / 4294967296.0);
}
double unit53(uint32_t hi, uint32_t lo) {
	return (double) ((((uint64_t) hi) << 21) | (((uint64_t) lo) >> 11)) * (1.0 / 9007199254740992.0);
}

void base_patches_host(double *base) {
	// system.c:20-31
	static const int sgn[4][3] = {{-1, -1, 1}, {1, -1, -1}, {1, 1, 1}, {-1, 1, -1}};
	const double hh = 0.5 / sqrt(3.);
	for(int i = 0; i < 4; i++) {
		double v[3];
		for(int k = 0; k < 3; k++) v[k] = sgn[i][k] > 0 ? hh : -hh;
		const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
		for(int k = 0; k < 3; k++) base[3 * i + k] = v[k] / n;
	}
}

void free_state(pp_handle *h) {
	peer_release(h); // (a neighbour's mapping of this arena must be gone first: the drivers close their connections before re-uploading)
	cudaFree(h->arena);
	h->arena = nullptr;
	h->arena_bytes = 0;
	cudaFree(h->slot_of);
	cudaFree(h->cb.cell_of);
	cudaFree(h->cb.cell_sorted);
	cudaFree(h->cb.perm);
	cudaFree(h->d_proposals);
	cudaFree(h->d_stage_r);
	cudaFree(h->d_stage_o);
	for(int k = 0; k < 2; k++) {
		cudaFree(h->pid[k]);
		cudaFree(h->outbox[k]);
		cudaFree(h->inbox[k]);
		h->outbox[k] = h->inbox[k] = nullptr;
	}
	memset(h->pos, 0, sizeof(h->pos));
	memset(h->ori, 0, sizeof(h->ori));
	memset(h->pid, 0, sizeof(h->pid));
	h->slot_of = h->cb.cell_of = h->cb.cell_sorted = h->cb.perm = nullptr;
	h->d_proposals = nullptr;
	h->cb.posf = nullptr;
	h->d_stage_r = h->d_stage_o = nullptr;
	h->cap_N = h->cap_ids = 0;
	h->cap_records = 0;
}

void free_cells(pp_handle *h) {
	cudaFree(h->cb.cell_count);
	cudaFree(h->cb.cell_fill);
	cudaFree(h->cb.cell_start);
	cudaFree(h->cb.tile_sums);
	cudaFree(h->cb.work);
	h->cb.cell_count = h->cb.cell_fill = h->cb.cell_start = h->cb.tile_sums = h->cb.work = nullptr;
	h->cap_cells = 0;
}

// n particles held at most, ids below n_ids, n_records migration / halo records per neighbour (0 without a decomposition)
int ensure_state(pp_handle *h, size_t n, size_t n_ids, size_t n_records) {
	if(n <= h->cap_N && n_ids + 32 <= h->cap_ids && n_records <= (size_t) h->cap_records) return 0;
	free_state(h);
	const size_t cap = n + 32;
	auto align = [](size_t x) { return (x + 255) / 256 * 256; };
	pp_peer_info &L = h->layout;
	memset(&L, 0, sizeof(L));
	size_t off = 0;
	for(int k = 0; k < 3; k++) {
		L.off_pos[k] = off;
		off += align(cap * sizeof(double4));
	}
	for(int k = 0; k < 2; k++) {
		L.off_ori[k] = off;
		off += align(cap * 9 * sizeof(double));
	}
	L.off_posf = off;
	off += align(cap * sizeof(uint4));
	for(int k = 0; k < 2; k++) {
		L.off_mig[k] = off;
		off += align(n_records * 13 * sizeof(double));
	}
	L.off_ctrl = off;
	off += 256;
	L.cap_records = (long long) n_records;
	CUDA_OK(h, cudaMalloc(&h->arena, off));
	CUDA_OK(h, cudaMemset(h->arena + L.off_ctrl, 0, 256));
	h->arena_bytes = off;
	for(int k = 0; k < 3; k++) h->pos[k] = reinterpret_cast<double4 *>(h->arena + L.off_pos[k]);
	for(int k = 0; k < 2; k++) h->ori[k] = reinterpret_cast<double *>(h->arena + L.off_ori[k]);
	h->cb.posf = reinterpret_cast<uint4 *>(h->arena + L.off_posf);
	for(int k = 0; k < 2; k++) {
		CUDA_OK(h, cudaMalloc(&h->pid[k], cap * sizeof(int)));
		if(n_records) {
			CUDA_OK(h, cudaMalloc(&h->outbox[k], sizeof(double) * 13 * n_records));
			CUDA_OK(h, cudaMalloc(&h->inbox[k], sizeof(double) * 13 * n_records));
		}
	}
	CUDA_OK(h, cudaMalloc(&h->slot_of, (n_ids + 32) * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.cell_of, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.perm, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.cell_sorted, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->d_proposals, cap * sizeof(Proposal)));
	CUDA_OK(h, cudaMalloc(&h->d_stage_r, cap * 3 * sizeof(double)));
	CUDA_OK(h, cudaMalloc(&h->d_stage_o, cap * 9 * sizeof(double)));
	h->cap_N = cap;
	h->cap_ids = n_ids + 32;
	h->cap_records = (int) n_records;
	return 0;
}

int ensure_cells(pp_handle *h, size_t n_cells) {
	if(n_cells <= h->cap_cells) return 0;
	free_cells(h);
	const size_t cap = n_cells + n_cells / 8 + 64;
	CUDA_OK(h, cudaMalloc(&h->cb.cell_count, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.cell_fill, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.cell_start, (cap + 1) * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.tile_sums, (cap / 2048 + 2) * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.work, cap * sizeof(int)));
	h->cap_cells = cap;
	return 0;
}

// Which configuration of the sweep kernel gets which cells.  The work list of a colour is ordered by occupancy
// (descending) and split in three ranges: cells with at most M1 = cells_max_occupancy members run on one thread per
// cell (highest throughput per instruction, longest serial chain), cells with more than M2 = cells_mid_occupancy on
// 8 lanes per cell (shortest chain), the ones in between on 4.  A small pass is bound by its longest chain, a large
// one by instruction throughput, so the thresholds grow with the number of particles an engine holds.
// lanes_per_cell = 1: one thread per cell wherever a cell fits it; 4..32: every cell on the tile kernel of that width.
// The thread-per-cell configurations need three distinct cells per axis.
void routing_limits(const pp_handle *h, int *m1, int *m2) {
	*m1 = *m2 = 0;
	const bool grid_ok = (h->slab_world > 1 || h->n_side[0] >= 4) && h->n_side[1] >= 4 && h->n_side[2] >= 4;
	if(!grid_ok) return;
	const int top = kWorkBuckets - 1;
	if(h->prm.lanes_per_cell == 1) *m1 = *m2 = top;
	else if(h->prm.cells_max_occupancy >= 0) {
		*m1 = h->prm.cells_max_occupancy < top ? h->prm.cells_max_occupancy : top;
		*m2 = h->prm.cells_mid_occupancy >= 0 ? h->prm.cells_mid_occupancy : *m1;
	}
	else if(h->prm.lanes_per_cell != 0) return; // an explicit tile width alone: tile kernel only
	else {
		// a dense system (a crystal, a compressed liquid): the mean neighbourhood does not fit the staging area of the thread-per-cell
		// configuration, nearly every cell would run on 8 lanes -- the tile kernel, which stages 64 neighbours per cell, is 2.2x faster
		// there (FCC at N = 256000: 4.8e8 against 2.2e8 attempts/s, profiles/README.md)
		if(h->n_cells > 0 && 27. * (double) h->N > (double) kCellsNbCap * (double) h->n_cells) return;
		// measured on B200 (profiles/README.md)
		*m1 = (h->N >= 800000) ? 3 : ((h->N >= 400000) ? 2 : ((h->N >= 100000) ? 1 : 0));
		*m2 = (h->N >= 800000) ? 3 : ((h->N >= 400000) ? 3 : ((h->N >= 100000) ? 2 : 0));
		if(h->prm.cells_mid_occupancy >= 0 && *m1 > 0) *m2 = h->prm.cells_mid_occupancy;
	}
	if(*m2 < *m1) *m2 = *m1;
	if(*m2 > top) *m2 = top;
	if(*m1 == 0) *m2 = 0;
}

int tile_lanes(const pp_handle *h) {
	const int l = h->prm.lanes_per_cell;
	if(l == 4 || l == 8 || l == 16 || l == 32) return l;
	return (h->N >= 400000) ? 8 : 16; // measured on B200: profiles/README.md
}

int build_cells(pp_handle *h, const SimConst &sc, const StateView &src, const StateView &dst) {
	routing_limits(h, &h->occ_limit, &h->occ_mid);
	prof_begin(h);
	const int nl = launch_cell_build(sc, h->cb, src, dst, h->slot_of, h->occ_limit, h->occ_mid, h->stream);
	prof_end(h, PP_K_CELLS, nl);
	CUDA_OK(h, cudaGetLastError());
	return 0;
}

// A trial box with a side below two cells of side r_cut cannot carry the checkerboard grid.  The reference clamps its
// grid to three cells per side there (cells.c:24-27) and evaluates the move with cells smaller than the interaction
// range; the engine counts the move as tried and rejected and leaves positions, box and cells untouched, so a
// compressing run goes on instead of aborting.
bool box_too_small(pp_handle *h, const double *nb) {
	int ns[3];
	if(even_grid(nb, h->r_cut, ns, h->prm.x_layers_multiple) == 0) return false;
	h->small_box_rejections++;
	return true;
}

} // namespace

// =================================================================================================
namespace ppi {

int fail(pp_handle *h, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	if(h) vsnprintf(h->err, sizeof(h->err), fmt, ap);
	va_end(ap);
	return 1;
}

void peer_release(pp_handle *h) {
	if(h->peer_connected && h->peer_ipc) {
		cudaIpcCloseMemHandle(h->peer_arena[0]);
		if(!h->peer_same) cudaIpcCloseMemHandle(h->peer_arena[1]);
	}
	h->peer_connected = false;
	h->peer_ipc = false;
	h->peer_arena[0] = h->peer_arena[1] = nullptr;
	h->peer_handle[0] = h->peer_handle[1] = nullptr;
	memset(&h->hm, 0, sizeof(h->hm));
}

// even grid with cells of side >= r_cut (cells.c:22-28 uses floor(box / r_cut), any parity)
// the 1e-9 margin keeps non-adjacent cells more than r_cut apart whatever the rounding of the cell index
// x_multiple (pp_params.x_layers_multiple, even; 0 or 2: none): the number of x layers is rounded down to a multiple of it -- 2 n for a
// run in n slabs gives every slab the same number of layers.  Any even grid with cells of side >= r_cut finds the same pairs.
int even_grid(const double *box, double r_cut, int *n_side, int x_multiple) {
	for(int k = 0; k < 3; k++) {
		int n = (int) floor(box[k] / r_cut);
		n &= ~1;
		if(n >= 2 && box[k] / n < r_cut * (1. + 1e-9)) n -= 2;
		if(k == 0 && x_multiple > 2 && n >= x_multiple) n -= n % x_multiple; // (a box too small for one period of it keeps the plain even grid)
		if(n < 2) return 1;
		n_side[k] = n;
	}
	return 0;
}

// host twin of pp::cell_coord (pp_common.cuh): same operations, same roundings (no contraction on the host: the
// expression has no multiply-add pair)
int cell_coord_host(double r, double shift, double inv_box, int n_side) {
	const double t = (r - shift) * inv_box;
	const double f = t - floor(t);
	const int c = (int) (f * (double) n_side);
	return c < n_side ? c : n_side - 1;
}

void slab_partition(int n_layers, int rank, int world, int *x0, int *x1) {
	// whole pairs of x layers per engine: every engine's first owned layer is even, its last one odd
	const int pairs = n_layers / 2, base = pairs / world, rem = pairs % world;
	*x0 = 2 * (rank * base + (rank < rem ? rank : rem));
	*x1 = *x0 + 2 * (base + (rank < rem ? 1 : 0));
}

SimConst make_const(const pp_handle *h, const double *box, const double *shift) {
	SimConst sc;
	memset(&sc, 0, sizeof(sc));
	for(int k = 0; k < 3; k++) {
		sc.box[k] = box[k];
		sc.shift[k] = shift[k];
		sc.n_side[k] = h->n_side[k];
		sc.cell_w[k] = box[k] / h->n_side[k];
		sc.inv_box[k] = 1. / box[k];
		sc.inv_w[k] = (double) h->n_side[k] * sc.inv_box[k];
	}
	sc.n_cells = h->n_cells;
	sc.N = h->N;
	if(h->slab_world > 1) {
		sc.slab = 1;
		sc.nxl = h->nxl;
		sc.x_lo = (h->x0 - 1 + h->n_side[0]) % h->n_side[0];
		sc.own_lo = 1;
		sc.own_hi = h->nxl - 1;
	}
	else {
		sc.slab = 0;
		sc.nxl = h->n_side[0];
		sc.x_lo = 0;
		sc.own_lo = 0;
		sc.own_hi = h->n_side[0];
	}
	sc.sqr_rcut = (1. + h->prm.kf_delta) * (1. + h->prm.kf_delta); // system.c:86
	sc.cosmax = h->prm.kf_cosmax;
	sc.cos2_lo = sc.cosmax * sc.cosmax * (1. - 1e-11);
	sc.cos2_hi = sc.cosmax * sc.cosmax * (1. + 1e-11);
	sc.inv_b2 = 1. / (h->base[2] * h->base[2]);
	sc.fast_patch = sc.cosmax > 1e-3;
	memcpy(sc.base, h->base, sizeof(sc.base));
	for(int k = 0; k < kMaxExp; k++) sc.exp_table[k] = exp(-(double) k / h->prm.temperature); // MC.c:332
	sc.disp_max = h->prm.disp_max;
	sc.theta_max = h->prm.theta_max;
	sc.seed_lo = (uint32_t) h->prm.seed;
	sc.seed_hi = (uint32_t) (h->prm.seed >> 32);
	sc.sweep_lo = (uint32_t) h->step;
	sc.sweep_hi8 = (uint32_t) (((uint64_t) h->step >> 32) << 8);
	return sc;
}

StateView view(pp_handle *h, int p, int o) {
	StateView v;
	v.pos = h->pos[p];
	v.ori = h->ori[o];
	v.pid = h->pid[o];
	return v;
}

// event-timed section helpers (only when profiling)
void prof_begin(pp_handle *h) {
	if(h->profiling) cudaEventRecord(h->ev0, h->stream);
}
void prof_end(pp_handle *h, int kclass, int n_launches) {
	h->launches[kclass] += n_launches;
	h->total_launches += n_launches;
	if(h->profiling) {
		cudaEventRecord(h->ev1, h->stream);
		cudaEventSynchronize(h->ev1);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, h->ev0, h->ev1);
		h->ms[kclass] += ms;
	}
}

// the colour passes `first_pass .. first_pass + n_pass - 1` of one sweep (pass order: x parity slowest): one launch each
int launch_passes(pp_handle *h, const SimConst &sc, int first_pass, int n_pass) {
	const int limit = h->occ_limit, lanes = tile_lanes(h); // the split the work lists were built with, not a fresh evaluation
	long long n_active = (long long) (sc.nxl / 2 + 1) * (sc.n_side[1] / 2) * (sc.n_side[2] / 2);
	if(n_active > sc.N) n_active = sc.N;
	const int cl = h->prm.cell_lanes;
	int nl = 0;
	for(int pass = first_pass; pass < first_pass + n_pass; pass++) {
		const int colour = ((pass >> 2) & 1) | (((pass >> 1) & 1) << 1) | ((pass & 1) << 2);
		const int *ranges = h->cb.bins + kWorkRanges + 6 * colour;
		if(limit == 0) { // every cell on the tile kernel (the three ranges are contiguous: the first one holds everything)
			nl += launch_sweep_pass(sc, ranges, lanes, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->hm, h->stream);
		}
		else if(cl == 2 || cl == 4 || cl == 8 || cl == -1) { // an explicit lane count for the sparse range (-1: one thread per cell): the three configurations one after the other
			nl += launch_sweep_cells(sc, ranges, n_active, 8, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->hm, h->stream);
			nl += launch_sweep_cells(sc, ranges + 2, n_active, 4, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->hm, h->stream);
			nl += launch_sweep_cells(sc, ranges + 4, n_active, cl < 0 ? 1 : cl, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->hm, h->stream);
		}
		else nl += launch_sweep_fused(sc, ranges, n_active, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->hm, h->stream);
	}
	return nl;
}

// slab decomposition: after a sort, particles outside the held layers sit in the trash cell at the end of the
// arrays and are dropped; the slot ranges of the boundary and halo layers are cached for the halo mirror
int slab_after_sort(pp_handle *h) {
	const int nyz = h->n_side[1] * h->n_side[2];
	const int idx[7] = {0, nyz, 2 * nyz, (h->nxl - 2) * nyz, (h->nxl - 1) * nyz, h->nxl * nyz, h->n_cells + 1};
	int *val = h->h_slab_ints; // one gather kernel + one copy into pinned memory
	launch_gather_ints(h->cb.cell_start, idx, 7, h->d_slab_counts + 8, h->stream);
	CUDA_OK(h, cudaMemcpyAsync(val, h->d_slab_counts + 8, 7 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	for(int k = 0; k < 6; k++) h->layer_slots[k] = val[k];
	h->N = val[5]; // everything before the trash cell
	return 0;
}

// the cells of the merged / current set on the grid with the given shift: sorted copy from `src` into `dst`
int sort_into(pp_handle *h, const double *shift, const StateView &src, const StateView &dst) {
	memcpy(h->shift, shift, sizeof(h->shift));
	const SimConst sc = make_const(h, h->box, h->shift);
	return build_cells(h, sc, src, dst);
}

// K1 on the current state with the given grid shift; the sorted copy becomes current.
int rebin(pp_handle *h, const double *shift) {
	if(h->slab_world <= 1) {
		if(even_grid(h->box, h->r_cut, h->n_side, h->prm.x_layers_multiple)) return fail(h, "box (%g, %g, %g) is smaller than two cells of side r_cut = %g", h->box[0], h->box[1], h->box[2], h->r_cut);
		h->nxl = h->n_side[0];
	}
	h->n_cells = h->nxl * h->n_side[1] * h->n_side[2];
	if(ensure_cells(h, (size_t) h->n_cells + 2)) return 1;
	const int np = (h->cp + 1) % 3, no = h->co ^ 1;
	if(sort_into(h, shift, view(h, h->cp, h->co), view(h, np, no))) return 1;
	h->cp = np;
	h->co = no;
	h->cells_valid = true;
	h->sweep_grid = false;
	if(h->slab_world > 1) {
		if(slab_after_sort(h)) return 1;
		if(h->peer_connected && slab_build_mirror(h)) return 1;
	}
	return 0;
}

int ensure_cells_valid(pp_handle *h) {
	if(h->cells_valid) return 0;
	const double zero[3] = {0., 0., 0.};
	return rebin(h, zero);
}

int fetch_counters(pp_handle *h) {
	CUDA_OK(h, cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

// K3 on (pos buffer p, current orientations) for the given box / shift
int count_bonds(pp_handle *h, int p, const double *box, const double *shift, long long *bonds, int *overlap) {
	const SimConst sc = make_const(h, box, shift);
	prof_begin(h);
	// the proposals of a sweep are dead between sweeps: their buffer holds the pair queue of the streaming count
	size_t scratch_bytes = (size_t) h->cap_N * sizeof(Proposal);
	if(const char *limit = getenv("PP_BONDS_QUEUE_BYTES")) { // test hook: a small queue exercises the overflow path below
		const long long b = atoll(limit);
		if(b >= 0 && (size_t) b < scratch_bytes) scratch_bytes = (size_t) b;
	}
	const int nl = launch_total_bonds(sc, h->cb.cell_start, h->cb.cell_sorted, h->pos[p], h->ori[h->co], h->cb.posf, h->d_counters, h->d_proposals,
			scratch_bytes, false, h->stream);
	prof_end(h, PP_K_BONDS, nl);
	CUDA_OK(h, cudaGetLastError());
	if(fetch_counters(h)) return 1;
	if(h->h_counters->pad != 0) { // more candidate pairs than the queue holds (a very dense, strongly bonded state): the exact tile kernel
		prof_begin(h);
		const int nt = launch_total_bonds(sc, h->cb.cell_start, h->cb.cell_sorted, h->pos[p], h->ori[h->co], h->cb.posf, h->d_counters, nullptr, 0, true,
				h->stream);
		prof_end(h, PP_K_BONDS, nt);
		CUDA_OK(h, cudaGetLastError());
		if(fetch_counters(h)) return 1;
	}
	*bonds = h->h_counters->bonds;
	*overlap = h->h_counters->overlap;
	return 0;
}

long long running_bonds(pp_handle *h) {
	return h->bonds_base + h->h_counters->dbonds;
}

// the box-move stream: uniforms keyed by (seed, draw number); the same on a single engine and on every engine of a group
double box_uniform(uint64_t seed, uint64_t draw) {
	uint32_t x[4];
	philox_host(kTagBoxMove, 0u, (uint32_t) draw, (uint32_t) (draw >> 32), (uint32_t) seed, (uint32_t) (seed >> 32), x);
	return unit53(x[0], x[1]);
}

// number of successes of N Bernoulli(1/N) draws (the per-iteration test of MC.c:24-27, 62-63), by inversion of u
int binomial_one_over_N(int N, double u) {
	if(N <= 0) return 0;
	if(N == 1) return 1;
	const double p = 1. / N, q = 1. - p;
	double pk = pow(q, (double) N), cdf = pk;
	int k = 0;
	while(u >= cdf && k < N) {
		pk *= ((double) (N - k) / (double) (k + 1)) * (p / q);
		k++;
		cdf += pk;
		if(pk < 1e-300) break;
	}
	return k;
}

// random grid shift in [0, cell width) of shift period `epoch`
void sweep_shift(uint64_t seed, uint64_t epoch, const double *box, const int *n_side, double *shift) {
	uint32_t x[4];
	philox_host(kTagShift, 0u, (uint32_t) epoch, (uint32_t) ((epoch >> 32) << 8), (uint32_t) seed, (uint32_t) (seed >> 32), x);
	for(int k = 0; k < 3; k++) shift[k] = unit_from_u32(x[k]) * (box[k] / n_side[k]);
}

// ---- upload ---------------------------------------------------------------------------------------------------------
// first part of every upload: box, grid, slab partition, books
static int upload_begin(pp_handle *h, int N_global, const double *box, bool keep_books) {
	CUDA_OK(h, cudaSetDevice(h->device));
	for(int k = 0; k < 3; k++)
		if(!(box[k] > 0.)) return fail(h, "box side %d is not positive", k);
	if(keep_books) { // fold what the device counters hold into the carry: the upload zeroes them
		if(fetch_counters(h)) return 1;
		const Counters &c = *h->h_counters;
		h->carry.tries += c.tries;
		h->carry.accepted += c.accepted;
		h->carry.confined += c.confined;
		h->carry.overlapped += c.overlapped;
		h->carry.w_tries += c.w_tries;
		h->carry.w_accepted += c.w_accepted;
		h->carry.pair_checks += c.pair_checks;
	}
	else memset(&h->carry, 0, sizeof(h->carry));
	h->N_global = N_global;
	memcpy(h->box, box, sizeof(h->box));
	if(!keep_books) h->V = box[0] * box[1] * box[2];
	if(even_grid(h->box, h->r_cut, h->n_side, h->prm.x_layers_multiple)) return fail(h, "box (%g, %g, %g) is smaller than two cells of side r_cut = %g", box[0], box[1], box[2], h->r_cut);
	h->nxl = h->n_side[0];
	if(h->slab_world > 1) {
		if(h->n_side[0] / 2 < h->slab_world) return fail(h, "%d x layers cannot be split over %d engines (at least 2 layers each)", h->n_side[0], h->slab_world);
		slab_partition(h->n_side[0], h->slab_rank, h->slab_world, &h->x0, &h->x1);
		h->nxl = (h->x1 - h->x0) + 2;
		if(h->d_slab_counts == nullptr) CUDA_OK(h, cudaMalloc(&h->d_slab_counts, 16 * sizeof(int)));
		if(h->h_slab_ints == nullptr) CUDA_OK(h, cudaMallocHost(&h->h_slab_ints, 8 * sizeof(int)));
		h->slab_pending = false;
	}
	return 0;
}

// array sizes of a slab engine from the number of particles it receives and the fullest x layer among them (an
// inhomogeneous system -- a slab of liquid, a gel -- has layers far above the mean)
static void slab_sizes(int n, int fullest, size_t *cap, size_t *n_records) {
	*n_records = 3 * (size_t) fullest + 8192;
	*cap = (size_t) n + (size_t) n / 8 + 4 * (size_t) fullest + 8192;
}

// last part: the state sits unsorted in buffer 0 with n particles; reset the books, build the cells
static int upload_finish(pp_handle *h, int n, bool check_held, bool keep_books) {
	h->N = n;
	h->cp = 0;
	h->co = 0;
	CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(Counters), h->stream));
	CUDA_OK(h, cudaMemsetAsync(h->arena + h->layout.off_ctrl, 0, 256, h->stream));
	memset(h->h_counters, 0, sizeof(Counters));
	h->cells_valid = false;
	h->sweep_grid = false;
	h->phase = 0;
	h->mig_epoch = 0;
	h->rebuilt_step = -1;
	if(!keep_books) {
		h->step = 0;
		h->bonds_base = 0;
		h->bonds_known = false;
		h->box_draws = 0;
		memset(h->tries, 0, sizeof(h->tries));
		memset(h->accepted, 0, sizeof(h->accepted));
	}
	memset(&h->bm, 0, sizeof(h->bm));
	if(h->slab_world > 1) {
		if(ensure_cells_valid(h)) return 1; // (a slab engine keeps only the particles of its layers: the sort tells which)
	}
	else {
		// single engine: the grid is fixed here (a box too small for it is refused now), the counting sort waits for the first
		// consumer -- a sweep re-sorts on its shifted grid anyway, so an upload followed by a sweep builds the cells once, not twice
		if(even_grid(h->box, h->r_cut, h->n_side, h->prm.x_layers_multiple)) return fail(h, "box (%g, %g, %g) is smaller than two cells of side r_cut = %g", h->box[0], h->box[1], h->box[2], h->r_cut);
		h->nxl = h->n_side[0];
		h->n_cells = h->nxl * h->n_side[1] * h->n_side[2];
	}
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(h->slab_world > 1 && check_held && h->N != n)
		return fail(h, "upload: %d of the %d particles lie outside the layers this engine holds", n - h->N, n);
	return 0;
}

int upload_impl(pp_handle *h, int N_global, int n, const int *ids, const double *box, const double *r, const double *orient, bool keep_books) {
	if(n < 0 || N_global < n) return fail(h, "bad particle numbers (%d local of %d)", n, N_global);
	if(upload_begin(h, N_global, box, keep_books)) return 1;
	size_t cap = (size_t) n, n_records = 0, n_ids = (size_t) N_global;
	if(h->slab_world <= 1 && (size_t) h->reserve_N > cap) cap = n_ids = (size_t) h->reserve_N; // room for insertions (pp_reserve)
	if(cap < 4) cap = 4;
	if(h->slab_world > 1) {
		const int nx = h->n_side[0];
		int *layer_count = (int *) calloc((size_t) nx, sizeof(int));
		if(layer_count == nullptr) return fail(h, "out of host memory");
		const double inv_box = 1. / h->box[0];
		for(int i = 0; i < n; i++) layer_count[cell_coord_host(r[3ll * i], 0., inv_box, nx)]++;
		int fullest = 0;
		for(int k = 0; k < nx; k++) fullest = layer_count[k] > fullest ? layer_count[k] : fullest;
		free(layer_count);
		slab_sizes(n, fullest, &cap, &n_records);
	}
	if(ensure_state(h, cap, n_ids, n_records)) return 1;
	peer_release(h); // a new configuration: the driver connects the neighbours again
	if(n > 0) {
		CUDA_OK(h, cudaMemcpyAsync(h->d_stage_r, r, sizeof(double) * 3 * (size_t) n, cudaMemcpyHostToDevice, h->stream));
		CUDA_OK(h, cudaMemcpyAsync(h->d_stage_o, orient, sizeof(double) * 9 * (size_t) n, cudaMemcpyHostToDevice, h->stream));
		if(ids) CUDA_OK(h, cudaMemcpyAsync(h->cb.perm, ids, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, h->stream));
	}
	launch_pack_state(n, h->d_stage_r, h->d_stage_o, ids ? h->cb.perm : nullptr, h->pos[0], h->ori[0], h->pid[0], h->stream);
	return upload_finish(h, n, ids != nullptr, keep_books);
}

// A slab engine handed the WHOLE configuration in pinned (device-readable) host memory takes the particles of its own
// layers itself: one kernel reads the positions over PCIe, selects by x layer with the engine's cell arithmetic, and
// gathers positions and orientations of the selected particles straight into the device arrays -- no host-side
// selection, no staging copy, an eighth of the orientations crosses the bus on eight engines.
int upload_select(pp_handle *h, int N, const double *box, const double *r, const double *orient) {
	if(upload_begin(h, N, box, false)) return 1;
	const int nx = h->n_side[0];
	const int lo = (h->x0 - 1 + nx) % nx, held = h->x1 - h->x0 + 2;
	int *d_hist = nullptr; // [0] particles selected, [1] overflow, [2 .. 2 + nx) particles per global x layer
	CUDA_OK(h, cudaMalloc(&d_hist, sizeof(int) * (size_t) (nx + 2)));
	int *hist = (int *) malloc(sizeof(int) * (size_t) (nx + 2));
	int rc = 0, n_sel = 0;
	for(int attempt = 0; attempt < 2 && rc == 0; attempt++) {
		// with arrays from an earlier upload the selection goes straight in; if they turn out too small (or there are none
		// yet) the first pass only counts and the second one packs
		const bool pack = h->cap_N > 0 && (size_t) N + 32 <= h->cap_ids;
		cudaMemsetAsync(d_hist, 0, sizeof(int) * (size_t) (nx + 2), h->stream);
		if(pack) peer_release(h);
		launch_select_pack(N, r, orient, 1. / h->box[0], nx, lo, held, pack ? h->pos[0] : nullptr, h->ori[0], h->pid[0], (int) h->cap_N, d_hist, h->stream);
		if(cudaMemcpyAsync(hist, d_hist, sizeof(int) * (size_t) (nx + 2), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
				cudaStreamSynchronize(h->stream) != cudaSuccess) {
			rc = fail(h, "upload: selection kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
			break;
		}
		n_sel = hist[0];
		int fullest = 0;
		for(int k = 0; k < held; k++) {
			const int c = hist[2 + (lo + k) % nx];
			fullest = c > fullest ? c : fullest;
		}
		size_t cap, n_records;
		slab_sizes(n_sel, fullest, &cap, &n_records);
		const bool fits = pack && !hist[1] && (size_t) n_sel + 32 <= h->cap_N && n_records <= (size_t) h->cap_records;
		if(fits) break;
		if(attempt == 1) {
			rc = fail(h, "upload: the selection does not fit the arrays sized for it (%d particles)", n_sel);
			break;
		}
		if(ensure_state(h, cap > h->cap_N ? cap : h->cap_N, (size_t) N, n_records > (size_t) h->cap_records ? n_records : (size_t) h->cap_records)) rc = 1;
	}
	free(hist);
	cudaFree(d_hist);
	if(rc) return rc;
	return upload_finish(h, n_sel, true, false);
}

int sweep_once(pp_handle *h) {
	const SimConst sc = make_const(h, h->box, h->shift);
	prof_begin(h);
	const int nr = launch_proposals_for_sweep(sc, h->cb, h->d_proposals, h->stream);
	prof_end(h, PP_K_RNG, nr);
	prof_begin(h);
	const int nl = launch_passes(h, sc, 0, 8);
	prof_end(h, PP_K_SWEEP, nl);
	return 0;
}

// ---- box moves ---------------------------------------------------------------------------------------------------
// propose: the trial box from the move's uniforms, in the reference's draw order.  *status = -1: rejected before any
// work by the Lyz window (MC.c:527-530), -2: the trial box is too small for the grid, 0: go on with box_trial_local.
int box_propose(pp_handle *h, int kind, const double *u, int *status) {
	BoxMove &bm = h->bm;
	memset(&bm, 0, sizeof(bm));
	bm.kind = kind;
	*status = 0;
	memcpy(bm.old_box, h->box, sizeof(bm.old_box));
	memcpy(bm.nb, h->box, sizeof(bm.nb));
	bm.f[0] = bm.f[1] = bm.f[2] = 1.;
	if(kind == PP_VOLUME) { // MC.c:450-472
		h->tries[PP_VOLUME]++;
		bm.initial_V = h->box[0] * h->box[1] * h->box[2];
		const int dir = (int) (u[0] * 3);
		const double ln_final_V = log(bm.initial_V) + (u[1] - 0.5) * h->prm.rescale_factor_max;
		bm.final_V = exp(ln_final_V);
		const double old_side = h->box[dir];
		double new_side = bm.final_V;
		for(int k = 0; k < 3; k++)
			if(k != dir) new_side /= h->box[k];
		bm.rescale = new_side / old_side;
		bm.nb[dir] = new_side;
		bm.f[dir] = bm.rescale;
		bm.u_acc = u[2];
	}
	else { // MC.c:515-538
		h->tries[PP_LX]++;
		bm.nb[0] = h->box[0] + (u[0] - 0.5) * h->prm.Lx_change_max;
		bm.nb[1] = bm.nb[2] = sqrt(h->V / bm.nb[0]); // MC.c:524: the stored volume, not the product of the current sides
		if(bm.nb[1] < h->prm.Lyz_min || bm.nb[1] > h->prm.Lyz_max) {
			*status = -1;
			return 0;
		}
		for(int k = 0; k < 3; k++) bm.f[k] = bm.nb[k] / bm.old_box[k];
		bm.u_acc = u[1];
	}
	if(box_too_small(h, bm.nb)) { // counted, rejected, nothing touched
		*status = -2;
		return 0;
	}
	bm.pending = true;
	return 0;
}

// trial: rescale into the trial buffer (K4) and recount (K3) on the engine's own layers.  *status = 2: the trial box
// needs another cell grid (a single engine re-sorts by itself; a decomposition gathers and uploads again).
int box_trial_local(pp_handle *h, long long *bonds_local, int *overlap_local, int *status) {
	BoxMove &bm = h->bm;
	*status = 0;
	if(!bm.pending) return fail(h, "box_trial_local without a proposed move");
	if(!h->bonds_known) return fail(h, "box move before the initial pp_total_bonds");
	if(ensure_cells_valid(h)) return 1;
	int ns[3];
	bool regrid = even_grid(bm.nb, h->r_cut, ns, h->prm.x_layers_multiple) != 0;
	for(int k = 0; k < 3 && !regrid; k++) regrid = (bm.nb[k] / h->n_side[k] < h->r_cut * (1. + 1e-9));
	if(regrid && h->slab_world > 1) {
		*status = 2;
		return 0;
	}
	const int tp = (h->cp + 2) % 3;
	prof_begin(h);
	const int nl = launch_rescale(h->N, h->pos[h->cp], h->pos[tp], bm.f[0], bm.f[1], bm.f[2], 0, h->stream);
	prof_end(h, PP_K_RESCALE, nl);
	if(regrid) { // single engine: make the scaled state current, re-sort it on a fresh grid and count
		*status = 2;
		bm.slow = true;
		h->cp = tp;
		memcpy(h->box, bm.nb, sizeof(h->box));
		const double zero[3] = {0., 0., 0.};
		if(rebin(h, zero)) return 1;
		return count_bonds(h, h->cp, h->box, h->shift, bonds_local, overlap_local);
	}
	const double sh[3] = {h->shift[0] * bm.f[0], h->shift[1] * bm.f[1], h->shift[2] * bm.f[2]};
	return count_bonds(h, tp, bm.nb, sh, bonds_local, overlap_local);
}

// the reference's acceptance tests (MC.c:497-498, 562), on the totals of the whole system
bool box_decide(pp_handle *h, long long bonds_new_total, int overlap, long long bonds_old_total, int N_total) {
	const BoxMove &bm = h->bm;
	const double delta_E = -(double) bonds_new_total + (double) bonds_old_total;
	double arg;
	if(bm.kind == PP_VOLUME) {
		const double delta_V = bm.final_V - bm.initial_V;
		arg = -(h->prm.P * delta_V + delta_E) / h->prm.temperature + (N_total + 1) * log(bm.rescale);
	}
	else arg = h->prm.Lx_reference_sign ? (delta_E / h->prm.temperature) : (-delta_E / h->prm.temperature);
	return !overlap && bm.u_acc < exp(arg);
}

int box_finish(pp_handle *h, bool acc, long long bonds_local_new) {
	BoxMove &bm = h->bm;
	if(!bm.pending) return fail(h, "box_finish without a proposed move");
	bm.pending = false;
	const int slot = bm.kind == PP_VOLUME ? PP_VOLUME : PP_LX;
	if(acc) {
		if(!bm.slow) {
			h->cp = (h->cp + 2) % 3;
			memcpy(h->box, bm.nb, sizeof(h->box));
			for(int k = 0; k < 3; k++) h->shift[k] *= bm.f[k];
			if(h->peer_connected && h->peer_ipc && slab_build_mirror(h)) return 1; // the neighbours switched buffers too (a group rebuilds its mirrors itself)
		}
		if(bm.kind == PP_VOLUME) h->V = bm.final_V; // MC.c:500
		h->bonds_base = bonds_local_new - h->h_counters->dbonds;
		h->accepted[slot]++;
		h->sweep_grid = false;
	}
	else if(bm.slow) {
		// go back: divide the coordinates back like the reference and rebuild the cells for the old box (MC.c:505-508, 567-581)
		const int tp = (h->cp + 2) % 3;
		prof_begin(h);
		const int nl = launch_rescale(h->N, h->pos[h->cp], h->pos[tp], bm.f[0], bm.f[1], bm.f[2], 1, h->stream);
		prof_end(h, PP_K_RESCALE, nl);
		h->cp = tp;
		memcpy(h->box, bm.old_box, sizeof(h->box));
		const double zero[3] = {0., 0., 0.};
		if(rebin(h, zero)) return 1;
	}
	return 0;
}

} // namespace ppi

namespace {

// one box move on a single engine; *accepted: 1, 0, -1 (Lyz window)
int box_move_single(pp_handle *h, int kind, const double *u, int *accepted) {
	if(fetch_counters(h)) return 1;
	int status = 0;
	if(box_propose(h, kind, u, &status)) return 1;
	if(status < 0) {
		*accepted = status == -1 ? -1 : 0;
		return 0;
	}
	const long long bonds_old = running_bonds(h);
	long long bonds_new = 0;
	int overlap = 0;
	if(box_trial_local(h, &bonds_new, &overlap, &status)) return 1;
	const bool acc = box_decide(h, bonds_new, overlap, bonds_old, h->N);
	if(box_finish(h, acc, bonds_new)) return 1;
	*accepted = acc ? 1 : 0;
	return 0;
}

} // namespace

// =================================================================================================
extern "C" {

int pp_abi_version(void) {
	return PP_ABI_VERSION;
}

void pp_default_params(pp_params *p) {
	memset(p, 0, sizeof(*p));
	p->ensemble = PP_NVT;
	p->temperature = 0.1;
	p->disp_max = 0.1;
	p->theta_max = 0.1;
	p->kf_delta = 0.119;
	p->kf_cosmax = 0.98;
	p->Lx_reference_sign = 1;
	p->seed = 12345;
	p->lanes_per_cell = 0;
	p->cells_max_occupancy = -1;
	p->cells_mid_occupancy = -1;
	p->x_layers_multiple = 0;
}

const char *pp_last_error(const pp_handle *h) {
	return h ? h->err : "null handle";
}

int pp_create(const pp_params *params, pp_handle **out) {
	if(out == nullptr || params == nullptr) return 1;
	*out = nullptr;
	pp_handle *h = (pp_handle *) calloc(1, sizeof(pp_handle));
	if(h == nullptr) return 1;
	h->prm = *params;
	h->slab_world = 1;
	*out = h; // returned even on failure so that the caller can read pp_last_error
	if(params->ensemble != PP_NVT && params->ensemble != PP_NPT) return fail(h, "Ensemble %d not supported on the GPU path (0 = NVT, 5 = NPT)", params->ensemble);
	if(params->ensemble == PP_NPT && params->Lx_move) return fail(h, "Lx_move is only available with Ensemble = 0 (as in the reference, system.c:165-171)");
	if(params->x_layers_multiple < 0 || (params->x_layers_multiple & 1)) return fail(h, "x_layers_multiple must be an even number (0: any even grid)");
	if(!(params->temperature > 0.)) return fail(h, "Temperature must be positive");
	int n_dev = 0;
	cudaError_t e = cudaGetDeviceCount(&n_dev);
	if(e != cudaSuccess || n_dev == 0) return fail(h, "no CUDA device available (%s): this engine has no CPU fallback", cudaGetErrorString(e));
	if(params->device < 0 || params->device >= n_dev) return fail(h, "device %d out of range (%d visible)", params->device, n_dev);
	h->device = params->device;
	CUDA_OK(h, cudaSetDevice(h->device));
	cudaDeviceProp prop;
	CUDA_OK(h, cudaGetDeviceProperties(&prop, h->device));
	if(prop.major < 10) return fail(h, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", h->device, prop.major, prop.minor);
	CUDA_OK(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	h->own_stream = true;
	CUDA_OK(h, cudaMalloc(&h->d_counters, sizeof(Counters)));
	CUDA_OK(h, cudaMemset(h->d_counters, 0, sizeof(Counters)));
	CUDA_OK(h, cudaMallocHost(&h->h_counters, sizeof(Counters)));
	memset(h->h_counters, 0, sizeof(Counters));
	CUDA_OK(h, cudaMalloc(&h->cb.bins, sizeof(int) * kBinInts));
	CUDA_OK(h, cudaMemset(h->cb.bins, 0, sizeof(int) * kBinInts));
	CUDA_OK(h, cudaMalloc(&h->d_push_done, sizeof(unsigned int)));
	CUDA_OK(h, cudaMemset(h->d_push_done, 0, sizeof(unsigned int)));
	CUDA_OK(h, cudaMalloc(&h->d_peer_status, sizeof(int)));
	CUDA_OK(h, cudaMemset(h->d_peer_status, 0, sizeof(int)));
	CUDA_OK(h, cudaEventCreate(&h->ev0));
	CUDA_OK(h, cudaEventCreate(&h->ev1));
	for(int k = 0; k < 2; k++) CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_phase[k], cudaEventDisableTiming));
	CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_mig, cudaEventDisableTiming));
	h->r_cut = 1. + params->kf_delta; // system.c:249
	base_patches_host(h->base);
	if(h->prm.lanes_per_cell != 1 && h->prm.lanes_per_cell != 4 && h->prm.lanes_per_cell != 8 && h->prm.lanes_per_cell != 16 && h->prm.lanes_per_cell != 32)
		h->prm.lanes_per_cell = 0; // automatic
	if(h->prm.shift_every < 1) h->prm.shift_every = 4;
	return 0;
}

int pp_destroy(pp_handle *h) {
	if(h == nullptr) return 0;
	if(h->stream) {
		cudaSetDevice(h->device);
		cudaStreamSynchronize(h->stream);
	}
	free_state(h);
	free_cells(h);
	cudaFree(h->d_counters);
	cudaFree(h->cb.bins);
	cudaFree(h->d_slab_counts);
	cudaFree(h->d_gc_out);
	cudaFreeHost(h->h_slab_ints);
	cudaFree(h->d_push_done);
	cudaFree(h->d_peer_status);
	cudaFreeHost(h->h_counters);
	if(h->ev0) cudaEventDestroy(h->ev0);
	if(h->ev1) cudaEventDestroy(h->ev1);
	for(int k = 0; k < 2; k++)
		if(h->ev_phase[k]) cudaEventDestroy(h->ev_phase[k]);
	if(h->ev_mig) cudaEventDestroy(h->ev_mig);
	if(h->own_stream && h->stream) cudaStreamDestroy(h->stream);
	free(h);
	return 0;
}

int pp_set_stream(pp_handle *h, void *cuda_stream) {
	if(h == nullptr) return 1;
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(h->own_stream) cudaStreamDestroy(h->stream);
	h->stream = (cudaStream_t) cuda_stream;
	h->own_stream = false;
	return 0;
}

int pp_set_profiling(pp_handle *h, int enable) {
	if(h == nullptr) return 1;
	h->profiling = enable != 0;
	return 0;
}

int pp_kernel_stats(pp_handle *h, int kclass, long long *launches, double *ms_total) {
	if(h == nullptr || kclass < 0 || kclass >= PP_K_CLASSES) return 1;
	if(launches) *launches = h->launches[kclass];
	if(ms_total) *ms_total = h->ms[kclass];
	return 0;
}

int pp_reset_kernel_stats(pp_handle *h) {
	if(h == nullptr) return 1;
	memset(h->launches, 0, sizeof(h->launches));
	memset(h->ms, 0, sizeof(h->ms));
	h->total_launches = 0;
	return 0;
}

int pp_upload(pp_handle *h, int N, const double *box, const double *r, const double *orient) {
	if(h == nullptr) return 1;
	if(N < 0) return fail(h, "negative particle number");
	return upload_impl(h, N, N, nullptr, box, r, orient, false);
}

int pp_upload_subset(pp_handle *h, int N_global, int n_local, const int *ids, const double *box, const double *r, const double *orient) {
	if(h == nullptr) return 1;
	if(ids == nullptr && n_local > 0) return fail(h, "pp_upload_subset needs the global ids of the particles");
	for(int i = 0; i < n_local; i++)
		if(ids[i] < 0 || ids[i] >= N_global) return fail(h, "pp_upload_subset: id %d out of range", ids[i]);
	return upload_impl(h, N_global, n_local, ids, box, r, orient, false);
}

int pp_upload_select(pp_handle *h, int N, const double *box, const double *r, const double *orient) {
	if(h == nullptr) return 1;
	if(N < 0) return fail(h, "negative particle number");
	if(h->slab_world <= 1) return upload_impl(h, N, N, nullptr, box, r, orient, false);
	// can the device read the caller's arrays (pinned host memory under unified addressing)?
	cudaPointerAttributes ar, ao;
	const bool readable = N > 0 && cudaPointerGetAttributes(&ar, r) == cudaSuccess && cudaPointerGetAttributes(&ao, orient) == cudaSuccess &&
			(ar.type == cudaMemoryTypeHost || ar.type == cudaMemoryTypeManaged) && (ao.type == cudaMemoryTypeHost || ao.type == cudaMemoryTypeManaged);
	cudaGetLastError();
	if(readable) return upload_select(h, N, box, r, orient);
	// pageable memory: select and pack on the host
	unsigned char *mine = (unsigned char *) malloc((size_t) N + 1);
	if(mine == nullptr) return fail(h, "out of host memory");
	if(pp_slab_select(&h->prm, h->slab_rank, h->slab_world, N, box, r, mine)) {
		free(mine);
		return fail(h, "the box (%g, %g, %g) cannot be split into %d slabs of at least two cell layers", box[0], box[1], box[2], h->slab_world);
	}
	size_t m = 0;
	for(int i = 0; i < N; i++) m += mine[i];
	int *ids = (int *) malloc(sizeof(int) * (m + 1));
	double *rs = (double *) malloc(sizeof(double) * 3 * (m + 1)), *os = (double *) malloc(sizeof(double) * 9 * (m + 1));
	int rc = 0;
	if(!ids || !rs || !os) rc = fail(h, "out of host memory");
	else {
		size_t k = 0;
		for(int i = 0; i < N; i++)
			if(mine[i]) {
				ids[k] = i;
				memcpy(rs + 3 * k, r + 3ll * i, 3 * sizeof(double));
				memcpy(os + 9 * k, orient + 9ll * i, 9 * sizeof(double));
				k++;
			}
		rc = upload_impl(h, N, (int) m, ids, box, rs, os, false);
	}
	free(mine);
	free(ids);
	free(rs);
	free(os);
	return rc;
}

int pp_slab_select(const pp_params *params, int rank, int world, int N, const double *box, const double *r, unsigned char *mine) {
	if(params == nullptr || box == nullptr || (N > 0 && (r == nullptr || mine == nullptr)) || world < 1 || rank < 0 || rank >= world) return 1;
	int n_side[3];
	if(even_grid(box, 1. + params->kf_delta, n_side, params->x_layers_multiple)) return 1;
	if(world == 1) {
		memset(mine, 1, (size_t) N);
		return 0;
	}
	const int nx = n_side[0];
	if(nx / 2 < world) return 1;
	int x0, x1;
	slab_partition(nx, rank, world, &x0, &x1);
	const int lo = (x0 - 1 + nx) % nx; // first held layer (the left halo); x1 - x0 + 2 layers are held
	const int held = x1 - x0 + 2;
	const double inv_box = 1. / box[0];
	for(int i = 0; i < N; i++) {
		int rel = cell_coord_host(r[3ll * i], 0., inv_box, nx) - lo;
		rel += rel < 0 ? nx : 0;
		mine[i] = rel < held;
	}
	return 0;
}

int pp_download(pp_handle *h, double *r, double *orient) {
	if(h == nullptr) return 1;
	if(h->slab_world > 1) return fail(h, "slab decomposition: use pp_slab_download");
	if(h->N == 0) return 0;
	launch_unpack_state(h->N, h->pos[h->cp], h->ori[h->co], h->pid[h->co], h->d_stage_r, h->d_stage_o, h->stream);
	CUDA_OK(h, cudaMemcpyAsync(r, h->d_stage_r, sizeof(double) * 3 * (size_t) h->N, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(orient, h->d_stage_o, sizeof(double) * 9 * (size_t) h->N, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

int pp_total_bonds(pp_handle *h, long long *bonds, int *overlap) {
	if(h == nullptr) return 1;
	CUDA_OK(h, cudaSetDevice(h->device));
	if(ensure_cells_valid(h)) return 1;
	long long b = 0;
	int ov = 0;
	if(count_bonds(h, h->cp, h->box, h->shift, &b, &ov)) return 1;
	h->bonds_base = b - h->h_counters->dbonds;
	h->bonds_known = true;
	if(bonds) *bonds = b;
	if(overlap) *overlap = ov;
	return 0;
}

int pp_bond_lists(pp_handle *h, int max_per_particle, int *counts, int *neighbours) {
	if(h == nullptr || counts == nullptr) return 1;
	if(neighbours != nullptr && max_per_particle < 1) return fail(h, "pp_bond_lists: max_per_particle must be positive");
	if(h->slab_world > 1) return fail(h, "pp_bond_lists / pp_bond_counts are single-engine calls");
	if(h->N == 0) return 0;
	CUDA_OK(h, cudaSetDevice(h->device));
	if(ensure_cells_valid(h)) return 1;
	const SimConst sc = make_const(h, h->box, h->shift);
	int *d_counts = h->cb.cell_of; // N ints of scratch, rewritten by the next cell build
	int *d_nb = nullptr;
	const size_t nb_bytes = sizeof(int) * (size_t) h->N * (size_t) (neighbours ? max_per_particle : 0);
	if(neighbours) {
		CUDA_OK(h, cudaMalloc(&d_nb, nb_bytes));
		CUDA_OK(h, cudaMemsetAsync(d_nb, 0xff, nb_bytes, h->stream)); // unused entries read -1
	}
	prof_begin(h);
	const int nl = launch_bond_counts(sc, h->cb.cell_start, h->pos[h->cp], h->ori[h->co], h->pid[h->co], d_counts, d_nb, max_per_particle, h->stream);
	prof_end(h, PP_K_BONDS, nl);
	cudaError_t e = cudaMemcpyAsync(counts, d_counts, sizeof(int) * (size_t) h->N, cudaMemcpyDeviceToHost, h->stream);
	if(e == cudaSuccess && neighbours) e = cudaMemcpyAsync(neighbours, d_nb, nb_bytes, cudaMemcpyDeviceToHost, h->stream);
	if(e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
	cudaFree(d_nb);
	if(e != cudaSuccess) return fail(h, "pp_bond_lists: %s", cudaGetErrorString(e));
	return 0;
}

int pp_bond_counts(pp_handle *h, int *counts) { return pp_bond_lists(h, 0, counts, nullptr); }

int pp_sweep(pp_handle *h, long long nsteps) {
	if(h == nullptr) return 1;
	if(h->slab_world > 1) return fail(h, "slab decomposition: drive the sweep with pp_slab_migrate / pp_slab_sweeps (or a pp_group)");
	if(h->N == 0) {
		h->step += nsteps;
		return 0;
	}
	CUDA_OK(h, cudaSetDevice(h->device));
	for(long long s = 0; s < nsteps; s++) {
		// box moves of this step: Binomial(N, 1/N) of them (MC.c:24-27, 62-63).  (A reference step is N iterations, each
		// EITHER a box move OR a particle move in NPT; here a step is N particle attempts plus the box moves.)
		if(h->prm.ensemble == PP_NPT || h->prm.Lx_move) {
			const int kind = h->prm.ensemble == PP_NPT ? PP_VOLUME : PP_LX;
			const int k = binomial_one_over_N(h->N, box_uniform(h->prm.seed, h->box_draws++));
			for(int m = 0; m < k; m++) {
				double u[3] = {0., 0., 0.};
				for(int j = 0; j < (kind == PP_VOLUME ? 3 : 2); j++) u[j] = box_uniform(h->prm.seed, h->box_draws++);
				int acc;
				if(box_move_single(h, kind, u, &acc)) return 1;
			}
		}
		// random grid shift in [0, cell width), redrawn every shift_every sweeps; between redraws the cell
		// lists stay valid because accepted moves never leave their cell
		const long long period = h->prm.shift_every;
		if(!h->sweep_grid || !h->cells_valid || h->step % period == 0) {
			if(even_grid(h->box, h->r_cut, h->n_side, h->prm.x_layers_multiple)) return fail(h, "box too small for the checkerboard");
			double shift[3];
			sweep_shift(h->prm.seed, (uint64_t) (h->step / period), h->box, h->n_side, shift);
			if(rebin(h, shift)) return 1;
			h->sweep_grid = true;
		}
		if(sweep_once(h)) return 1;
		h->step++;
	}
	CUDA_OK(h, cudaGetLastError());
	return 0;
}

int pp_sync(pp_handle *h) {
	if(h == nullptr) return 1;
	CUDA_OK(h, cudaSetDevice(h->device));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(h->peer_connected) {
		int status = 0;
		CUDA_OK(h, cudaMemcpy(&status, h->d_peer_status, sizeof(int), cudaMemcpyDeviceToHost));
		if(status) return fail(h, "a neighbouring engine did not reach the phase this engine waited for within 20 s");
	}
	return 0;
}

int pp_get_observables(pp_handle *h, pp_observables *obs, int reset) {
	if(h == nullptr || obs == nullptr) return 1;
	CUDA_OK(h, cudaSetDevice(h->device));
	if(fetch_counters(h)) return 1;
	const Counters &c = *h->h_counters, &k = h->carry;
	memset(obs, 0, sizeof(*obs));
	obs->step = h->step;
	obs->N = (h->slab_world > 1) ? h->N_global : h->N;
	memcpy(obs->box, h->box, sizeof(obs->box));
	obs->energy = -(double) running_bonds(h);
	for(int m = 0; m < PP_N_MOVES; m++) {
		obs->tries[m] = h->tries[m];
		obs->accepted[m] = h->accepted[m];
	}
	obs->tries[PP_ROTO_TRASL] += c.tries + k.tries;
	obs->accepted[PP_ROTO_TRASL] += c.accepted + k.accepted;
	obs->confined = c.confined + k.confined;
	obs->overlapped = c.overlapped + k.overlapped;
	obs->w_tries = c.w_tries + k.w_tries;
	obs->w_accepted = c.w_accepted + k.w_accepted;
	obs->acceptance_roto = (obs->w_tries > 0) ? (double) obs->w_accepted / (double) obs->w_tries : NAN;
	obs->pair_checks = c.pair_checks + k.pair_checks;
	memcpy(obs->n_side, h->n_side, sizeof(obs->n_side));
	obs->volume = h->V;
	if(reset) {
		// keep the energy: fold dbonds into the base, zero everything else (output.c:189)
		h->bonds_base += c.dbonds;
		CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(Counters), h->stream));
		memset(h->h_counters, 0, sizeof(Counters));
		memset(&h->carry, 0, sizeof(h->carry));
		memset(h->tries, 0, sizeof(h->tries));
		memset(h->accepted, 0, sizeof(h->accepted));
	}
	return 0;
}

int pp_check_energy(pp_handle *h, long long *bonds_scratch, int *overlap, int *ok) {
	if(h == nullptr) return 1;
	CUDA_OK(h, cudaSetDevice(h->device));
	if(ensure_cells_valid(h)) return 1;
	if(fetch_counters(h)) return 1;
	const long long running = running_bonds(h);
	long long b = 0;
	int ov = 0;
	if(count_bonds(h, h->cp, h->box, h->shift, &b, &ov)) return 1;
	// MC.c:774: re-orthonormalise every orientation (halo copies included: same input, same arithmetic as their owners)
	prof_begin(h);
	const int nl = launch_orthonormalise(h->N, h->ori[h->co], h->stream);
	prof_end(h, PP_K_ORTHO, nl);
	// cells_check (cells.c:65-76): the CSR offsets must account for every particle
	int total = 0;
	CUDA_OK(h, cudaMemcpyAsync(&total, h->cb.cell_start + h->n_cells, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	int good = 1;
	if(ov) good = 0;
	const double E_run = -(double) running, E_new = -(double) b;
	// (with a slab decomposition only the sum over engines is comparable: the driver checks it)
	if(h->slab_world <= 1 && h->bonds_known && fabs(E_run) > 1e-5 && fabs((E_new - E_run) / E_run) > 1e-5) good = 0; // MC.c:779
	if(total != h->N) good = 0;
	// E <- E_scratch (MC.c:783)
	h->bonds_base = b - h->h_counters->dbonds;
	h->bonds_known = true;
	if(bonds_scratch) *bonds_scratch = b;
	if(overlap) *overlap = ov;
	if(ok) *ok = good;
	if(!good) fail(h, "energy check failed: running bonds %lld, recomputed %lld, overlap %d, particles in cells %d / %d", running, b, ov, total, h->N);
	return 0;
}

int pp_eval_proposals(pp_handle *h, int n, const int *idx, const double *disp, const double *orient_new, const double *u, int commit, int *dE,
		int *overlap, int *accept) {
	if(h == nullptr) return 1;
	if(h->slab_world > 1) return fail(h, "pp_eval_proposals is a single-engine test hook");
	if(n <= 0) return 0;
	for(int k = 0; k < n; k++)
		if(idx[k] < 0 || idx[k] >= h->N) return fail(h, "proposal %d: particle index %d out of range", k, idx[k]);
	if(ensure_cells_valid(h)) return 1;
	int *d_idx, *d_out;
	double *d_disp, *d_orient, *d_u;
	CUDA_OK(h, cudaMalloc(&d_idx, sizeof(int) * (size_t) n));
	CUDA_OK(h, cudaMalloc(&d_out, sizeof(int) * 4 * (size_t) n));
	CUDA_OK(h, cudaMalloc(&d_disp, sizeof(double) * 3 * (size_t) n));
	CUDA_OK(h, cudaMalloc(&d_orient, sizeof(double) * 9 * (size_t) n));
	CUDA_OK(h, cudaMalloc(&d_u, sizeof(double) * (size_t) n));
	CUDA_OK(h, cudaMemcpyAsync(d_idx, idx, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(d_disp, disp, sizeof(double) * 3 * (size_t) n, cudaMemcpyHostToDevice, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(d_orient, orient_new, sizeof(double) * 9 * (size_t) n, cudaMemcpyHostToDevice, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(d_u, u, sizeof(double) * (size_t) n, cudaMemcpyHostToDevice, h->stream));
	int rc = 0;
	if(!commit) {
		const SimConst sc = make_const(h, h->box, h->shift);
		ProposalIO io = {d_idx, d_disp, d_orient, d_u, d_out, d_out + n, d_out + 2 * n, d_out + 3 * n};
		prof_begin(h);
		const int nl = launch_proposals(sc, n, io, h->cb.cell_start, h->pos[h->cp], h->ori[h->co], h->slot_of, h->stream);
		prof_end(h, PP_K_PROPOSALS, nl);
		cudaMemcpyAsync(dE, d_out, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
		cudaMemcpyAsync(overlap, d_out + n, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
		cudaMemcpyAsync(accept, d_out + 2 * n, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
		if(cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, "proposal kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
	}
	else {
		// sequential move list: one proposal at a time; accepted moves are written back, and the
		// cell structure is rebuilt when a particle changed cell (MC_change_cell, MC.c:180-206)
		for(int k = 0; k < n && rc == 0; k++) {
			if(ensure_cells_valid(h)) {
				rc = 1;
				break;
			}
			const SimConst sc = make_const(h, h->box, h->shift);
			ProposalIO io = {d_idx + k, d_disp + 3 * k, d_orient + 9 * k, d_u + k, d_out + k, d_out + n + k, d_out + 2 * n + k, d_out + 3 * n + k};
			prof_begin(h);
			const int nl = launch_proposals(sc, 1, io, h->cb.cell_start, h->pos[h->cp], h->ori[h->co], h->slot_of, h->stream);
			prof_end(h, PP_K_PROPOSALS, nl);
			int res[3];
			cudaMemcpyAsync(&res[0], d_out + 2 * n + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
			cudaMemcpyAsync(&res[1], d_out + 3 * n + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
			cudaMemcpyAsync(&res[2], d_out + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
			int slot;
			cudaMemcpyAsync(&slot, h->slot_of + idx[k], sizeof(int), cudaMemcpyDeviceToHost, h->stream);
			if(cudaStreamSynchronize(h->stream) != cudaSuccess) {
				rc = fail(h, "proposal kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
				break;
			}
			if(res[0]) {
				launch_commit(sc, slot, d_disp + 3 * k, d_orient + 9 * k, h->pos[h->cp], h->ori[h->co], h->cb.posf, h->stream);
				if(res[1]) h->cells_valid = false;
				h->bonds_base -= res[2]; // energy += dE, energy = -bonds
			}
		}
		if(rc == 0) {
			cudaMemcpyAsync(dE, d_out, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
			cudaMemcpyAsync(overlap, d_out + n, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
			cudaMemcpyAsync(accept, d_out + 2 * n, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream);
			if(cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, "proposal copy-back failed");
		}
	}
	cudaFree(d_idx);
	cudaFree(d_out);
	cudaFree(d_disp);
	cudaFree(d_orient);
	cudaFree(d_u);
	return rc;
}

int pp_volume_move(pp_handle *h, double u_dir, double u_vol, double u_acc, int *accepted) {
	if(h == nullptr) return 1;
	if(h->slab_world > 1) return fail(h, "slab decomposition: pp_slab_box_trial / pp_slab_box_finish (or a pp_group)");
	const double u[3] = {u_dir, u_vol, u_acc};
	int acc = 0;
	if(box_move_single(h, PP_VOLUME, u, &acc)) return 1;
	if(accepted) *accepted = acc;
	return 0;
}

int pp_Lx_move(pp_handle *h, double u_lx, double u_acc, int *accepted) {
	if(h == nullptr) return 1;
	if(h->slab_world > 1) return fail(h, "slab decomposition: pp_slab_box_trial / pp_slab_box_finish (or a pp_group)");
	const double u[3] = {u_lx, u_acc, 0.};
	int acc = 0;
	if(box_move_single(h, PP_LX, u, &acc)) return 1;
	if(accepted) *accepted = acc;
	return 0;
}

} // extern "C"
