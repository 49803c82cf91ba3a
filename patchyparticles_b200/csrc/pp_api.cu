// pp_api.cu -- the C ABI (include/pp_b200.h): handle, device memory, and the host-side scheduling
// that replaces do_NVT / do_NPT (MC.c:19-29, 60-68), MC_change_volume (MC.c:450-513),
// MC_change_Lx (MC.c:515-585) and MC_check_energy (MC.c:766-784).  No CPU fallback: every
// numerical result comes from the kernels in pp_cells.cu / pp_sweep.cu / pp_bonds.cu.
#include "../../include/pp_b200.h"
#include "pp_kernels.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace pp;

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

} // namespace

struct pp_handle {
	pp_params prm;
	char err[512];
	int device;
	cudaStream_t stream;
	bool own_stream;

	int N;
	size_t cap_N;
	size_t cap_cells;
	double box[3];
	double V;      // System.V: box product at upload (system.c:198), replaced by an accepted volume move (MC.c:500)
	double r_cut;
	double base[12];

	// device state: positions are triple buffered (current, box-move trial, sort destination),
	// orientations / ids double buffered
	double4 *pos[3];
	double *ori[2];
	int *pid[2];
	int cp, co;
	int *slot_of;
	CellBuffers cb;
	Counters *d_counters;
	Counters *h_counters; // pinned
	// staging for upload / download / proposals
	double *d_stage_r, *d_stage_o;
	Proposal *d_proposals; // one per trial of a sweep

	// grid
	int n_side[3];
	int n_cells;
	double shift[3];
	bool cells_valid;
	bool sweep_grid; // the current cells were built for the current sweep period's shift
	int occ_limit;   // the routing threshold the current work lists were split with (cells_occupancy_limit at build time)

	// bookkeeping
	long long step;
	long long bonds_base;  // bond count at the last full count
	bool bonds_known;
	long long tries[PP_N_MOVES], accepted[PP_N_MOVES];
	uint64_t box_draws;
	long long small_box_rejections; // box moves refused because the trial box could not hold two cells per side

	// x-slab decomposition (one engine per GPU; world == 1: the whole periodic box)
	int slab_rank, slab_world;
	int x0, x1;          // global x layers [x0, x1) owned by this engine
	int nxl;             // layers held: owned + 2 halo layers (n_side[0] when world == 1)
	int N_global;        // particles of the whole system
	double *outbox[2], *inbox[2]; // migration / halo records (13 doubles each) to / from the left (0) and right (1) neighbour
	int cap_records;
	int *d_slab_counts;  // keep, left, right, error, received from the left, received from the right, (2 spare)
	int *h_slab_ints;    // pinned: 8 ints read back per rebuild
	int keep_count, out_count[2], in_count[2];
	double pending_shift[3];
	bool slab_pending;   // a classify has been done and waits for pp_slab_sweep_rebuild
	int layer_slots[6];  // slot offsets: left halo begin, first owned begin, first owned end, last owned begin, right halo begin, right halo end

	// peer-memory halo refresh: my mailbox block ([from-left mailbox][from-right mailbox][2 flags]) and the
	// neighbours' blocks mapped with CUDA IPC
	char *mbox;
	size_t mbox_stride;           // bytes per mailbox
	long long mbox_cap;           // particles per mailbox
	char *peer_mbox[2];           // left / right neighbour's block
	bool peer_same, peer_connected;
	unsigned long long halo_epoch[2], mig_epoch;
	unsigned int *d_push_done;
	int *d_peer_status;

	// the two sweep kernels of a colour pass run concurrently: crowded cells (tile kernel) on `stream`, sparsely
	// populated ones (thread-per-cell kernel) on `stream2`
	cudaStream_t stream2;
	cudaEvent_t ev_a, ev_b;

	// profiling
	bool profiling;
	long long launches[PP_K_CLASSES];
	double ms[PP_K_CLASSES];
	cudaEvent_t ev0, ev1;
	long long total_launches;
};

namespace {

int fail(pp_handle *h, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	if(h) vsnprintf(h->err, sizeof(h->err), fmt, ap);
	va_end(ap);
	return 1;
}

#define CUDA_OK(h, call)                                                                                      \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                              \
		if(e_ != cudaSuccess) return fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while(0)

void peer_release(pp_handle *h) {
	if(h->peer_connected) {
		cudaIpcCloseMemHandle(h->peer_mbox[0]);
		if(!h->peer_same) cudaIpcCloseMemHandle(h->peer_mbox[1]);
	}
	h->peer_connected = false;
	h->peer_mbox[0] = h->peer_mbox[1] = nullptr;
	cudaFree(h->mbox);
	cudaFree(h->d_push_done);
	cudaFree(h->d_peer_status);
	h->mbox = nullptr;
	h->d_push_done = nullptr;
	h->d_peer_status = nullptr;
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

// even grid with cells of side >= r_cut (cells.c:22-28 uses floor(box / r_cut), any parity)
// the 1e-9 margin keeps non-adjacent cells more than r_cut apart whatever the rounding of the cell index
int even_grid(const double *box, double r_cut, int *n_side) {
	for(int k = 0; k < 3; k++) {
		int n = (int) floor(box[k] / r_cut);
		n &= ~1;
		if(n >= 2 && box[k] / n < r_cut * (1. + 1e-9)) n -= 2;
		if(n < 2) return 1;
		n_side[k] = n;
	}
	return 0;
}

void free_state(pp_handle *h) {
	for(int k = 0; k < 3; k++) cudaFree(h->pos[k]);
	for(int k = 0; k < 2; k++) {
		cudaFree(h->ori[k]);
		cudaFree(h->pid[k]);
	}
	cudaFree(h->slot_of);
	cudaFree(h->cb.cell_of);
	cudaFree(h->cb.cell_sorted);
	cudaFree(h->cb.posf);
	cudaFree(h->cb.perm);
	cudaFree(h->d_proposals);
	cudaFree(h->d_stage_r);
	cudaFree(h->d_stage_o);
	memset(h->pos, 0, sizeof(h->pos));
	memset(h->ori, 0, sizeof(h->ori));
	memset(h->pid, 0, sizeof(h->pid));
	h->slot_of = h->cb.cell_of = h->cb.cell_sorted = h->cb.perm = nullptr;
	h->d_proposals = nullptr;
	h->cb.posf = nullptr;
	h->d_stage_r = h->d_stage_o = nullptr;
	h->cap_N = 0;
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

int ensure_state(pp_handle *h, size_t n) {
	if(n <= h->cap_N) return 0;
	free_state(h);
	const size_t cap = n + 32;
	for(int k = 0; k < 3; k++) CUDA_OK(h, cudaMalloc(&h->pos[k], cap * sizeof(double4)));
	for(int k = 0; k < 2; k++) {
		CUDA_OK(h, cudaMalloc(&h->ori[k], cap * 9 * sizeof(double)));
		CUDA_OK(h, cudaMalloc(&h->pid[k], cap * sizeof(int)));
	}
	CUDA_OK(h, cudaMalloc(&h->slot_of, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.cell_of, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.perm, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.cell_sorted, cap * sizeof(int)));
	CUDA_OK(h, cudaMalloc(&h->cb.posf, cap * sizeof(uint2)));
	CUDA_OK(h, cudaMalloc(&h->d_proposals, cap * sizeof(Proposal)));
	CUDA_OK(h, cudaMalloc(&h->d_stage_r, cap * 3 * sizeof(double)));
	CUDA_OK(h, cudaMalloc(&h->d_stage_o, cap * 9 * sizeof(double)));
	h->cap_N = cap;
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

// Which sweep kernel gets which cells.  lanes_per_cell = 1: every cell to the thread-per-cell kernel; 4..32: every
// cell to the tile kernel of that width; 0 (default): cells with at most cells_max_occupancy members to the
// thread-per-cell kernel, the crowded ones to the tile kernel, concurrently.  The thread-per-cell kernel needs
// three distinct cells per axis.
int cells_occupancy_limit(const pp_handle *h) {
	const bool grid_ok = (h->slab_world > 1 || h->n_side[0] >= 4) && h->n_side[1] >= 4 && h->n_side[2] >= 4;
	if(!grid_ok) return 0;
	if(h->prm.lanes_per_cell == 1) return kWorkBuckets - 1;
	if(h->prm.cells_max_occupancy >= 0) return h->prm.cells_max_occupancy < kWorkBuckets ? h->prm.cells_max_occupancy : kWorkBuckets - 1;
	if(h->prm.lanes_per_cell != 0) return 0; // an explicit tile width alone: tile kernel only
	// measured on B200 (profiles/README.md): the thread-per-cell kernel wins once a colour pass fills the GPU; the
	// smaller the pass, the more its serial chain per cell matters and the fewer particles a cell may hold there
	return (h->N >= 800000) ? 3 : ((h->N >= 400000) ? 2 : ((h->N >= 150000) ? 1 : 0));
}

// lanes per cell of the thread-per-cell kernel: one lane per cell has the highest throughput, more lanes shorten
// the serial chain of a cell when a colour pass is too small to fill the GPU
int cell_lanes(const pp_handle *h) {
	const int l = h->prm.cell_lanes;
	if(l == 1 || l == 2 || l == 4 || l == 8) return l;
	return 1;
}

int tile_lanes(const pp_handle *h) {
	const int l = h->prm.lanes_per_cell;
	if(l == 4 || l == 8 || l == 16 || l == 32) return l;
	return (h->N >= 400000) ? 8 : 16; // measured on B200: profiles/README.md
}

// the colour passes `first_pass .. first_pass + n_pass - 1` of one sweep (pass order: x parity slowest)
int launch_passes(pp_handle *h, const SimConst &sc, int first_pass, int n_pass) {
	const int limit = h->occ_limit, lanes = tile_lanes(h); // the split the work lists were built with, not a fresh evaluation
	long long n_active = (long long) (sc.nxl / 2 + 1) * (sc.n_side[1] / 2) * (sc.n_side[2] / 2);
	if(n_active > sc.N) n_active = sc.N;
	int nl = 0;
	if(limit > 0) { // the second stream joins after everything already queued on the first
		cudaEventRecord(h->ev_a, h->stream);
		cudaStreamWaitEvent(h->stream2, h->ev_a, 0);
	}
	for(int pass = first_pass; pass < first_pass + n_pass; pass++) {
		const int colour = ((pass >> 2) & 1) | (((pass >> 1) & 1) << 1) | ((pass & 1) << 2);
		const int *ranges = h->cb.bins + kWorkRanges + 4 * colour;
		// crowded cells: the tile kernel, or (hybrid routing, no explicit tile width) the 8-lane latency configuration
		// of the thread-per-cell kernel
		if(limit > 0 && h->prm.lanes_per_cell == 0)
			nl += launch_sweep_cells(sc, ranges, n_active, 8, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->stream);
		else nl += launch_sweep_pass(sc, ranges, lanes, h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->stream);
		if(limit > 0) {
			nl += launch_sweep_cells(sc, ranges + 2, n_active, cell_lanes(h), h->cb, h->d_proposals, h->pos[h->cp], h->ori[h->co], h->d_counters, h->stream2);
			// the next pass (either kernel) starts when both kernels of this one are done
			cudaEventRecord(h->ev_a, h->stream);
			cudaEventRecord(h->ev_b, h->stream2);
			cudaStreamWaitEvent(h->stream, h->ev_b, 0);
			cudaStreamWaitEvent(h->stream2, h->ev_a, 0);
		}
	}
	return nl;
}

// slab decomposition: after a sort, particles outside the held layers sit in the trash cell at the end of the
// arrays and are dropped; the slot ranges of the boundary and halo layers are cached for the halo refresh
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

// K1 on the current state with the given grid shift; the sorted copy becomes current.
int rebin(pp_handle *h, const double *shift) {
	if(h->slab_world <= 1) {
		if(even_grid(h->box, h->r_cut, h->n_side)) return fail(h, "box (%g, %g, %g) is smaller than two cells of side r_cut = %g", h->box[0], h->box[1], h->box[2], h->r_cut);
		h->nxl = h->n_side[0];
	}
	h->n_cells = h->nxl * h->n_side[1] * h->n_side[2];
	if(ensure_cells(h, (size_t) h->n_cells + 2)) return 1;
	memcpy(h->shift, shift, sizeof(h->shift));
	const SimConst sc = make_const(h, h->box, h->shift);
	const int np = (h->cp + 1) % 3, no = h->co ^ 1;
	prof_begin(h);
	h->occ_limit = cells_occupancy_limit(h);
	const int nl = launch_cell_build(sc, h->cb, view(h, h->cp, h->co), view(h, np, no), h->slot_of, h->occ_limit, h->stream);
	prof_end(h, PP_K_CELLS, nl);
	h->cp = np;
	h->co = no;
	h->cells_valid = true;
	h->sweep_grid = false;
	CUDA_OK(h, cudaGetLastError());
	if(h->slab_world > 1 && slab_after_sort(h)) return 1;
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
	const int nl = launch_total_bonds(sc, h->cb.cell_start, h->pos[p], h->ori[h->co], h->d_counters, h->stream);
	prof_end(h, PP_K_BONDS, nl);
	CUDA_OK(h, cudaGetLastError());
	if(fetch_counters(h)) return 1;
	*bonds = h->h_counters->bonds;
	*overlap = h->h_counters->overlap;
	return 0;
}

long long running_bonds(pp_handle *h) {
	return h->bonds_base + h->h_counters->dbonds;
}

double next_box_uniform(pp_handle *h) {
	uint32_t x[4];
	philox_host(kTagBoxMove, 0u, (uint32_t) h->box_draws, (uint32_t) (h->box_draws >> 32), (uint32_t) h->prm.seed, (uint32_t) (h->prm.seed >> 32), x);
	h->box_draws++;
	return unit53(x[0], x[1]);
}

// number of successes of N Bernoulli(1/N) draws (the per-iteration test of MC.c:24-27, 62-63), by inversion
int binomial_one_over_N(pp_handle *h) {
	const int N = h->N;
	if(N <= 0) return 0;
	if(N == 1) return 1;
	const double p = 1. / N, q = 1. - p;
	const double u = next_box_uniform(h);
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

// A trial box with a side below two cells of side r_cut cannot carry the checkerboard grid.  The reference clamps its
// grid to three cells per side there (cells.c:24-27) and evaluates the move with cells smaller than the interaction
// range; the engine counts the move as tried and rejected and leaves positions, box and cells untouched, so a
// compressing run goes on instead of aborting.
bool box_too_small(pp_handle *h, const double *nb) {
	int ns[3];
	if(even_grid(nb, h->r_cut, ns) == 0) return false;
	h->small_box_rejections++;
	return true;
}

// One box move: rescale into the trial buffer, recount, Metropolis on the host.
// factors f[3], new box nb[3]; arg_fn computes the acceptance exponent from (bonds_new).
int box_trial(pp_handle *h, const double *nb, const double *f, long long *bonds_new, int *overlap) {
	// the running bond count must be current
	if(!h->bonds_known) return fail(h, "box move before the initial pp_total_bonds");
	if(ensure_cells_valid(h)) return 1;
	const int tp = (h->cp + 2) % 3;
	prof_begin(h);
	const int nl = launch_rescale(h->N, h->pos[h->cp], h->pos[tp], f[0], f[1], f[2], 0, h->stream);
	prof_end(h, PP_K_RESCALE, nl);
	int ns[3];
	bool regrid = even_grid(nb, h->r_cut, ns) != 0;
	for(int k = 0; k < 3 && !regrid; k++) regrid = (nb[k] / h->n_side[k] < h->r_cut * (1. + 1e-9));
	if(regrid) return 2; // caller takes the slow path
	double sh[3] = {h->shift[0] * f[0], h->shift[1] * f[1], h->shift[2] * f[2]};
	return count_bonds(h, tp, nb, sh, bonds_new, overlap);
}

// Slow path of a box move whose new box invalidates the grid: make the scaled state current, re-sort
// it on a fresh grid and count.  Returns with the scaled, sorted state current; *undo receives what
// is needed to go back (the reference divides back and rebuilds its cells, MC.c:505-508, 567-581).
int box_trial_regrid(pp_handle *h, const double *nb, long long *bonds_new, int *overlap) {
	const int tp = (h->cp + 2) % 3;
	h->cp = tp; // scaled copy becomes the source of the sort
	memcpy(h->box, nb, sizeof(h->box));
	const double zero[3] = {0., 0., 0.};
	if(rebin(h, zero)) return 1;
	return count_bonds(h, h->cp, h->box, h->shift, bonds_new, overlap);
}

int volume_move(pp_handle *h, double u_dir, double u_vol, double u_acc, int *accepted) {
	// MC.c:450-513
	if(fetch_counters(h)) return 1;
	h->tries[PP_VOLUME]++;
	const double initial_V = h->box[0] * h->box[1] * h->box[2];
	const int dir = (int) (u_dir * 3);
	const double ln_final_V = log(initial_V) + (u_vol - 0.5) * h->prm.rescale_factor_max;
	const double final_V = exp(ln_final_V);
	const double old_side = h->box[dir];
	double new_side = final_V;
	for(int k = 0; k < 3; k++)
		if(k != dir) new_side /= h->box[k];
	const double rescale = new_side / old_side;
	double nb[3] = {h->box[0], h->box[1], h->box[2]}, f[3] = {1., 1., 1.};
	nb[dir] = new_side;
	f[dir] = rescale;
	double old_box[3] = {h->box[0], h->box[1], h->box[2]};

	if(box_too_small(h, nb)) { // counted, rejected, nothing touched
		*accepted = 0;
		return 0;
	}
	long long bonds_new = 0;
	int overlap = 0;
	int rc = box_trial(h, nb, f, &bonds_new, &overlap);
	bool slow = false;
	if(rc == 2) {
		slow = true;
		rc = box_trial_regrid(h, nb, &bonds_new, &overlap);
	}
	if(rc) return 1;
	const long long bonds_old = running_bonds(h);
	const double delta_E = -(double) bonds_new + (double) bonds_old;
	const double delta_V = final_V - initial_V;
	const double arg = -(h->prm.P * delta_V + delta_E) / h->prm.temperature + (h->N + 1) * log(rescale);
	const bool acc = !overlap && u_acc < exp(arg);
	if(acc) {
		if(!slow) {
			h->cp = (h->cp + 2) % 3;
			memcpy(h->box, nb, sizeof(h->box));
			for(int k = 0; k < 3; k++) h->shift[k] *= f[k];
		}
		h->V = final_V; // MC.c:500
		h->bonds_base = bonds_new - h->h_counters->dbonds;
		h->accepted[PP_VOLUME]++;
		h->sweep_grid = false;
	}
	else if(slow) {
		// go back: divide the coordinates back like the reference and rebuild the cells for the old box
		const int tp = (h->cp + 2) % 3;
		prof_begin(h);
		const int nl = launch_rescale(h->N, h->pos[h->cp], h->pos[tp], f[0], f[1], f[2], 1, h->stream);
		prof_end(h, PP_K_RESCALE, nl);
		h->cp = tp;
		memcpy(h->box, old_box, sizeof(h->box));
		const double zero[3] = {0., 0., 0.};
		if(rebin(h, zero)) return 1;
	}
	*accepted = acc ? 1 : 0;
	return 0;
}

int Lx_move(pp_handle *h, double u_lx, double u_acc, int *accepted) {
	// MC.c:515-585
	if(fetch_counters(h)) return 1;
	h->tries[PP_LX]++;
	double old_box[3] = {h->box[0], h->box[1], h->box[2]};
	double nb[3];
	nb[0] = h->box[0] + (u_lx - 0.5) * h->prm.Lx_change_max;
	nb[1] = nb[2] = sqrt(h->V / nb[0]); // MC.c:524: the stored volume, not the product of the current sides
	if(nb[1] < h->prm.Lyz_min || nb[1] > h->prm.Lyz_max) {
		*accepted = -1;
		return 0;
	}
	double f[3] = {nb[0] / old_box[0], nb[1] / old_box[1], nb[2] / old_box[2]};
	if(box_too_small(h, nb)) { // counted, rejected, nothing touched
		*accepted = 0;
		return 0;
	}
	long long bonds_new = 0;
	int overlap = 0;
	int rc = box_trial(h, nb, f, &bonds_new, &overlap);
	bool slow = false;
	if(rc == 2) {
		slow = true;
		rc = box_trial_regrid(h, nb, &bonds_new, &overlap);
	}
	if(rc) return 1;
	const double delta_E = -(double) bonds_new + (double) running_bonds(h);
	const double arg = h->prm.Lx_reference_sign ? (delta_E / h->prm.temperature) : (-delta_E / h->prm.temperature);
	const bool acc = !overlap && u_acc < exp(arg);
	if(acc) {
		if(!slow) {
			h->cp = (h->cp + 2) % 3;
			memcpy(h->box, nb, sizeof(h->box));
			for(int k = 0; k < 3; k++) h->shift[k] *= f[k];
		}
		h->bonds_base = bonds_new - h->h_counters->dbonds;
		h->accepted[PP_LX]++;
		h->sweep_grid = false;
	}
	else if(slow) {
		const int tp = (h->cp + 2) % 3;
		prof_begin(h);
		const int nl = launch_rescale(h->N, h->pos[h->cp], h->pos[tp], f[0], f[1], f[2], 1, h->stream);
		prof_end(h, PP_K_RESCALE, nl);
		h->cp = tp;
		memcpy(h->box, old_box, sizeof(h->box));
		const double zero[3] = {0., 0., 0.};
		if(rebin(h, zero)) return 1;
	}
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
	*out = h; // returned even on failure so that the caller can read pp_last_error
	if(params->ensemble != PP_NVT && params->ensemble != PP_NPT) return fail(h, "Ensemble %d not supported on the GPU path (0 = NVT, 5 = NPT)", params->ensemble);
	if(params->ensemble == PP_NPT && params->Lx_move) return fail(h, "Lx_move is only available with Ensemble = 0 (as in the reference, system.c:165-171)");
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
	CUDA_OK(h, cudaEventCreate(&h->ev0));
	CUDA_OK(h, cudaEventCreate(&h->ev1));
	CUDA_OK(h, cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking));
	CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
	CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
	h->r_cut = 1. + params->kf_delta; // system.c:249
	base_patches_host(h->base);
	if(h->prm.lanes_per_cell != 1 && h->prm.lanes_per_cell != 4 && h->prm.lanes_per_cell != 8 && h->prm.lanes_per_cell != 16 && h->prm.lanes_per_cell != 32)
		h->prm.lanes_per_cell = 0; // automatic
	if(h->prm.shift_every < 1) h->prm.shift_every = 4;
	return 0;
}

int pp_destroy(pp_handle *h) {
	if(h == nullptr) return 0;
	if(h->stream) cudaStreamSynchronize(h->stream);
	free_state(h);
	free_cells(h);
	cudaFree(h->d_counters);
	cudaFree(h->cb.bins);
	for(int k = 0; k < 2; k++) {
		cudaFree(h->outbox[k]);
		cudaFree(h->inbox[k]);
	}
	cudaFree(h->d_slab_counts);
	cudaFreeHost(h->h_slab_ints);
	peer_release(h);
	cudaFreeHost(h->h_counters);
	if(h->ev0) cudaEventDestroy(h->ev0);
	if(h->ev1) cudaEventDestroy(h->ev1);
	if(h->ev_a) cudaEventDestroy(h->ev_a);
	if(h->ev_b) cudaEventDestroy(h->ev_b);
	if(h->stream2) cudaStreamDestroy(h->stream2);
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
	CUDA_OK(h, cudaSetDevice(h->device));
	for(int k = 0; k < 3; k++)
		if(!(box[k] > 0.)) return fail(h, "box side %d is not positive", k);
	h->N = N;
	h->N_global = N;
	memcpy(h->box, box, sizeof(h->box));
	h->V = box[0] * box[1] * box[2];
	if(even_grid(h->box, h->r_cut, h->n_side)) return fail(h, "box (%g, %g, %g) is smaller than two cells of side r_cut = %g", box[0], box[1], box[2], h->r_cut);
	h->nxl = h->n_side[0];
	if(h->slab_world > 1) {
		// whole pairs of x layers per engine: every engine's first owned layer is even, its last one odd
		const int pairs = h->n_side[0] / 2, base = pairs / h->slab_world, rem = pairs % h->slab_world, r_ = h->slab_rank;
		if(base < 1) return fail(h, "%d x layers cannot be split over %d engines (at least 2 layers each)", h->n_side[0], h->slab_world);
		h->x0 = 2 * (r_ * base + (r_ < rem ? r_ : rem));
		h->x1 = h->x0 + 2 * (base + (r_ < rem ? 1 : 0));
		h->nxl = (h->x1 - h->x0) + 2;
		if(h->prm.ensemble != PP_NVT || h->prm.Lx_move) return fail(h, "the slab decomposition supports NVT roto-translations only (box moves: single engine)");
		h->cap_records = 4 * (N / h->n_side[0] + 1) + 8192;
		for(int k = 0; k < 2; k++) {
			cudaFree(h->outbox[k]);
			cudaFree(h->inbox[k]);
			CUDA_OK(h, cudaMalloc(&h->outbox[k], sizeof(double) * 13 * (size_t) h->cap_records));
			CUDA_OK(h, cudaMalloc(&h->inbox[k], sizeof(double) * 13 * (size_t) h->cap_records));
		}
		if(h->d_slab_counts == nullptr) CUDA_OK(h, cudaMalloc(&h->d_slab_counts, 16 * sizeof(int)));
		if(h->h_slab_ints == nullptr) CUDA_OK(h, cudaMallocHost(&h->h_slab_ints, 8 * sizeof(int)));
		h->slab_pending = false;
	}
	if(ensure_state(h, (size_t) N)) return 1;
	if(N > 0) {
		CUDA_OK(h, cudaMemcpyAsync(h->d_stage_r, r, sizeof(double) * 3 * (size_t) N, cudaMemcpyHostToDevice, h->stream));
		CUDA_OK(h, cudaMemcpyAsync(h->d_stage_o, orient, sizeof(double) * 9 * (size_t) N, cudaMemcpyHostToDevice, h->stream));
	}
	h->cp = 0;
	h->co = 0;
	launch_pack_state(N, h->d_stage_r, h->d_stage_o, h->pos[0], h->ori[0], h->pid[0], h->stream);
	CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(Counters), h->stream));
	memset(h->h_counters, 0, sizeof(Counters));
	h->cells_valid = false;
	h->sweep_grid = false;
	h->step = 0;
	h->bonds_base = 0;
	h->bonds_known = false;
	h->box_draws = 0;
	memset(h->tries, 0, sizeof(h->tries));
	memset(h->accepted, 0, sizeof(h->accepted));
	if(ensure_cells_valid(h)) return 1;
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
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
	if(h->slab_world > 1) return fail(h, "slab decomposition: drive the sweep with pp_slab_sweep_begin / _rebuild / _pass_block / _sweep_end");
	if(h->N == 0) {
		h->step += nsteps;
		return 0;
	}
	CUDA_OK(h, cudaSetDevice(h->device));
	for(long long s = 0; s < nsteps; s++) {
		// box moves of this step: Binomial(N, 1/N) of them (MC.c:24-27, 62-63)
		if(h->prm.ensemble == PP_NPT) {
			const int k = binomial_one_over_N(h);
			for(int m = 0; m < k; m++) {
				const double u_dir = next_box_uniform(h), u_vol = next_box_uniform(h), u_acc = next_box_uniform(h);
				int acc;
				if(volume_move(h, u_dir, u_vol, u_acc, &acc)) return 1;
			}
		}
		else if(h->prm.Lx_move) {
			const int k = binomial_one_over_N(h);
			for(int m = 0; m < k; m++) {
				const double u_lx = next_box_uniform(h), u_acc = next_box_uniform(h);
				int acc;
				if(Lx_move(h, u_lx, u_acc, &acc)) return 1;
			}
		}
		// random grid shift in [0, cell width), redrawn every shift_every sweeps; between redraws the cell
		// lists stay valid because accepted moves never leave their cell
		const long long period = h->prm.shift_every;
		if(!h->sweep_grid || !h->cells_valid || h->step % period == 0) {
			const uint64_t epoch = (uint64_t) (h->step / period);
			uint32_t x[4];
			philox_host(kTagShift, 0u, (uint32_t) epoch, (uint32_t) ((epoch >> 32) << 8), (uint32_t) h->prm.seed, (uint32_t) (h->prm.seed >> 32), x);
			if(even_grid(h->box, h->r_cut, h->n_side)) return fail(h, "box too small for the checkerboard");
			double shift[3];
			for(int k = 0; k < 3; k++) shift[k] = unit_from_u32(x[k]) * (h->box[k] / h->n_side[k]);
			if(rebin(h, shift)) return 1;
			h->sweep_grid = true;
		}
		const SimConst sc = make_const(h, h->box, h->shift);
		prof_begin(h);
		const int nr = launch_proposals_for_sweep(sc, h->cb, h->d_proposals, h->stream);
		prof_end(h, PP_K_RNG, nr);
		prof_begin(h);
		const int nl = launch_passes(h, sc, 0, 8);
		prof_end(h, PP_K_SWEEP, nl);
		h->step++;
	}
	CUDA_OK(h, cudaGetLastError());
	return 0;
}

int pp_sync(pp_handle *h) {
	if(h == nullptr) return 1;
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(h->peer_connected) {
		int status = 0;
		CUDA_OK(h, cudaMemcpy(&status, h->d_peer_status, sizeof(int), cudaMemcpyDeviceToHost));
		if(status) return fail(h, "a halo refresh timed out waiting for the neighbouring engine");
	}
	return 0;
}

int pp_get_observables(pp_handle *h, pp_observables *obs, int reset) {
	if(h == nullptr || obs == nullptr) return 1;
	if(fetch_counters(h)) return 1;
	const Counters &c = *h->h_counters;
	memset(obs, 0, sizeof(*obs));
	obs->step = h->step;
	obs->N = (h->slab_world > 1) ? h->N_global : h->N;
	memcpy(obs->box, h->box, sizeof(obs->box));
	obs->energy = -(double) running_bonds(h);
	for(int k = 0; k < PP_N_MOVES; k++) {
		obs->tries[k] = h->tries[k];
		obs->accepted[k] = h->accepted[k];
	}
	obs->tries[PP_ROTO_TRASL] += c.tries;
	obs->accepted[PP_ROTO_TRASL] += c.accepted;
	obs->confined = c.confined;
	obs->overlapped = c.overlapped;
	obs->acceptance_roto = (c.w_tries > 0) ? (double) c.w_accepted / (double) c.w_tries : NAN;
	obs->pair_checks = c.pair_checks;
	memcpy(obs->n_side, h->n_side, sizeof(obs->n_side));
	obs->volume = h->V;
	if(reset) {
		// keep the energy: fold dbonds into the base, zero everything else (output.c:189)
		h->bonds_base += c.dbonds;
		CUDA_OK(h, cudaMemsetAsync(h->d_counters, 0, sizeof(Counters), h->stream));
		memset(h->h_counters, 0, sizeof(Counters));
		memset(h->tries, 0, sizeof(h->tries));
		memset(h->accepted, 0, sizeof(h->accepted));
	}
	return 0;
}

int pp_check_energy(pp_handle *h, long long *bonds_scratch, int *overlap, int *ok) {
	if(h == nullptr) return 1;
	if(ensure_cells_valid(h)) return 1;
	if(fetch_counters(h)) return 1;
	const long long running = running_bonds(h);
	long long b = 0;
	int ov = 0;
	if(count_bonds(h, h->cp, h->box, h->shift, &b, &ov)) return 1;
	// MC.c:774: re-orthonormalise every orientation
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
	if(h->slab_world <= 1 && total != h->N) good = 0;
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
				launch_commit(slot, d_disp + 3 * k, d_orient + 9 * k, h->pos[h->cp], h->ori[h->co], h->stream);
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
	int acc = 0;
	if(volume_move(h, u_dir, u_vol, u_acc, &acc)) return 1;
	if(accepted) *accepted = acc;
	return 0;
}

int pp_Lx_move(pp_handle *h, double u_lx, double u_acc, int *accepted) {
	if(h == nullptr) return 1;
	int acc = 0;
	if(Lx_move(h, u_lx, u_acc, &acc)) return 1;
	if(accepted) *accepted = acc;
	return 0;
}

// ---------------------------------------------------------------------------------------------------
// x-slab decomposition hooks (one engine per GPU).  The driver (patchyparticles_b200/slab.py under
// torchrun, or any host with a transport) moves the outbox / halo buffers between neighbours.
// ---------------------------------------------------------------------------------------------------
int pp_slab_setup(pp_handle *h, int rank, int world) {
	if(h == nullptr) return 1;
	if(world < 1 || rank < 0 || rank >= world) return fail(h, "bad slab rank %d / world %d", rank, world);
	h->slab_rank = rank;
	h->slab_world = world;
	return 0;
}

int pp_slab_range(pp_handle *h, int *x0, int *x1, int *nx, int *n_local) {
	if(h == nullptr) return 1;
	if(x0) *x0 = (h->slab_world > 1) ? h->x0 : 0;
	if(x1) *x1 = (h->slab_world > 1) ? h->x1 : h->n_side[0];
	if(nx) *nx = h->n_side[0];
	if(n_local) *n_local = h->N;
	return 0;
}

int pp_slab_sweep_begin(pp_handle *h, int *exchange) {
	if(h == nullptr || exchange == nullptr) return 1;
	if(h->slab_world <= 1) return fail(h, "pp_slab_sweep_begin without a slab decomposition");
	CUDA_OK(h, cudaSetDevice(h->device));
	*exchange = 0;
	const long long period = h->prm.shift_every;
	if(h->sweep_grid && h->cells_valid && h->step % period != 0) return 0;
	// new shift period: same stream as the single-engine path, so every engine draws the same shift
	const uint64_t epoch = (uint64_t) (h->step / period);
	uint32_t x[4];
	philox_host(kTagShift, 0u, (uint32_t) epoch, (uint32_t) ((epoch >> 32) << 8), (uint32_t) h->prm.seed, (uint32_t) (h->prm.seed >> 32), x);
	for(int k = 0; k < 3; k++) h->pending_shift[k] = unit_from_u32(x[k]) * (h->box[k] / h->n_side[k]);
	if(ensure_cells_valid(h)) return 1;
	const SimConst sc_new = make_const(h, h->box, h->pending_shift);
	const int tp = (h->cp + 2) % 3, to = h->co ^ 1;
	prof_begin(h);
	const int nl = launch_slab_classify(sc_new, h->layer_slots[1], h->layer_slots[4], h->x0, h->x1 - h->x0, view(h, h->cp, h->co), view(h, tp, to),
			h->outbox[0], h->outbox[1], h->cap_records, h->d_slab_counts, h->stream);
	prof_end(h, PP_K_CELLS, nl);
	int counts[4];
	CUDA_OK(h, cudaMemcpyAsync(counts, h->d_slab_counts, sizeof(counts), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(counts[3]) return fail(h, "%d particles moved by more than one x layer at a grid shift", counts[3]);
	if(counts[1] > h->cap_records || counts[2] > h->cap_records) return fail(h, "slab outbox overflow (%d, %d records, capacity %d)", counts[1], counts[2], h->cap_records);
	h->keep_count = counts[0];
	h->out_count[0] = counts[1];
	h->out_count[1] = counts[2];
	h->in_count[0] = h->in_count[1] = 0;
	h->slab_pending = true;
	*exchange = 1;
	return 0;
}

int pp_slab_outbox(pp_handle *h, int side, void **dev, int *count) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(dev) *dev = h->outbox[side];
	if(count) *count = h->out_count[side];
	return 0;
}

int pp_slab_inbox(pp_handle *h, int side, int count, void **dev) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(count < 0 || count > h->cap_records) return fail(h, "slab inbox overflow (%d records, capacity %d)", count, h->cap_records);
	h->in_count[side] = count;
	if(dev) *dev = h->inbox[side];
	return 0;
}

int pp_slab_sweep_rebuild(pp_handle *h) {
	if(h == nullptr) return 1;
	if(!h->slab_pending) return fail(h, "pp_slab_sweep_rebuild without a pending exchange");
	const int tp = (h->cp + 2) % 3, to = h->co ^ 1;
	const int n_new = h->keep_count + h->in_count[0] + h->in_count[1];
	if((size_t) n_new > h->cap_N) return fail(h, "slab: %d local particles exceed the capacity %zu", n_new, h->cap_N);
	launch_slab_unpack(h->in_count[0], h->inbox[0], h->keep_count, view(h, tp, to), h->stream);
	launch_slab_unpack(h->in_count[1], h->inbox[1], h->keep_count + h->in_count[0], view(h, tp, to), h->stream);
	// sort the merged set on the new grid; the previous current buffers are free to be the destination
	h->N = n_new;
	memcpy(h->shift, h->pending_shift, sizeof(h->shift));
	const SimConst sc = make_const(h, h->box, h->shift);
	prof_begin(h);
	h->occ_limit = cells_occupancy_limit(h);
	const int nl = launch_cell_build(sc, h->cb, view(h, tp, to), view(h, h->cp, h->co), h->slot_of, h->occ_limit, h->stream);
	prof_end(h, PP_K_CELLS, nl);
	CUDA_OK(h, cudaGetLastError());
	h->cells_valid = true;
	h->sweep_grid = true;
	h->slab_pending = false;
	if(slab_after_sort(h)) return 1;
	if(h->N != n_new) return fail(h, "slab: %d of %d particles fell outside the held layers after the exchange", n_new - h->N, n_new);
	return 0;
}

int pp_slab_pass_block(pp_handle *h, int cx) {
	if(h == nullptr || cx < 0 || cx > 1) return 1;
	if(h->slab_world <= 1) return fail(h, "pp_slab_pass_block without a slab decomposition");
	if(!h->sweep_grid || !h->cells_valid) return fail(h, "pp_slab_pass_block before pp_slab_sweep_begin / _rebuild");
	const SimConst sc = make_const(h, h->box, h->shift);
	if(cx == 0) {
		prof_begin(h);
		const int nr = launch_proposals_for_sweep(sc, h->cb, h->d_proposals, h->stream);
		prof_end(h, PP_K_RNG, nr);
	}
	prof_begin(h);
	const int nl = launch_passes(h, sc, 4 * cx, 4);
	prof_end(h, PP_K_SWEEP, nl);
	CUDA_OK(h, cudaGetLastError());
	return 0;
}

int pp_slab_halo(pp_handle *h, int side, int recv, void **pos, void **ori, void **posf, int *count) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(h->slab_world <= 1 || !h->cells_valid) return fail(h, "pp_slab_halo: no valid slab layout");
	// send: my first (side 0) / last (side 1) owned layer; recv: my left (side 0) / right (side 1) halo layer
	int b, e;
	if(recv) {
		b = side == 0 ? h->layer_slots[0] : h->layer_slots[4];
		e = side == 0 ? h->layer_slots[1] : h->layer_slots[5];
	}
	else {
		b = side == 0 ? h->layer_slots[1] : h->layer_slots[3];
		e = side == 0 ? h->layer_slots[2] : h->layer_slots[4];
	}
	if(pos) *pos = h->pos[h->cp] + b;
	if(ori) *ori = h->ori[h->co] + 9ll * b;
	if(posf) *posf = h->cb.posf + b;
	if(count) *count = e - b;
	return 0;
}

// ---- halo refresh over peer memory -------------------------------------------------------------------
int pp_slab_peer_export(pp_handle *h, void *ipc_handle_64) {
	if(h == nullptr || ipc_handle_64 == nullptr) return 1;
	if(h->slab_world <= 1 || !h->cells_valid) return fail(h, "pp_slab_peer_export: upload a configuration in slab mode first");
	CUDA_OK(h, cudaSetDevice(h->device));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	peer_release(h);
	h->mbox_cap = h->cap_records;
	h->mbox_stride = ((size_t) h->mbox_cap * 112 + 255) / 256 * 256;
	const size_t bytes = 4 * h->mbox_stride + 256; // halo from-left, halo from-right, migration from-left, migration from-right, flags
	CUDA_OK(h, cudaMalloc(&h->mbox, bytes));
	CUDA_OK(h, cudaMemset(h->mbox, 0, bytes));
	CUDA_OK(h, cudaMalloc(&h->d_push_done, sizeof(unsigned int)));
	CUDA_OK(h, cudaMemset(h->d_push_done, 0, sizeof(unsigned int)));
	CUDA_OK(h, cudaMalloc(&h->d_peer_status, sizeof(int)));
	CUDA_OK(h, cudaMemset(h->d_peer_status, 0, sizeof(int)));
	h->halo_epoch[0] = h->halo_epoch[1] = 0;
	h->mig_epoch = 0;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI passes IPC handles as 64 bytes");
	cudaIpcMemHandle_t mh;
	CUDA_OK(h, cudaIpcGetMemHandle(&mh, h->mbox));
	memcpy(ipc_handle_64, &mh, 64);
	return 0;
}

int pp_slab_peer_connect(pp_handle *h, const void *left_handle_64, const void *right_handle_64) {
	if(h == nullptr || left_handle_64 == nullptr || right_handle_64 == nullptr) return 1;
	if(h->mbox == nullptr) return fail(h, "pp_slab_peer_connect before pp_slab_peer_export");
	CUDA_OK(h, cudaSetDevice(h->device));
	cudaIpcMemHandle_t l, r;
	memcpy(&l, left_handle_64, 64);
	memcpy(&r, right_handle_64, 64);
	h->peer_same = memcmp(&l, &r, 64) == 0; // two engines: both neighbours are the same one
	CUDA_OK(h, cudaIpcOpenMemHandle((void **) &h->peer_mbox[0], l, cudaIpcMemLazyEnablePeerAccess));
	if(h->peer_same) h->peer_mbox[1] = h->peer_mbox[0];
	else CUDA_OK(h, cudaIpcOpenMemHandle((void **) &h->peer_mbox[1], r, cudaIpcMemLazyEnablePeerAccess));
	h->peer_connected = true;
	return 0;
}

int pp_slab_peer_close(pp_handle *h) {
	if(h == nullptr) return 1;
	if(h->stream) cudaStreamSynchronize(h->stream);
	peer_release(h);
	return 0;
}

// side 0: my first owned layer -> the left neighbour's right halo, then my right halo <- what the right neighbour
// pushed; side 1: my last owned layer -> the right neighbour's left halo, then my left halo <- the left neighbour.
int pp_slab_refresh(pp_handle *h, int side) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(!h->peer_connected || !h->cells_valid) return fail(h, "pp_slab_refresh: no peer connection / no valid slab layout");
	CUDA_OK(h, cudaSetDevice(h->device));
	const int sb = side == 0 ? h->layer_slots[1] : h->layer_slots[3], se = side == 0 ? h->layer_slots[2] : h->layer_slots[4];
	const int rb = side == 0 ? h->layer_slots[4] : h->layer_slots[0], re = side == 0 ? h->layer_slots[5] : h->layer_slots[1];
	if(se - sb > h->mbox_cap || re - rb > h->mbox_cap) return fail(h, "halo layer of %d particles exceeds the mailbox capacity %lld", se - sb > re - rb ? se - sb : re - rb, h->mbox_cap);
	const unsigned long long epoch = ++h->halo_epoch[side];
	// the neighbour's mailbox for data coming from me: its from-right box (1) when I am its right neighbour (side 0)
	const int box = side == 0 ? 1 : 0;
	char *peer = h->peer_mbox[side == 0 ? 0 : 1];
	unsigned long long *peer_flags = reinterpret_cast<unsigned long long *>(peer + 4 * h->mbox_stride);
	const unsigned long long *my_flags = reinterpret_cast<const unsigned long long *>(h->mbox + 4 * h->mbox_stride);
	double4 *pos = h->pos[h->cp];
	double *ori = h->ori[h->co];
	prof_begin(h);
	int nl = launch_halo_push(pos + sb, ori + 9ll * sb, h->cb.posf + sb, se - sb, peer + box * h->mbox_stride, h->mbox_cap, peer_flags + box, epoch,
			h->d_push_done, h->stream);
	nl += launch_halo_pull(h->mbox + box * h->mbox_stride, h->mbox_cap, my_flags + box, epoch, pos + rb, ori + 9ll * rb, h->cb.posf + rb, re - rb,
			h->d_peer_status, h->stream);
	prof_end(h, PP_K_HALO, nl);
	CUDA_OK(h, cudaGetLastError());
	return 0;
}

// Start of a shift period over peer memory: classify on the new grid, store the outgoing records straight into the
// neighbours' migration mailboxes, take in what they stored here, rebuild the cells.  One host synchronisation (the
// counts), no library transfer.  *migrated = 0 when the current period is still running.
int pp_slab_migrate(pp_handle *h, int *migrated) {
	if(h == nullptr || migrated == nullptr) return 1;
	if(!h->peer_connected) return fail(h, "pp_slab_migrate needs the peer-memory connection (pp_slab_peer_connect)");
	CUDA_OK(h, cudaSetDevice(h->device));
	*migrated = 0;
	const long long period = h->prm.shift_every;
	if(h->sweep_grid && h->cells_valid && h->step % period != 0) return 0;
	const uint64_t epoch = (uint64_t) (h->step / period);
	uint32_t x[4];
	philox_host(kTagShift, 0u, (uint32_t) epoch, (uint32_t) ((epoch >> 32) << 8), (uint32_t) h->prm.seed, (uint32_t) (h->prm.seed >> 32), x);
	for(int k = 0; k < 3; k++) h->pending_shift[k] = unit_from_u32(x[k]) * (h->box[k] / h->n_side[k]);
	if(ensure_cells_valid(h)) return 1;
	const SimConst sc_new = make_const(h, h->box, h->pending_shift);
	const int tp = (h->cp + 2) % 3, to = h->co ^ 1;
	prof_begin(h);
	int nl = launch_slab_classify(sc_new, h->layer_slots[1], h->layer_slots[4], h->x0, h->x1 - h->x0, view(h, h->cp, h->co), view(h, tp, to),
			h->outbox[0], h->outbox[1], h->cap_records, h->d_slab_counts, h->stream);
	prof_end(h, PP_K_CELLS, nl);
	prof_begin(h);
	nl = 0;
	const unsigned long long me = ++h->mig_epoch;
	char *flags_area[2] = {h->peer_mbox[0] + 4 * h->mbox_stride, h->peer_mbox[1] + 4 * h->mbox_stride};
	char *my_area = h->mbox + 4 * h->mbox_stride; // 4 flags (u64), then the two migration counts (int)
	// to the left neighbour: its from-right migration box (3); to the right neighbour: its from-left box (2)
	nl += launch_records_push(h->outbox[0], h->d_slab_counts, 1, h->cap_records, reinterpret_cast<double *>(h->peer_mbox[0] + 3 * h->mbox_stride),
			reinterpret_cast<int *>(flags_area[0] + 32) + 1, reinterpret_cast<unsigned long long *>(flags_area[0]) + 3, me, h->d_push_done, h->stream);
	nl += launch_records_push(h->outbox[1], h->d_slab_counts, 2, h->cap_records, reinterpret_cast<double *>(h->peer_mbox[1] + 2 * h->mbox_stride),
			reinterpret_cast<int *>(flags_area[1] + 32) + 0, reinterpret_cast<unsigned long long *>(flags_area[1]) + 2, me, h->d_push_done, h->stream);
	nl += launch_records_pull(reinterpret_cast<const double *>(h->mbox + 2 * h->mbox_stride), reinterpret_cast<const int *>(my_area + 32) + 0,
			reinterpret_cast<const unsigned long long *>(my_area) + 2, me, h->cap_records, h->inbox[0], h->d_slab_counts + 4, h->d_peer_status, h->stream);
	nl += launch_records_pull(reinterpret_cast<const double *>(h->mbox + 3 * h->mbox_stride), reinterpret_cast<const int *>(my_area + 32) + 1,
			reinterpret_cast<const unsigned long long *>(my_area) + 3, me, h->cap_records, h->inbox[1], h->d_slab_counts + 5, h->d_peer_status, h->stream);
	prof_end(h, PP_K_HALO, nl);
	int *c = h->h_slab_ints;
	CUDA_OK(h, cudaMemcpyAsync(c, h->d_slab_counts, 6 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(c[3]) return fail(h, "%d particles moved by more than one x layer at a grid shift", c[3]);
	if(c[1] > h->cap_records || c[2] > h->cap_records) return fail(h, "slab outbox overflow (%d, %d records, capacity %d)", c[1], c[2], h->cap_records);
	h->keep_count = c[0];
	h->out_count[0] = c[1];
	h->out_count[1] = c[2];
	h->in_count[0] = c[4];
	h->in_count[1] = c[5];
	h->slab_pending = true;
	if(pp_slab_sweep_rebuild(h)) return 1;
	*migrated = 1;
	return 0;
}

// Runs whole sweeps (pass block, refresh, pass block, refresh) over peer memory until `max_sweeps` are done or the
// next grid shift needs a migration exchange, which is the driver's job.  -> number of sweeps done in *done.
int pp_slab_sweeps(pp_handle *h, int max_sweeps, int *done) {
	if(h == nullptr || done == nullptr) return 1;
	*done = 0;
	if(!h->peer_connected) return fail(h, "pp_slab_sweeps needs the peer-memory halo refresh (pp_slab_peer_connect)");
	const long long period = h->prm.shift_every;
	while(*done < max_sweeps) {
		if(!(h->sweep_grid && h->cells_valid && h->step % period != 0)) break; // a new shift period starts here
		for(int cx = 0; cx < 2; cx++) {
			if(pp_slab_pass_block(h, cx)) return 1;
			if(pp_slab_refresh(h, cx)) return 1;
		}
		h->step++;
		(*done)++;
	}
	return 0;
}

int pp_slab_sweep_end(pp_handle *h) {
	if(h == nullptr) return 1;
	h->step++;
	return 0;
}

int pp_slab_download(pp_handle *h, int *n_owned, int *ids, double *r, double *orient) {
	if(h == nullptr) return 1;
	if(h->slab_world <= 1) return fail(h, "pp_slab_download without a slab decomposition");
	if(ensure_cells_valid(h)) return 1;
	const int b = h->layer_slots[1], e = h->layer_slots[4], n = e - b;
	if(n_owned) *n_owned = n;
	if(ids == nullptr || n == 0) return 0;
	int *d_ids = h->cb.cell_of; // scratch
	launch_slab_download(b, e, view(h, h->cp, h->co), d_ids, h->d_stage_r, h->d_stage_o, h->stream);
	CUDA_OK(h, cudaMemcpyAsync(ids, d_ids, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(r, h->d_stage_r, sizeof(double) * 3 * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(orient, h->d_stage_o, sizeof(double) * 9 * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

} // extern "C"
