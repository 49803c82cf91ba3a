// pp_internal.h -- the engine handle and the host-side helpers shared by pp_api.cu (single engine, slab hooks) and
// pp_group.cu (several engines driven from one host thread).  Not part of the ABI.
#pragma once

#include "../../include/pp_b200.h"
#include "pp_kernels.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

// one box move in flight (MC_change_volume / MC_change_Lx split into propose - trial - decide - finish so that the
// engines of a slab decomposition can take the decision together)
struct BoxMove {
	int kind;            // PP_VOLUME or PP_LX
	double nb[3], f[3], old_box[3];
	double initial_V, final_V, rescale, u_acc;
	bool slow;           // the trial box needed a fresh cell grid: the scaled, re-sorted state is current
	bool pending;
	long long bonds_old_total, bonds_local_new; // slab decomposition: kept between pp_slab_box_trial and pp_slab_box_finish
};

struct pp_handle {
	pp_params prm;
	char err[512];
	int device;
	cudaStream_t stream;
	bool own_stream;

	int N;
	size_t cap_N, cap_ids;
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
	pp::CellBuffers cb;
	pp::Counters *d_counters;
	pp::Counters *h_counters; // pinned
	pp::Counters carry;       // counters of earlier uploads of the same run (a box move that re-uploads the configuration keeps the books)
	// staging for upload / download / proposals
	double *d_stage_r, *d_stage_o;
	pp::Proposal *d_proposals; // one per trial of a sweep
	int reserve_N;             // pp_reserve: the arrays are sized for this many particles (insertions, pp_gc.cu)
	int *d_gc_out;             // bonds and overlap flag of a single-particle probe

	// grid
	int n_side[3];
	int n_cells;
	double shift[3];
	bool cells_valid;
	bool sweep_grid; // the current cells were built for the current sweep period's shift
	int occ_limit;   // the routing thresholds the current work lists were split with (evaluated at build time)
	int occ_mid;

	// bookkeeping
	long long step;
	long long bonds_base;  // bond count at the last full count
	bool bonds_known;
	long long tries[PP_N_MOVES], accepted[PP_N_MOVES];
	uint64_t box_draws;
	long long small_box_rejections; // box moves refused because the trial box could not hold two cells per side
	BoxMove bm;

	// x-slab decomposition (one engine per slab; world == 1: the whole periodic box)
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
	bool slab_pending;   // a classify has been done and waits for the rebuild
	long long rebuilt_step; // the sweep index the current slab cells were rebuilt for (-1: none yet)
	int layer_slots[6];  // slot offsets: left halo begin, first owned begin, first owned end, last owned begin, right halo begin, right halo end

	// Every device array of the state lives in one allocation (the arena), so that a neighbouring engine in another
	// process can map all of it with a single CUDA IPC handle: positions x 3, orientations x 2, shadows, the two
	// migration mailboxes (records stored by the left / right neighbour) and the control words.
	char *arena;
	size_t arena_bytes;
	pp_peer_info layout;          // offsets of the above inside the arena
	// neighbours: [0] left, [1] right
	char *peer_arena[2];          // the neighbour's arena as mapped here (IPC), or its plain device pointer (same process)
	pp_peer_info peer_layout[2];
	pp_handle *peer_handle[2];    // same-process neighbours (pp_group): ordering by CUDA events instead of device-side flags
	bool peer_same, peer_connected, peer_ipc;
	cudaEvent_t ev_phase[2], ev_mig;
	long long phase;              // phases (rebuilds, pass blocks, box moves) completed; the same sequence on every engine
	unsigned long long mig_epoch;
	unsigned int *d_push_done;
	int *d_peer_status;
	pp::HaloMirror hm;            // where accepted moves of the boundary layers are mirrored; rebuilt after every re-sort

	// profiling
	bool profiling;
	long long launches[PP_K_CLASSES];
	double ms[PP_K_CLASSES];
	cudaEvent_t ev0, ev1;
	long long total_launches;
};

namespace ppi {

constexpr int kSideLeft = 1, kSideRight = 2;

int fail(pp_handle *h, const char *fmt, ...);
void peer_release(pp_handle *h);
int cell_coord_host(double r, double shift, double inv_box, int n_side);
void slab_partition(int n_layers, int rank, int world, int *x0, int *x1);
int launch_passes(pp_handle *h, const pp::SimConst &sc, int first_pass, int n_pass);
int slab_after_sort(pp_handle *h);
int sort_into(pp_handle *h, const double *shift, const pp::StateView &src, const pp::StateView &dst);
pp::SimConst make_const(const pp_handle *h, const double *box, const double *shift);
pp::StateView view(pp_handle *h, int p, int o);
void prof_begin(pp_handle *h);
void prof_end(pp_handle *h, int kclass, int n_launches);
int even_grid(const double *box, double r_cut, int *n_side, int x_multiple);
int rebin(pp_handle *h, const double *shift);
int ensure_cells_valid(pp_handle *h);
int fetch_counters(pp_handle *h);
int count_bonds(pp_handle *h, int p, const double *box, const double *shift, long long *bonds, int *overlap);
long long running_bonds(pp_handle *h);
double box_uniform(uint64_t seed, uint64_t draw);
int binomial_one_over_N(int N, double u);
void sweep_shift(uint64_t seed, uint64_t epoch, const double *box, const int *n_side, double *shift);
int upload_impl(pp_handle *h, int N_global, int n, const int *ids, const double *box, const double *r, const double *orient, bool keep_books);
int sweep_once(pp_handle *h);  // proposals + the eight colour passes of one sweep on a single engine
// box moves (MC.c:450-585)
int box_propose(pp_handle *h, int kind, const double *u, int *status);
int box_trial_local(pp_handle *h, long long *bonds_local, int *overlap_local, int *status);
bool box_decide(pp_handle *h, long long bonds_new_total, int overlap, long long bonds_old_total, int N_total);
int box_finish(pp_handle *h, bool accepted, long long bonds_local_new);
// slab phases
int slab_wait(pp_handle *h, int sides);      // before a phase: the neighbours on these sides have completed as many phases as this engine
int slab_signal(pp_handle *h, bool layout);  // after a phase (layout: the slot layout changed, publish it with the phase number)
int slab_rebuild_begin(pp_handle *h);        // wait for both neighbours, classify on the new grid, store the outgoing records
int slab_rebuild_end(pp_handle *h);          // take in the neighbours' records, re-sort
int slab_block(pp_handle *h, int cx);        // one block of four colour passes between its wait and its signal
int slab_build_mirror(pp_handle *h);
bool slab_period_starts(const pp_handle *h);

#define CUDA_OK(h, call)                                                                                      \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                              \
		if(e_ != cudaSuccess) return ppi::fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
	} while(0)

} // namespace ppi
