// pp_kernels.cuh -- host-visible launchers of the engine's kernels (K1..K6 of SURVEY.md section 2).
#pragma once

#include "pp_common.cuh"

namespace pp {

// Particle state in cell order (structure of arrays).
struct StateView {
	double4 *pos; // x, y, z, (double) particle id
	double *ori;  // 9 doubles per particle, row-major orientation matrix
	int *pid;     // particle id per slot
};

// work lists of the sweep: 8 colours x kWorkBuckets occupancy buckets (descending occupancy) x kWorkSub classes of the
// neighbourhood size (descending, four neighbours wide: the lanes of a warp of the thread-per-cell kernel then run the
// same number of trials over the same number of staged entries)
constexpr int kWorkBuckets = 16;
constexpr int kWorkSub = 8;
constexpr int kWorkBins = 8 * kWorkBuckets * kWorkSub;
constexpr int kCellsNbCap = 32; // neighbours the thread-per-cell sweep kernel can stage; cells with more go to the tile kernel
constexpr int kBinCount = 0;                              // [kWorkBins] cells per bin
constexpr int kBinStart = kWorkBins;                      // [kWorkBins + 1] exclusive prefix
constexpr int kBinFill = kBinStart + kWorkBins + 1;       // [kWorkBins] scatter cursors
constexpr int kWorkRanges = kBinFill + kWorkBins;         // [8][6] per colour: {begin, end} of the crowded cells (8 lanes per cell or the tile kernel),
                                                          // of the middle range (4 lanes per cell) and of the sparsely populated cells (one thread per cell)
constexpr int kBinInts = kWorkRanges + 48;

struct CellBuffers {
	int *cell_of;     // N: cell of each particle in source order
	int *cell_sorted; // N: cell of each slot after the sort
	int *cell_count;  // n_cells
	int *cell_fill;   // n_cells
	int *cell_start;  // n_cells + 1
	int *tile_sums;   // scan scratch
	int *perm;        // N
	int *work;        // n_cells: non-empty cells grouped by colour, occupancy descending
	int *bins;        // kBinInts
	uint4 *posf;      // N: fixed-point shadow of each slot: box fractions x, y, z (32 bits each) + the slot (pre-screens only)
};

// A pre-generated roto-translation proposal (MC.c:312-324): rotation matrix R (new = old . R),
// displacement, Metropolis uniform and the rank of the picked particle inside its cell.
struct Proposal {
	double R[9];
	double dx, dy, dz;
	double u_acc;
	int pick;
	unsigned int wfix; // weight of the trial in the acceptance estimator, fixed point 2^-24
};

// one proposal of the K5 evaluator (device layout)
struct ProposalIO {
	const int *idx;          // particle ids
	const double *disp;      // n x 3
	const double *orient;    // n x 9
	const double *u;         // n
	int *dE, *overlap, *accept, *cell_changed;
};

// K1: returns the number of kernels launched.  cells_max_occupancy: cells with at most that many members go to
// the thread-per-cell sweep kernel, the more crowded ones to the tile kernel (0: everything to the tile kernel)
// cells_mid_occupancy >= cells_max_occupancy: cells with more members than cells_max_occupancy and at most cells_mid_occupancy form the middle range
int launch_cell_build(const SimConst &sc, const CellBuffers &cb, const StateView &src, const StateView &dst, int *slot_of, int cells_max_occupancy,
		int cells_mid_occupancy, cudaStream_t st);

// K2a: proposals of all N trials of one sweep (trial t of a cell lives at index cell_start[cell] + t)
int launch_proposals_for_sweep(const SimConst &sc, const CellBuffers &cb, Proposal *proposals, cudaStream_t st);
// K2b: one colour pass of the checkerboard sweep over the work items [range[0], range[1]) (device-resident range):
// tile of `lanes` threads per cell ...
int launch_sweep_pass(const SimConst &sc, const int *range, int lanes, const CellBuffers &cb, const Proposal *proposals, double4 *pos, double *ori,
		Counters *counters, const HaloMirror &hm, cudaStream_t st);
// ... or one thread per cell (sparsely populated cells; needs at least 3 distinct cells per axis)
int launch_sweep_cells(const SimConst &sc, const int *range, long long max_items, int lanes, const CellBuffers &cb, const Proposal *proposals, double4 *pos,
		double *ori, Counters *counters, const HaloMirror &hm, cudaStream_t st);
// ... or the three configurations side by side in one launch: ranges = {8-lane begin, end, 4-lane begin, end, 1-lane begin, end}
int launch_sweep_fused(const SimConst &sc, const int *ranges, long long max_items, const CellBuffers &cb, const Proposal *proposals, double4 *pos,
		double *ori, Counters *counters, const HaloMirror &hm, cudaStream_t st);

// K3: full bond count + overlap flag into counters->bonds / counters->overlap (zeroed by the launcher)
// (posf: the shadows of the same slots; nullptr or a grid with fewer than four cells on an axis: the tile kernel without the FP32 screen)
// scratch: device memory for the queue of pairs that pass the FP32 screens (nullptr: tile kernel); a queue overflow is reported in
// counters->pad and the caller repeats the count with force_tile
int launch_total_bonds(const SimConst &sc, const int *cell_start, const int *cell_sorted, const double4 *pos, const double *ori, const uint4 *posf,
		Counters *counters, void *scratch, size_t scratch_bytes, bool force_tile, cudaStream_t st);
// per-particle bonded-neighbour counts (indexed by particle id)
int launch_bond_counts(const SimConst &sc, const int *cell_start, const double4 *pos, const double *ori, const int *pid, int *counts, int *neighbours,
		int max_nb, cudaStream_t st);

// K4: dst.xyz = src.xyz * factor (or / factor when divide != 0) per axis, id carried over
int launch_rescale(int N, const double4 *src, double4 *dst, double fx, double fy, double fz, int divide, cudaStream_t st);

// K5: proposal evaluator against the frozen configuration
int launch_proposals(const SimConst &sc, int n, const ProposalIO &io, const int *cell_start, const double4 *pos, const double *ori, const int *slot_of,
		cudaStream_t st);
int launch_commit(const SimConst &sc, int slot, const double *disp3, const double *orient9, double4 *pos, double *ori, uint4 *posf, cudaStream_t st);

// K6: Gram-Schmidt on every orientation matrix
int launch_orthonormalise(int N, double *ori, cudaStream_t st);

// x-slab decomposition (pp_slab.cu)
int launch_slab_classify(const SimConst &sc_new, int slot_begin, int slot_end, int x0, int n_owned_layers, const StateView &cur, const StateView &keep,
		double *out_left, double *out_right, int cap_records, int *counts, cudaStream_t st);
int launch_slab_unpack(int n, const double *records, int offset, const StateView &dst, cudaStream_t st);
int launch_slab_download(int slot_begin, int slot_end, const StateView &cur, int *ids, double *r, double *orient, cudaStream_t st);
// phase numbers between neighbouring engines: release store into the neighbours' control words / acquire spin on one's own
int launch_phase_signal(unsigned long long *peer_a, unsigned long long *peer_b, unsigned long long value, int *layout_dst, int layout_val, const int *status,
		cudaStream_t st);
int launch_phase_wait(const unsigned long long *flag_a, const unsigned long long *flag_b, unsigned long long need, int *status, cudaStream_t st);
int launch_records_push(const double *outbox, const int *counts, int which, int cap, double *peer_box, int *peer_count, unsigned long long *peer_flag,
		unsigned long long epoch, unsigned int *done, cudaStream_t st);
int launch_records_pull(const double *mbox, const int *my_count, const unsigned long long *flag, unsigned long long epoch, int cap, double *inbox,
		int *in_count, int *status, cudaStream_t st);
int launch_shadow_slots(uint4 *posf, int begin, int end, cudaStream_t st);
int launch_gather_ints(const int *src, const int *idx, int n, int *out, cudaStream_t st);

// upload / download helpers (caller order <-> slot order)
int launch_select_pack(int N, const double *r, const double *orient, double inv_box, int nx, int lo, int held, double4 *pos, double *ori, int *pid, int cap,
		int *hist, cudaStream_t st);
int launch_pack_state(int N, const double *r, const double *orient, const int *ids, double4 *pos, double *ori, int *pid, cudaStream_t st);
int launch_unpack_state(int N, const double4 *pos, const double *ori, const int *pid, double *r, double *orient, cudaStream_t st);

} // namespace pp
