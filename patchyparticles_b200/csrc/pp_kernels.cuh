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

struct CellBuffers {
	int *cell_of;    // N
	int *cell_count; // n_cells
	int *cell_fill;  // n_cells
	int *cell_start; // n_cells + 1
	int *tile_sums;  // scan scratch
	int *perm;       // N
};

// one proposal of the K5 evaluator (device layout)
struct ProposalIO {
	const int *idx;          // particle ids
	const double *disp;      // n x 3
	const double *orient;    // n x 9
	const double *u;         // n
	int *dE, *overlap, *accept, *cell_changed;
};

// K1: returns the number of kernels launched
int launch_cell_build(const SimConst &sc, const CellBuffers &cb, const StateView &src, const StateView &dst, int *slot_of, cudaStream_t st);

// K2: one colour pass of the checkerboard sweep
int launch_sweep_pass(const SimConst &sc, int colour, int lanes, const int *cell_start, double4 *pos, double *ori, Counters *counters, cudaStream_t st);

// K3: full bond count + overlap flag into counters->bonds / counters->overlap (zeroed by the launcher)
int launch_total_bonds(const SimConst &sc, const int *cell_start, const double4 *pos, const double *ori, Counters *counters, cudaStream_t st);
// per-particle bonded-neighbour counts (indexed by particle id)
int launch_bond_counts(const SimConst &sc, const int *cell_start, const double4 *pos, const double *ori, const int *pid, int *counts, cudaStream_t st);

// K4: dst.xyz = src.xyz * factor (or / factor when divide != 0) per axis, id carried over
int launch_rescale(int N, const double4 *src, double4 *dst, double fx, double fy, double fz, int divide, cudaStream_t st);

// K5: proposal evaluator against the frozen configuration
int launch_proposals(const SimConst &sc, int n, const ProposalIO &io, const int *cell_start, const double4 *pos, const double *ori, const int *slot_of,
		cudaStream_t st);
int launch_commit(int slot, const double *disp3, const double *orient9, double4 *pos, double *ori, cudaStream_t st);

// K6: Gram-Schmidt on every orientation matrix
int launch_orthonormalise(int N, double *ori, cudaStream_t st);

// upload / download helpers (caller order <-> slot order)
int launch_pack_state(int N, const double *r, const double *orient, double4 *pos, double *ori, int *pid, cudaStream_t st);
int launch_unpack_state(int N, const double4 *pos, const double *ori, const int *pid, double *r, double *orient, cudaStream_t st);

} // namespace pp
