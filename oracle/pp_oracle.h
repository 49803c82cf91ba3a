/*
 * pp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, FP64, no FMA contraction) of the PatchyParticles hot path:
 * Kern-Frenkel pair test, per-particle and total bond counts over cell lists, roto-translation
 * proposal evaluation with the reference's Metropolis rule, volume / Lx move arithmetic, insertion / removal, and a
 * serial restatement of this repo's checkerboard sweep (same Philox streams as the CUDA kernel).
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 * the unmodified reference built into oracle/_ref/libppref.so, and against the golden vectors in
 * tests/golden/ that were generated from that library (tools/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libpatchy_b200.so, host/) never does.
 */
#ifndef PP_ORACLE_H_
#define PP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ppo_system ppo_system;

/* ---- construction ---------------------------------------------------------------------- */
/* r: N x 3, orient: N x 3 x 3 (row-major, rows as in the .rrr file + their cross product) */
ppo_system *ppo_create(int N, const double *box, const double *r, const double *orient,
		double kf_delta, double kf_cosmax, double temperature);
void ppo_destroy(ppo_system *s);

/* reads a .rrr file the way system.c:80,213-245 does (normalise rows, cross, Gram-Schmidt).
 * On success returns 0 and mallocs *r (N*3) and *orient (N*9); caller frees with ppo_free. */
int ppo_read_rrr(const char *path, long long *step, int *N, double *box, double **r, double **orient);
int ppo_write_rrr(const char *path, long long step, int N, const double *box, const double *r, const double *orient);
void ppo_free(void *p);

/* ---- small pure functions ------------------------------------------------------------- */
void ppo_base_patches(double *base /* 4 x 3 */);                                   /* system.c:20-31 */
void ppo_orientation_from_rows(const double *row0, const double *row1, double *m); /* system.c:225-233 */
void ppo_patches_from_orientation(const double *m, double *patches /* 4 x 3 */);  /* MC.c:225-227 */
void ppo_gram_schmidt(double *m);                                                  /* utils.c:91-113 */
void ppo_rotate(const double *m_old, const double *axis, double theta, double *m_new); /* utils.c:186-209 */
void ppo_axis_from_pair(double ran1, double ran2, double *axis);                   /* utils.c:127-131 */
int ppo_cell_coords(const double *r, const double *box, const int *n_side, int *idx); /* cells.c:43-49 */
/* -1 overlap, 0 no bond, 1 bond */
int ppo_pair(const double *rp, const double *patches_p, const double *rq, const double *patches_q,
		const double *box, double sqr_rcut, double cosmax);                         /* MC.c:232-266 */

/* ---- system-level -------------------------------------------------------------------- */
int ppo_N(const ppo_system *s);
void ppo_get_box(const ppo_system *s, double *box);
void ppo_get_cells(const ppo_system *s, int *n_side);
void ppo_get_state(const ppo_system *s, double *r, double *orient);
void ppo_set_state(ppo_system *s, const double *r, const double *orient, const double *box);

/* number of bonded neighbours of particle i at its current state; *overlap = 1 if any */
int ppo_particle_bonds(ppo_system *s, int i, int *overlap);                        /* MC.c:272-309 */
/* number of bonded PAIRS in the system (energy = -bonds), *overlap = 1 if any pair overlaps */
long long ppo_total_bonds(ppo_system *s, int *overlap);                            /* main.c:55-61 */
void ppo_bond_counts(ppo_system *s, int *counts, int *overlaps);
void ppo_bond_lists(ppo_system *s, int max_nb, int *counts, int *neigh);                   /* output.c:213-268 */

/* same contract as ref_eval_proposals in ref_harness.c                          MC.c:328-339 */
void ppo_eval_proposals(ppo_system *s, int n, const int *idx, const double *disp, const double *orient_new,
		const double *u, int commit, int *dE_out, int *overlap_out, int *accept_out);

/* volume move arithmetic (MC.c:450-513) given its three uniforms; returns 1 if accepted.
 * energy_bonds is the running bond count (updated on accept).  dir_out/new box reported. */
int ppo_volume_move(ppo_system *s, double P, double rescale_factor_max, double u_dir, double u_vol,
		double u_acc, long long *energy_bonds, int *overlap_out);
/* Lx move (MC.c:515-585) given its two uniforms; returns 1 accepted, 0 rejected, -1 early reject */
int ppo_Lx_move(ppo_system *s, double Lx_change_max, double Lyz_min, double Lyz_max, int faithful_sign,
		double u_lx, double u_acc, long long *energy_bonds, int *overlap_out);

/* neighbour queries of the cluster moves: mode 0, _set_neighbours (avb.c:43-75): list[4] = neighbour bonded through each patch, returns
 * the number of bonded neighbours; mode 1, the range test of _populate_possible_links (vmmc.c:178): list[max_nb] = particles in range */
int ppo_neighbour_query(ppo_system *s, int i, int mode, int max_nb, int *list);

/* MC_add_remove (MC.c:342-448) given its uniforms in the reference's draw order; returns 2 (insertion), 3 (removal) or -1 */
int ppo_add_remove(ppo_system *s, double z, int n_min, int n_max, const double *u, int n_u, int *accepted, int *used, long long *energy_bonds);

/* ---- serial restatement of this repo's checkerboard sweep ----------------------------- */
typedef struct {
	long long tries, accepted, confined, overlapped;
	uint64_t w_tries, w_accepted; /* fixed point 2^-24, see DESIGN.md "acceptance estimator" */
	long long pair_checks;
} ppo_sweep_stats;

/* one sweep = 8 colour passes on an even grid of side >= r_cut shifted by shift[3];
 * RNG = Philox4x32-10 keyed by seed, counter (sweep, cell, trial, draw block).
 * Updates s in place; *energy_bonds is the running bond count. */
void ppo_sweep(ppo_system *s, uint64_t seed, uint64_t sweep_index, int shift_every, double disp_max, double theta_max,
		long long *energy_bonds, ppo_sweep_stats *stats);
void ppo_even_grid(const double *box, double r_cut, int *n_side);
void ppo_sweep_shift(uint64_t seed, uint64_t sweep_index, const double *box, const int *n_side, double *shift);
void ppo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
