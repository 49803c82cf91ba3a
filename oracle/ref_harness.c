/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin driver around the UNMODIFIED reference objects (compiled by oracle/Makefile
 * straight from /root/reference/*.c into oracle/_ref/).  It exposes the reference's own
 * functions on the hot path through a flat C interface so that tests can pin the CPU
 * restatement (pp_oracle.c) and the CUDA kernels against the reference itself:
 *
 *   total energy      = 0.5 * sum_i MC_energy(i)            (main.c:55-61, MC.c:766-784)
 *   proposal replay   = MC_energy / MC_rototraslate_particle / MC_rollback_particle
 *                       with the acceptance rule of MC.c:332   (MC.c:311-340)
 *   rotation proposal = utils_rotate_matrix                  (utils.c:204-209)
 *   whole moves       = MC_move_rototranslate fed by an interposed drand48 stream
 *
 * drand48/srand48 are defined here (the library is linked with -Bsymbolic so the
 * reference objects bind to these): they implement the POSIX 48-bit LCG
 * (X' = 0x5DEECE66D * X + 0xB mod 2^48, srand48(s): X = s<<16 | 0x330E) unless a recorded
 * stream has been installed with ref_set_stream().
 *
 * This file contains no reference source; it only calls the reference's public symbols.
 */
#include "defs.h"
#include "MC.h"
#include "output.h"
#include "parse_input.h"
#include "system.h"
#include "utils.h"
#include "cells.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- RNG interposition */
static uint64_t g_x48 = 0x1234ABCD330EULL;
static const double *g_stream = NULL;
static long g_stream_len = 0, g_stream_pos = 0;
static long g_draws = 0;

double drand48(void) {
	g_draws++;
	if(g_stream != NULL) {
		if(g_stream_pos >= g_stream_len) {
			fprintf(stderr, "ref_harness: recorded drand48 stream exhausted\n");
			abort();
		}
		return g_stream[g_stream_pos++];
	}
	g_x48 = (0x5DEECE66DULL * g_x48 + 0xBULL) & 0xFFFFFFFFFFFFULL;
	return (double) g_x48 / 281474976710656.0;
}

void srand48(long seed) {
	g_x48 = ((((uint64_t) seed) << 16) | 0x330EULL) & 0xFFFFFFFFFFFFULL;
}

void ref_set_stream(const double *stream, long n) {
	g_stream = stream;
	g_stream_len = n;
	g_stream_pos = 0;
}
long ref_stream_pos(void) { return g_stream_pos; }
long ref_draws(void) { return g_draws; }
void ref_seed(long seed) { srand48(seed); }

/* ---------------------------------------------------------------- state */
static System g_syst;
static Output g_out;
static input_file g_input;
static int g_open = 0;

int ref_open(const char *input_path) {
	if(g_open) return -1;
	memset(&g_syst, 0, sizeof(g_syst));
	memset(&g_out, 0, sizeof(g_out));
	g_out.log = stderr;
	loadInputFile(&g_input, input_path);
	if(g_input.state == ERROR) return -2;
	system_init(&g_input, &g_syst, &g_out);
	MC_init(&g_input, &g_syst, &g_out);
	g_open = 1;
	return 0;
}

void ref_close(void) {
	if(!g_open) return;
	MC_free(&g_syst);
	system_free(&g_syst);
	cleanInputFile(&g_input);
	g_open = 0;
}

int ref_N(void) { return g_syst.N; }
void ref_box(double *box) { memcpy(box, g_syst.box, 3 * sizeof(double)); }
void ref_cells(int *n_side) { memcpy(n_side, g_syst.cells->N_side, 3 * sizeof(int)); }
double ref_running_energy(void) { return g_syst.energy; }
void ref_set_running_energy(double e) { g_syst.energy = e; }
void ref_counters(int *tries, int *accepted) {
	memcpy(tries, g_syst.tries, N_MOVES * sizeof(int));
	memcpy(accepted, g_syst.accepted, N_MOVES * sizeof(int));
}

void ref_get_state(double *r, double *orient) {
	int i;
	for(i = 0; i < g_syst.N; i++) {
		memcpy(r + 3 * i, g_syst.particles[i].r, 3 * sizeof(double));
		memcpy(orient + 9 * i, g_syst.particles[i].orientation, 9 * sizeof(double));
	}
}

void ref_get_patches(double *patches) {
	int i;
	for(i = 0; i < g_syst.N; i++) memcpy(patches + 12 * i, g_syst.particles[i].patches, 12 * sizeof(double));
}

/* overwrite positions/orientations (N unchanged), recompute patches, rebuild cells */
void ref_set_state(const double *r, const double *orient, const double *box) {
	int i;
	if(box != NULL) {
		memcpy(g_syst.box, box, 3 * sizeof(double));
		g_syst.V = box[0] * box[1] * box[2];
	}
	for(i = 0; i < g_syst.N; i++) {
		PatchyParticle *p = g_syst.particles + i;
		memcpy(p->r, r + 3 * i, 3 * sizeof(double));
		memcpy(p->orientation, orient + 9 * i, 9 * sizeof(double));
		set_patches(&g_syst, p);
	}
	cells_free(g_syst.cells);
	cells_init(&g_syst, &g_out, g_syst.r_cut);
	cells_fill(&g_syst);
}

/* main.c:55-61 -- returns energy (= -bonds); *overlap set if any MC_energy flagged one */
double ref_total_energy(int *overlap) {
	int i;
	double E = 0.;
	*overlap = 0;
	for(i = 0; i < g_syst.N; i++) {
		E += MC_energy(&g_syst, g_syst.particles + i);
		if(g_syst.overlap) *overlap = 1;
	}
	g_syst.overlap = 0;
	return 0.5 * E;
}

/* per-particle bonded-neighbour counts (what Print_bonds reports, output.c:213-268) */
void ref_bond_counts(int *counts, int *overlaps) {
	int i;
	for(i = 0; i < g_syst.N; i++) {
		double e = MC_energy(&g_syst, g_syst.particles + i);
		counts[i] = (int) lrint(-e);
		overlaps[i] = g_syst.overlap;
	}
	g_syst.overlap = 0;
}

/* bonded neighbours of every particle by the reference's own pair function, over ALL pairs (no cells): an
 * independent statement of what output_print_bonds (output.c:213-268) is meant to list.  O(N^2): test sizes only. */
void ref_bond_lists(int max_nb, int *counts, int *neigh) {
	int i, j, onp, onq;
	for(i = 0; i < g_syst.N; i++) {
		int n = 0;
		for(j = 0; j < g_syst.N; j++) {
			if(j == i) continue;
			if(MC_interact(&g_syst, g_syst.particles + i, g_syst.particles + j, &onp, &onq) == PATCH_BOND) {
				if(neigh != NULL && n < max_nb) neigh[(size_t) i * max_nb + n] = j;
				n++;
			}
		}
		counts[i] = n;
	}
}

/* what AVB's _set_neighbours (avb.c:43-75) collects, by the same MC_interact over ALL pairs in ascending index: per particle the
 * bonded neighbours with the patch of p (onp) and of q (onq) the reference reports */
void ref_patch_pairs(int max_nb, int *counts, int *q_idx, int *q_onp, int *q_onq) {
	int i, j, onp, onq;
	for(i = 0; i < g_syst.N; i++) {
		int n = 0;
		for(j = 0; j < g_syst.N; j++) {
			if(j == i) continue;
			if(MC_interact(&g_syst, g_syst.particles + i, g_syst.particles + j, &onp, &onq) == PATCH_BOND) {
				if(n < max_nb) {
					q_idx[(size_t) i * max_nb + n] = j;
					q_onp[(size_t) i * max_nb + n] = onp;
					q_onq[(size_t) i * max_nb + n] = onq;
				}
				n++;
			}
		}
		counts[i] = n;
	}
}

void ref_rotate(const double *orient_old, const double *axis, double theta, double *orient_new) {
	matrix o, n;
	vector a = {axis[0], axis[1], axis[2]};
	memcpy(o, orient_old, sizeof(matrix));
	utils_rotate_matrix(o, n, a, theta);
	memcpy(orient_new, n, sizeof(matrix));
}

/*
 * Replay a list of proposals with the reference's own functions (MC.c:328-339).
 * commit == 0: every proposal is evaluated against the same frozen configuration (always rolled back).
 * commit != 0: accepted proposals are kept, i.e. the list is a sequential move list.
 * u[k] is the uniform the reference would draw for the Metropolis test (only consulted when the
 * reference would draw it: no overlap and deltaE >= 0).
 * dE_out is the reference's local deltaE (E_new - E_old with E_new = 0 when the new position overlaps).
 */
void ref_eval_proposals(int n, const int *idx, const double *disp, const double *orient_new, const double *u,
		int commit, int *dE_out, int *overlap_out, int *accept_out) {
	int k;
	for(k = 0; k < n; k++) {
		PatchyParticle *p = g_syst.particles + idx[k];
		vector d = {disp[3 * k], disp[3 * k + 1], disp[3 * k + 2]};
		matrix o;
		memcpy(o, orient_new + 9 * k, sizeof(matrix));

		double deltaE = -MC_energy(&g_syst, p);
		MC_rototraslate_particle(&g_syst, p, d, o);
		deltaE += MC_energy(&g_syst, p);
		int ov = g_syst.overlap;
		int acc = (!ov && (deltaE < 0. || u[k] < exp(-deltaE / g_syst.T)));

		dE_out[k] = (int) lrint(deltaE);
		overlap_out[k] = ov;
		accept_out[k] = acc;

		if(acc && commit) {
			g_syst.energy += deltaE;
		}
		else {
			MC_rollback_particle(&g_syst, p);
		}
		g_syst.overlap = 0;
	}
}

/* run whole reference moves (MC_move_rototranslate, MC.c:311-340) n times, drawing from drand48 above */
void ref_run_moves(long n) {
	long k;
	for(k = 0; k < n; k++) MC_move_rototranslate(&g_syst, &g_out);
}

/* run n full ensemble steps (do_NVT / do_NPT, main.c:119) */
void ref_run_steps(long n) {
	long k;
	for(k = 0; k < n; k++) g_syst.do_ensemble(&g_syst, &g_out);
}

void ref_check_energy(void) { MC_check_energy(&g_syst, &g_out); }
void ref_volume_move(void) { MC_change_volume(&g_syst, &g_out); }
/* one MC_add_remove (MC.c:342-448) of a system opened with Ensemble = 1 (GC_N_max particles allocated, system.c:94, 199) */
void ref_add_remove(void) { MC_add_remove(&g_syst, &g_out); }
void ref_Lx_move(void) { MC_change_Lx(&g_syst, &g_out); }

/* the reference's own helpers, driven directly (utils.c:115-131 draws from the drand48 above; utils.c:91-113) */
void ref_random_vector_on_sphere(double *out) {
	vector v;
	random_vector_on_sphere(v);
	memcpy(out, v, sizeof(vector));
}

void ref_gram_schmidt(double *m) {
	gram_schmidt(m, m + 3, m + 6);
}

/* per-move bookkeeping the tests read after ref_volume_move / ref_Lx_move */
double ref_volume(void) { return g_syst.V; }
