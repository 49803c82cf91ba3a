/*
 * pp_oracle.c -- TEST INFRASTRUCTURE ONLY (see pp_oracle.h).
 *
 * CPU restatement of the PatchyParticles roto-translation hot path.  Plain C, FP64,
 * compiled with -ffp-contract=off so that every product/sum is rounded separately,
 * exactly like the strict flavour of the reference in oracle/_ref/libppref.so.
 *
 * Data layout differs from the reference on purpose (flat arrays + CSR cell lists
 * instead of an array of structs with linked lists); the arithmetic of each formula
 * follows the cited reference lines operation by operation, because bit-exact bond
 * counts and accept/reject decisions are the contract.
 */
#include "pp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct ppo_system {
	int N;
	double box[3];
	double V;        /* System.V: set from the box at start-up (system.c:198), replaced by an accepted volume move (MC.c:500) */
	double T;
	double kf_delta, kf_cosmax, sqr_rcut, r_cut;
	double base[12];
	double *r;       /* N x 3 */
	double *m;       /* N x 9 */
	double *patches; /* N x 12, always = m . base */
	/* CSR cell list on the reference grid (cells.c:17-37): floor(box / r_cut), at least 3 */
	int n_side[3];
	int n_cells;
	int *cell_start; /* n_cells + 1 */
	int *cell_items; /* N */
	int cells_valid;
};

/* ------------------------------------------------------------------ small helpers */
static double dot3(const double *a, const double *b) {
	/* utils.h:15 SCALAR: ((a0*b0 + a1*b1) + a2*b2) */
	return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

static void normalize3(double *v) {
	/* utils.c:31-36 */
	double n = sqrt(dot3(v, v));
	v[0] /= n;
	v[1] /= n;
	v[2] /= n;
}

static void matvec3(const double *m, const double *v, double *out) {
	/* utils.h:17-21 */
	out[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
	out[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
	out[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}

void ppo_free(void *p) {
	free(p);
}

void ppo_base_patches(double *base) {
	/* system.c:20-31: four tetrahedral directions, built from 0.5/sqrt(3) then normalised */
	static const int sgn[4][3] = {{-1, -1, 1}, {1, -1, -1}, {1, 1, 1}, {-1, 1, -1}};
	double h = 0.5 / sqrt(3);
	int i, k;
	for(i = 0; i < 4; i++) {
		for(k = 0; k < 3; k++) base[3 * i + k] = (sgn[i][k] > 0) ? h : -h;
		normalize3(base + 3 * i);
	}
}

void ppo_gram_schmidt(double *m) {
	/* utils.c:91-113 on the three rows of m */
	double *v1 = m, *v2 = m + 3, *v3 = m + 6;
	double v1_norm2 = dot3(v1, v1);
	double v2_v1 = dot3(v2, v1);
	double b1 = v2_v1 / v1_norm2;
	int k;
	for(k = 0; k < 3; k++) v2[k] -= b1 * v1[k];

	double v3_v1 = dot3(v3, v1);
	double v3_v2 = dot3(v3, v2);
	double v2_norm2 = dot3(v2, v2);
	b1 = v3_v1 / v1_norm2;
	double b2 = v3_v2 / v2_norm2;
	for(k = 0; k < 3; k++) v3[k] -= b1 * v1[k] + b2 * v2[k];

	normalize3(v1);
	normalize3(v2);
	normalize3(v3);
}

void ppo_orientation_from_rows(const double *row0, const double *row1, double *m) {
	/* system.c:225-233: normalise both rows, third = row0 x row1 (utils.c:25-29), Gram-Schmidt */
	memcpy(m, row0, 3 * sizeof(double));
	memcpy(m + 3, row1, 3 * sizeof(double));
	normalize3(m);
	normalize3(m + 3);
	m[6] = m[1] * m[5] - m[2] * m[4];
	m[7] = m[2] * m[3] - m[0] * m[5];
	m[8] = m[0] * m[4] - m[1] * m[3];
	ppo_gram_schmidt(m);
}

static void patches_from(const double *base, const double *m, double *patches) {
	int i;
	for(i = 0; i < 4; i++) matvec3(m, base + 3 * i, patches + 3 * i);
}

void ppo_patches_from_orientation(const double *m, double *patches) {
	double base[12];
	ppo_base_patches(base);
	patches_from(base, m, patches);
}

void ppo_rotate(const double *o, const double *axis, double t, double *out) {
	/* utils.c:186-202 Rodrigues matrix, then utils.c:172-184 out = o . R */
	double st = sin(t), ct = cos(t), uct = 1. - ct;
	double x = axis[0], y = axis[1], z = axis[2];
	double R[9];
	R[0] = x * x + (1. - x * x) * ct;
	R[1] = x * y * uct - z * st;
	R[2] = x * z * uct + y * st;
	R[3] = x * y * uct + z * st;
	R[4] = y * y + (1. - y * y) * ct;
	R[5] = y * z * uct - x * st;
	R[6] = x * z * uct - y * st;
	R[7] = y * z * uct + x * st;
	R[8] = z * z + (1. - z * z) * ct;
	int i, j;
	for(i = 0; i < 3; i++)
		for(j = 0; j < 3; j++) out[3 * i + j] = o[3 * i] * R[j] + o[3 * i + 1] * R[3 + j] + o[3 * i + 2] * R[6 + j];
}

void ppo_axis_from_pair(double ran1, double ran2, double *axis) {
	/* utils.c:127-131 (the accepted Marsaglia pair) */
	double ransq = ran1 * ran1 + ran2 * ran2;
	double ranh = 2. * sqrt(1. - ransq);
	axis[0] = ran1 * ranh;
	axis[1] = ran2 * ranh;
	axis[2] = 1. - 2. * ransq;
}

int ppo_cell_coords(const double *r, const double *box, const int *n_side, int *idx) {
	/* cells.c:43-49 */
	int k;
	for(k = 0; k < 3; k++) idx[k] = (int) ((r[k] / box[k] - floor(r[k] / box[k])) * (1. - DBL_EPSILON) * n_side[k]);
	return (idx[0] * n_side[1] + idx[1]) * n_side[2] + idx[2];
}

int ppo_pair(const double *rp, const double *patches_p, const double *rq, const double *patches_q, const double *box, double sqr_rcut, double cosmax) {
	/* MC.c:232-266 */
	double d[3] = {rq[0] - rp[0], rq[1] - rp[1], rq[2] - rp[2]};
	int k;
	for(k = 0; k < 3; k++) d[k] -= box[k] * rint(d[k] / box[k]);
	double d2 = dot3(d, d);
	if(d2 < 1.) return -1;
	if(d2 < sqr_rcut) {
		double norm = sqrt(d2);
		d[0] /= norm;
		d[1] /= norm;
		d[2] /= norm;
		int pp, pq;
		for(pp = 0; pp < 4; pp++) {
			if(dot3(d, patches_p + 3 * pp) > cosmax) {
				for(pq = 0; pq < 4; pq++) {
					if(-dot3(d, patches_q + 3 * pq) > cosmax) return 1;
				}
			}
		}
	}
	return 0;
}

/* ------------------------------------------------------------------ cell lists */
static void cells_release(ppo_system *s) {
	free(s->cell_start);
	free(s->cell_items);
	s->cell_start = NULL;
	s->cell_items = NULL;
	s->cells_valid = 0;
}

static void cells_build(ppo_system *s) {
	/* grid of cells.c:17-37, filled as a counting sort instead of linked lists */
	int k, i;
	cells_release(s);
	for(k = 0; k < 3; k++) {
		s->n_side[k] = (int) floor(s->box[k] / s->r_cut);
		if(s->n_side[k] < 3) s->n_side[k] = 3;
	}
	s->n_cells = s->n_side[0] * s->n_side[1] * s->n_side[2];
	s->cell_start = calloc((size_t) s->n_cells + 1, sizeof(int));
	s->cell_items = malloc(sizeof(int) * (size_t) (s->N > 0 ? s->N : 1));
	int *cell_of = malloc(sizeof(int) * (size_t) (s->N > 0 ? s->N : 1));
	int idx[3];
	for(i = 0; i < s->N; i++) {
		cell_of[i] = ppo_cell_coords(s->r + 3 * i, s->box, s->n_side, idx);
		s->cell_start[cell_of[i] + 1]++;
	}
	for(i = 0; i < s->n_cells; i++) s->cell_start[i + 1] += s->cell_start[i];
	int *fill = calloc((size_t) s->n_cells, sizeof(int));
	for(i = 0; i < s->N; i++) {
		int c = cell_of[i];
		s->cell_items[s->cell_start[c] + fill[c]++] = i;
	}
	free(fill);
	free(cell_of);
	s->cells_valid = 1;
}

/* ------------------------------------------------------------------ system */
ppo_system *ppo_create(int N, const double *box, const double *r, const double *orient, double kf_delta, double kf_cosmax, double temperature) {
	ppo_system *s = calloc(1, sizeof(ppo_system));
	s->N = N;
	memcpy(s->box, box, 3 * sizeof(double));
	s->V = box[0] * box[1] * box[2];
	s->T = temperature;
	s->kf_delta = kf_delta;
	s->kf_cosmax = kf_cosmax;
	s->sqr_rcut = (1. + kf_delta) * (1. + kf_delta); /* system.c:86 */
	s->r_cut = 1. + kf_delta;                       /* system.c:249 */
	ppo_base_patches(s->base);
	size_t n = (size_t) (N > 0 ? N : 1);
	s->r = malloc(n * 3 * sizeof(double));
	s->m = malloc(n * 9 * sizeof(double));
	s->patches = malloc(n * 12 * sizeof(double));
	ppo_set_state(s, r, orient, NULL);
	return s;
}

void ppo_destroy(ppo_system *s) {
	if(s == NULL) return;
	cells_release(s);
	free(s->r);
	free(s->m);
	free(s->patches);
	free(s);
}

int ppo_N(const ppo_system *s) {
	return s->N;
}
void ppo_get_box(const ppo_system *s, double *box) {
	memcpy(box, s->box, 3 * sizeof(double));
}
void ppo_get_cells(const ppo_system *s, int *n_side) {
	memcpy(n_side, s->n_side, 3 * sizeof(int));
}
void ppo_get_state(const ppo_system *s, double *r, double *orient) {
	memcpy(r, s->r, (size_t) s->N * 3 * sizeof(double));
	memcpy(orient, s->m, (size_t) s->N * 9 * sizeof(double));
}

void ppo_set_state(ppo_system *s, const double *r, const double *orient, const double *box) {
	int i;
	if(box != NULL) {
		memcpy(s->box, box, 3 * sizeof(double));
		s->V = box[0] * box[1] * box[2];
	}
	memcpy(s->r, r, (size_t) s->N * 3 * sizeof(double));
	memcpy(s->m, orient, (size_t) s->N * 9 * sizeof(double));
	for(i = 0; i < s->N; i++) patches_from(s->base, s->m + 9 * i, s->patches + 12 * i);
	cells_build(s);
}

/* bonds of a probe (position rp, patches pp) standing in for particle `self`.   MC.c:272-309 */
static int probe_bonds(const ppo_system *s, int self, const double *rp, const double *pp, int *overlap, long long *checks) {
	int idx[3], j, k, l, bonds = 0;
	*overlap = 0;
	ppo_cell_coords(rp, s->box, s->n_side, idx);
	for(j = -1; j < 2; j++) {
		int cx = (idx[0] + j + s->n_side[0]) % s->n_side[0];
		for(k = -1; k < 2; k++) {
			int cy = (idx[1] + k + s->n_side[1]) % s->n_side[1];
			for(l = -1; l < 2; l++) {
				int cz = (idx[2] + l + s->n_side[2]) % s->n_side[2];
				int c = (cx * s->n_side[1] + cy) * s->n_side[2] + cz;
				int a;
				for(a = s->cell_start[c]; a < s->cell_start[c + 1]; a++) {
					int q = s->cell_items[a];
					if(q == self) continue;
					if(checks) (*checks)++;
					int v = ppo_pair(rp, pp, s->r + 3 * q, s->patches + 12 * q, s->box, s->sqr_rcut, s->kf_cosmax);
					if(v == 1) bonds++;
					else if(v == -1) {
						*overlap = 1;
						return 0; /* the reference returns E = 0 at the first overlap */
					}
				}
			}
		}
	}
	return bonds;
}

int ppo_particle_bonds(ppo_system *s, int i, int *overlap) {
	return probe_bonds(s, i, s->r + 3 * i, s->patches + 12 * i, overlap, NULL);
}

long long ppo_total_bonds(ppo_system *s, int *overlap) {
	/* main.c:55-61 / MC.c:766-784: E = 0.5 * sum_i E_i; here in integer bonds */
	long long twice = 0;
	int i, ov;
	*overlap = 0;
	for(i = 0; i < s->N; i++) {
		twice += ppo_particle_bonds(s, i, &ov);
		if(ov) *overlap = 1;
	}
	return twice / 2;
}

void ppo_bond_counts(ppo_system *s, int *counts, int *overlaps) {
	int i;
	for(i = 0; i < s->N; i++) counts[i] = ppo_particle_bonds(s, i, overlaps + i);
}

/* the two lines output_print_bonds writes per particle (output.c:222-264): bonded-neighbour count and indices, walking
 * the 27 cells in the reference's order; no early exit on overlaps (the reference calls MC_interact pair by pair). */
void ppo_bond_lists(ppo_system *s, int max_nb, int *counts, int *neigh) {
	int i, j, k, l, idx[3];
	for(i = 0; i < s->N; i++) {
		int n = 0;
		ppo_cell_coords(s->r + 3 * i, s->box, s->n_side, idx);
		for(j = -1; j < 2; j++) {
			int cx = (idx[0] + j + s->n_side[0]) % s->n_side[0];
			for(k = -1; k < 2; k++) {
				int cy = (idx[1] + k + s->n_side[1]) % s->n_side[1];
				for(l = -1; l < 2; l++) {
					int cz = (idx[2] + l + s->n_side[2]) % s->n_side[2];
					int c = (cx * s->n_side[1] + cy) * s->n_side[2] + cz, a;
					for(a = s->cell_start[c]; a < s->cell_start[c + 1]; a++) {
						int q = s->cell_items[a];
						if(q == i) continue;
						if(ppo_pair(s->r + 3 * i, s->patches + 12 * i, s->r + 3 * q, s->patches + 12 * q, s->box, s->sqr_rcut, s->kf_cosmax) == 1) {
							if(neigh != NULL && n < max_nb) neigh[(size_t) i * max_nb + n] = q;
							n++;
						}
					}
				}
			}
		}
		counts[i] = n;
	}
}

/* move particle i between CSR cells (the reference relinks its lists, MC.c:180-206) */
static void cells_move(ppo_system *s, int i, const double *r_new) {
	int idx[3];
	int c_old = ppo_cell_coords(s->r + 3 * i, s->box, s->n_side, idx);
	int c_new = ppo_cell_coords(r_new, s->box, s->n_side, idx);
	memcpy(s->r + 3 * i, r_new, 3 * sizeof(double));
	if(c_old != c_new) cells_build(s); /* simple and rare at test sizes */
}

void ppo_eval_proposals(ppo_system *s, int n, const int *idx, const double *disp, const double *orient_new, const double *u, int commit,
		int *dE_out, int *overlap_out, int *accept_out) {
	/* MC.c:328-339 with the proposal (disp, new orientation, uniform) supplied by the caller */
	int k;
	for(k = 0; k < n; k++) {
		int i = idx[k], ov_old, ov_new;
		double r_new[3] = {s->r[3 * i] + disp[3 * k], s->r[3 * i + 1] + disp[3 * k + 1], s->r[3 * i + 2] + disp[3 * k + 2]};
		double p_new[12];
		patches_from(s->base, orient_new + 9 * k, p_new);

		int b_old = probe_bonds(s, i, s->r + 3 * i, s->patches + 12 * i, &ov_old, NULL);
		int b_new = probe_bonds(s, i, r_new, p_new, &ov_new, NULL);
		double deltaE = -(double) b_new + (double) b_old;
		int acc = (!ov_new && (deltaE < 0. || u[k] < exp(-deltaE / s->T)));

		dE_out[k] = b_old - b_new;
		overlap_out[k] = ov_new;
		accept_out[k] = acc;
		if(acc && commit) {
			cells_move(s, i, r_new);
			memcpy(s->m + 9 * i, orient_new + 9 * k, 9 * sizeof(double));
			memcpy(s->patches + 12 * i, p_new, 12 * sizeof(double));
		}
	}
}

/* ------------------------------------------------------------------ box moves */
static int grid_too_fine(const ppo_system *s, int k) {
	return (s->box[k] / s->n_side[k]) < s->r_cut;
}

int ppo_volume_move(ppo_system *s, double P, double rescale_factor_max, double u_dir, double u_vol, double u_acc, long long *energy_bonds,
		int *overlap_out) {
	/* MC.c:450-513 */
	int i, k;
	double initial_V = s->box[0] * s->box[1] * s->box[2];
	int dir = (int) (u_dir * 3);
	double ln_final_V = log(initial_V) + (u_vol - 0.5) * rescale_factor_max;
	double final_V = exp(ln_final_V);
	double old_side = s->box[dir];
	double new_side = final_V;
	for(k = 0; k < 3; k++)
		if(k != dir) new_side /= s->box[k];
	s->box[dir] = new_side;
	double f = new_side / old_side;
	for(i = 0; i < s->N; i++) s->r[3 * i + dir] *= f;
	if(grid_too_fine(s, dir)) cells_build(s);

	int ov;
	long long bonds_new = ppo_total_bonds(s, &ov);
	double delta_E = -(double) bonds_new + (double) *energy_bonds;
	double delta_V = final_V - initial_V;
	double arg = -(P * delta_V + delta_E) / s->T + (s->N + 1) * log(f);
	*overlap_out = ov;
	if(!ov && u_acc < exp(arg)) {
		s->V = final_V; /* MC.c:500 */
		*energy_bonds = bonds_new;
		return 1;
	}
	for(i = 0; i < s->N; i++) s->r[3 * i + dir] /= f;
	s->box[dir] = old_side;
	return 0;
}

int ppo_Lx_move(ppo_system *s, double Lx_change_max, double Lyz_min, double Lyz_max, int faithful_sign, double u_lx, double u_acc,
		long long *energy_bonds, int *overlap_out) {
	/* MC.c:515-585.  faithful_sign != 0 reproduces the reference's exp(+dE/T) test (MC.c:562). */
	int i, k;
	double old[3] = {s->box[0], s->box[1], s->box[2]};
	*overlap_out = 0;
	s->box[0] += (u_lx - 0.5) * Lx_change_max;
	s->box[1] = s->box[2] = sqrt(s->V / s->box[0]); /* MC.c:524: the stored volume, not the product of the current sides */
	if(s->box[1] < Lyz_min || s->box[1] > Lyz_max) {
		memcpy(s->box, old, sizeof(old));
		return -1;
	}
	double f[3] = {s->box[0] / old[0], s->box[1] / old[1], s->box[2] / old[2]};
	for(i = 0; i < s->N; i++)
		for(k = 0; k < 3; k++) s->r[3 * i + k] *= f[k];
	if(grid_too_fine(s, 0) || grid_too_fine(s, 1)) cells_build(s);

	int ov;
	long long bonds_new = ppo_total_bonds(s, &ov);
	double delta_E = -(double) bonds_new + (double) *energy_bonds;
	*overlap_out = ov;
	double arg = faithful_sign ? (delta_E / s->T) : (-delta_E / s->T);
	if(!ov && u_acc < exp(arg)) {
		*energy_bonds = bonds_new;
		return 1;
	}
	for(i = 0; i < s->N; i++)
		for(k = 0; k < 3; k++) s->r[3 * i + k] /= f[k];
	memcpy(s->box, old, sizeof(old));
	if(grid_too_fine(s, 0) || grid_too_fine(s, 1)) cells_build(s);
	return 0;
}

/* ------------------------------------------------------------------ .rrr I/O */
int ppo_read_rrr(const char *path, long long *step, int *N, double *box, double **r_out, double **m_out) {
	/* system.c:80 header, system.c:213-245 body */
	FILE *f = fopen(path, "r");
	if(f == NULL) return -1;
	long long st;
	int n;
	if(fscanf(f, "%lld %d %lf %lf %lf", &st, &n, box, box + 1, box + 2) != 5) {
		fclose(f);
		return -2;
	}
	double *r = malloc(sizeof(double) * 3 * (size_t) (n > 0 ? n : 1));
	double *m = malloc(sizeof(double) * 9 * (size_t) (n > 0 ? n : 1));
	int i;
	for(i = 0; i < n; i++) {
		double a[3], b[3];
		if(fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf", a, a + 1, a + 2, b, b + 1, b + 2, r + 3 * i, r + 3 * i + 1, r + 3 * i + 2) != 9) {
			free(r);
			free(m);
			fclose(f);
			return -3;
		}
		ppo_orientation_from_rows(a, b, m + 9 * i);
	}
	double extra;
	int more = (fscanf(f, "%lf", &extra) == 1);
	fclose(f);
	if(more) {
		free(r);
		free(m);
		return -4; /* more particles than the header says: system.c:245 is fatal */
	}
	if(step) *step = st;
	*N = n;
	*r_out = r;
	*m_out = m;
	return 0;
}

int ppo_write_rrr(const char *path, long long step, int N, const double *box, const double *r, const double *m) {
	/* output.c:302-317 */
	FILE *f = fopen(path, "w");
	if(f == NULL) return -1;
	fprintf(f, "%lld %d %.12lf %.12lf %.12lf\n", step, N, box[0], box[1], box[2]);
	int i;
	for(i = 0; i < N; i++) {
		fprintf(f, "%lf %lf %lf\n", m[9 * i], m[9 * i + 1], m[9 * i + 2]);
		fprintf(f, "%lf %lf %lf\n", m[9 * i + 3], m[9 * i + 4], m[9 * i + 5]);
		fprintf(f, "%.12lf %.12lf %.12lf\n", r[3 * i], r[3 * i + 1], r[3 * i + 2]);
	}
	fclose(f);
	return 0;
}

/* ------------------------------------------------------------------ checkerboard sweep (this repo's algorithm) */
void ppo_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
	/* Philox4x32-10 (Salmon et al., SC'11): public algorithm, constants from the paper */
	int round;
	for(round = 0; round < 10; round++) {
		uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
		uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t) p1;
		uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t) p0;
		c0 = n0;
		c1 = n1;
		c2 = n2;
		c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0;
	out[1] = c1;
	out[2] = c2;
	out[3] = c3;
}

static double u32_to_unit(uint32_t x) {
	return ((double) x + 0.5) * (1.0 / 4294967296.0);
}

/* ------------------------------------------------------------------ neighbour queries of the cluster moves */
/* what AVB's _set_neighbours (avb.c:43-75) collects for particle i: the neighbour bonded through each patch (MC_interact's onp,
 * MC.c:249-261), walking the 27 cells in the reference's order; returns the number of bonded neighbours.  mode 1: the particles
 * with dist2 < kf_sqr_rcut (vmmc.c:178) into list[max_nb] instead. */
int ppo_neighbour_query(ppo_system *s, int i, int mode, int max_nb, int *list) {
	int idx[3], j, k, l, n = 0, a;
	const double *rp = s->r + 3 * i, *pp = s->patches + 12 * i;
	for(a = 0; a < (mode == 0 ? 4 : max_nb); a++) list[a] = -1;
	ppo_cell_coords(rp, s->box, s->n_side, idx);
	for(j = -1; j < 2; j++) {
		int cx = (idx[0] + j + s->n_side[0]) % s->n_side[0];
		for(k = -1; k < 2; k++) {
			int cy = (idx[1] + k + s->n_side[1]) % s->n_side[1];
			for(l = -1; l < 2; l++) {
				int cz = (idx[2] + l + s->n_side[2]) % s->n_side[2];
				int c = (cx * s->n_side[1] + cy) * s->n_side[2] + cz;
				for(a = s->cell_start[c]; a < s->cell_start[c + 1]; a++) {
					const int q = s->cell_items[a];
					if(q == i) continue;
					const double *rq = s->r + 3 * q, *pq = s->patches + 12 * q;
					double d[3] = {rq[0] - rp[0], rq[1] - rp[1], rq[2] - rp[2]};
					int m, u, v, onp = -1;
					for(m = 0; m < 3; m++) d[m] -= s->box[m] * rint(d[m] / s->box[m]);
					const double d2 = dot3(d, d);
					if(mode == 1) {
						if(d2 < s->sqr_rcut) {
							if(n < max_nb) list[n] = q;
							n++;
						}
						continue;
					}
					if(d2 < 1. || d2 >= s->sqr_rcut) continue;
					const double norm = sqrt(d2);
					d[0] /= norm;
					d[1] /= norm;
					d[2] /= norm;
					for(u = 0; u < 4 && onp < 0; u++)
						if(dot3(d, pp + 3 * u) > s->kf_cosmax)
							for(v = 0; v < 4; v++)
								if(-dot3(d, pq + 3 * v) > s->kf_cosmax) {
									onp = u;
									break;
								}
					if(onp >= 0) {
						list[onp] = q; /* avb.c:64: the last one in the walk keeps the patch */
						n++;
					}
				}
			}
		}
	}
	return n;
}

/* ------------------------------------------------------------------ grand-canonical insertion / removal */
/* random_vector_on_sphere (utils.c:115-131) from a recorded stream */
static void gc_sphere(const double *u, int n_u, int *used, double *res) {
	double ran1, ran2, ransq;
	do {
		ran1 = 1. - 2. * (*used < n_u ? u[(*used)++] : 0.75);
		ran2 = 1. - 2. * (*used < n_u ? u[(*used)++] : 0.75);
		ransq = ran1 * ran1 + ran2 * ran2;
	} while(ransq >= 1. && *used < n_u);
	const double ranh = 2. * sqrt(1. - ransq);
	res[0] = ran1 * ranh;
	res[1] = ran2 * ranh;
	res[2] = 1. - 2. * ransq;
}

/* MC_add_remove (MC.c:342-448) with the uniforms the reference would draw, in its order.  Returns the kind of move (2 insertion,
 * 3 removal, -1 returned before a try); *accepted, *used (uniforms consumed); *energy_bonds is the running bond count.  A new
 * particle gets index N; when particle p leaves, the last one takes its index (MC.c:414-438). */
int ppo_add_remove(ppo_system *s, double z, int n_min, int n_max, const double *u, int n_u, int *accepted, int *used, long long *energy_bonds) {
	int k, ov;
	*accepted = 0;
	*used = 0;
	if(u[(*used)++] < 0.5) { /* MC.c:344 */
		double r[3], x[3], y[3], zz[3], m[9], pp[12];
		if(s->N == n_max) return -1;
		for(k = 0; k < 3; k++) r[k] = u[(*used)++] * s->box[k]; /* MC.c:352-354 */
		/* random_orientation (utils.c:133-169): three points, Gram-Schmidt, rows (z, x, y), first two swapped if det < 0 */
		gc_sphere(u, n_u, used, x);
		gc_sphere(u, n_u, used, y);
		gc_sphere(u, n_u, used, zz);
		{
			double g[9];
			memcpy(g, x, sizeof(x));
			memcpy(g + 3, y, sizeof(y));
			memcpy(g + 6, zz, sizeof(zz));
			ppo_gram_schmidt(g);
			memcpy(x, g, sizeof(x));
			memcpy(y, g + 3, sizeof(y));
			memcpy(zz, g + 6, sizeof(zz));
		}
		memcpy(m, zz, sizeof(zz));
		memcpy(m + 3, x, sizeof(x));
		memcpy(m + 6, y, sizeof(y));
		{
			double det = 0.;
			det += m[0] * m[4] * m[8];
			det += m[1] * m[5] * m[6];
			det += m[2] * m[3] * m[7];
			det -= m[2] * m[4] * m[6];
			det -= m[0] * m[5] * m[7];
			det -= m[1] * m[3] * m[8];
			if(det < 0) {
				memcpy(m, x, sizeof(x));
				memcpy(m + 3, zz, sizeof(zz));
			}
		}
		patches_from(s->base, m, pp);
		const int bonds = probe_bonds(s, -1, r, pp, &ov, NULL);          /* MC.c:361 */
		const double delta_E = -(double) bonds;
		const double acc = exp(-delta_E / s->T) * z * s->V / (s->N + 1.); /* MC.c:362 */
		if(!ov && u[(*used)++] < acc) {                                   /* MC.c:364 */
			const size_t n = (size_t) s->N + 1;
			s->r = realloc(s->r, n * 3 * sizeof(double));
			s->m = realloc(s->m, n * 9 * sizeof(double));
			s->patches = realloc(s->patches, n * 12 * sizeof(double));
			memcpy(s->r + 3 * s->N, r, sizeof(r));
			memcpy(s->m + 9 * s->N, m, sizeof(m));
			memcpy(s->patches + 12 * s->N, pp, sizeof(pp));
			s->N++;
			*energy_bonds += bonds;
			cells_build(s);
			*accepted = 1;
		}
		return 2;
	}
	if(s->N == n_min) return -1;
	{
		const int p = (int) (u[(*used)++] * s->N);                       /* MC.c:383 */
		const int bonds = ppo_particle_bonds(s, p, &ov);
		const double delta_E = (double) bonds;
		const double acc = exp(-delta_E / s->T) * s->N / (s->V * z);     /* MC.c:386 */
		if(u[(*used)++] < acc) {
			const int q = s->N - 1;
			if(p != q) {
				memcpy(s->r + 3 * p, s->r + 3 * q, 3 * sizeof(double));
				memcpy(s->m + 9 * p, s->m + 9 * q, 9 * sizeof(double));
				patches_from(s->base, s->m + 9 * p, s->patches + 12 * p);   /* MC.c:428-430 */
			}
			s->N--;
			*energy_bonds -= bonds;
			cells_build(s);
			*accepted = 1;
		}
		return 3;
	}
}

void ppo_even_grid(const double *box, double r_cut, int *n_side) {
	int k;
	for(k = 0; k < 3; k++) {
		int n = (int) floor(box[k] / r_cut);
		n &= ~1;
		if(n >= 2 && box[k] / n < r_cut * (1. + 1e-9)) n -= 2;
		n_side[k] = n; /* < 2 means the box is too small for the checkerboard */
	}
}

#define PPO_TAG_SHIFT 0xFFFFFFFFu

void ppo_sweep_shift(uint64_t seed, uint64_t sweep, const double *box, const int *n_side, double *shift) {
	uint32_t x[4];
	int k;
	ppo_philox4x32(PPO_TAG_SHIFT, 0u, (uint32_t) sweep, (uint32_t) ((sweep >> 32) << 8), (uint32_t) seed, (uint32_t) (seed >> 32), x);
	for(k = 0; k < 3; k++) shift[k] = u32_to_unit(x[k]) * (box[k] / n_side[k]);
}

/* cell of a position on the sweep's own grid (the engine's cell_coord, pp_common.cuh) */
static int grid_cell(const double *r, const double *shift, const double *inv_box, const int *n_side, int *idx) {
	int k;
	for(k = 0; k < 3; k++) {
		double t = (r[k] - shift[k]) * inv_box[k];
		double f = t - floor(t);
		int c = (int) (f * (double) n_side[k]);
		idx[k] = c < n_side[k] ? c : n_side[k] - 1;
	}
	return (idx[0] * n_side[1] + idx[1]) * n_side[2] + idx[2];
}

void ppo_sweep(ppo_system *s, uint64_t seed, uint64_t sweep, int shift_every, double disp_max, double theta_max, long long *energy_bonds,
		ppo_sweep_stats *st) {
	int n_side[3], k, i;
	double shift[3], inv_box[3], inv_w[3];
	ppo_even_grid(s->box, s->r_cut, n_side);
	if(n_side[0] < 2 || n_side[1] < 2 || n_side[2] < 2) {
		fprintf(stderr, "ppo_sweep: box too small for the checkerboard\n");
		abort();
	}
	/* the shift is redrawn every shift_every sweeps; in between the cell membership cannot change
	 * (moves are confined), so re-binning every sweep with the period's shift is equivalent */
	ppo_sweep_shift(seed, sweep / (uint64_t) shift_every, s->box, n_side, shift);
	for(k = 0; k < 3; k++) {
		inv_box[k] = 1. / s->box[k];
		inv_w[k] = (double) n_side[k] * inv_box[k];
	}
	int n_cells = n_side[0] * n_side[1] * n_side[2];

	/* counting sort by shifted cell; ascending particle id inside a cell */
	int *start = calloc((size_t) n_cells + 1, sizeof(int));
	int *items = malloc(sizeof(int) * (size_t) s->N);
	int *cell_of = malloc(sizeof(int) * (size_t) s->N);
	int idx[3];
	for(i = 0; i < s->N; i++) {
		cell_of[i] = grid_cell(s->r + 3 * i, shift, inv_box, n_side, idx);
		start[cell_of[i] + 1]++;
	}
	for(i = 0; i < n_cells; i++) start[i + 1] += start[i];
	int *fill = calloc((size_t) n_cells, sizeof(int));
	for(i = 0; i < s->N; i++) items[start[cell_of[i]] + fill[cell_of[i]]++] = i;
	free(fill);

	double exp_table[32];
	for(k = 0; k < 32; k++) exp_table[k] = exp(-(double) k / s->T);

	uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
	uint32_t c2 = (uint32_t) sweep, c3hi = (uint32_t) ((sweep >> 32) << 8);

	int colour;
	for(colour = 0; colour < 8; colour++) {
		/* pass order: x parity slowest, so that a slab decomposition along x exchanges halos twice per sweep */
		int ox = (colour >> 2) & 1, oy = (colour >> 1) & 1, oz = colour & 1;
		int ix, iy, iz;
		for(ix = ox; ix < n_side[0]; ix += 2)
			for(iy = oy; iy < n_side[1]; iy += 2)
				for(iz = oz; iz < n_side[2]; iz += 2) {
					int cell = (ix * n_side[1] + iy) * n_side[2] + iz;
					int n_here = start[cell + 1] - start[cell];
					if(n_here == 0) continue;

					/* deduplicated neighbourhood (an axis with 2 cells has only offsets 0 and +1) */
					int nb[27], n_nb = 0, dx, dy, dz;
					for(dx = -1; dx < 2; dx++) {
						if(n_side[0] == 2 && dx == -1) continue;
						for(dy = -1; dy < 2; dy++) {
							if(n_side[1] == 2 && dy == -1) continue;
							for(dz = -1; dz < 2; dz++) {
								if(n_side[2] == 2 && dz == -1) continue;
								int cx = (ix + dx + n_side[0]) % n_side[0];
								int cy = (iy + dy + n_side[1]) % n_side[1];
								int cz = (iz + dz + n_side[2]) % n_side[2];
								nb[n_nb++] = (cx * n_side[1] + cy) * n_side[2] + cz;
							}
						}
					}

					int t;
					for(t = 0; t < n_here; t++) {
						uint32_t x[4], y[4], z[4];
						ppo_philox4x32((uint32_t) cell, (uint32_t) t, c2, c3hi | 0u, k0, k1, x);
						ppo_philox4x32((uint32_t) cell, (uint32_t) t, c2, c3hi | 1u, k0, k1, y);
						int pick = (int) (((uint64_t) x[0] * (uint64_t) n_here) >> 32);
						int p = items[start[cell] + pick];
						double disp[3];
						for(k = 0; k < 3; k++) disp[k] = (u32_to_unit(x[1 + k]) - 0.5) * disp_max;
						double theta = u32_to_unit(y[0]) * theta_max;
						double u_acc = (double) ((((uint64_t) y[1]) << 21) | (((uint64_t) y[2]) >> 11)) * (1.0 / 9007199254740992.0);
						double ran1 = 0., ran2 = 0.;
						uint32_t block = 2;
						for(;;) {
							ppo_philox4x32((uint32_t) cell, (uint32_t) t, c2, c3hi | block, k0, k1, z);
							ran1 = 1. - 2. * u32_to_unit(z[0]);
							ran2 = 1. - 2. * u32_to_unit(z[1]);
							if(ran1 * ran1 + ran2 * ran2 < 1.) break;
							ran1 = 1. - 2. * u32_to_unit(z[2]);
							ran2 = 1. - 2. * u32_to_unit(z[3]);
							if(ran1 * ran1 + ran2 * ran2 < 1.) break;
							block++;
						}
						st->tries++;

						double r_new[3] = {s->r[3 * p] + disp[0], s->r[3 * p + 1] + disp[1], s->r[3 * p + 2] + disp[2]};
						if(grid_cell(r_new, shift, inv_box, n_side, idx) != cell) {
							st->confined++;
							continue;
						}
						double keep = (1. - fabs(disp[0]) * inv_w[0]) * (1. - fabs(disp[1]) * inv_w[1]) * (1. - fabs(disp[2]) * inv_w[2]);
						if(keep < 0.0078125) keep = 0.0078125; /* Disp_max comparable to the cell: the 2^-24 fixed point must not wrap */
						double W = 1. / keep;
						uint64_t Wfix = (uint64_t) (W * 16777216.0);
						st->w_tries += Wfix;

						double axis[3], m_new[9], p_new[12];
						ppo_axis_from_pair(ran1, ran2, axis);
						ppo_rotate(s->m + 9 * p, axis, theta, m_new);
						patches_from(s->base, m_new, p_new);

						int b_old = 0, b_new = 0, ov_new = 0, a, c;
						for(c = 0; c < n_nb; c++) {
							for(a = start[nb[c]]; a < start[nb[c] + 1]; a++) {
								int q = items[a];
								if(q == p) continue;
								st->pair_checks += 2;
								int v_old = ppo_pair(s->r + 3 * p, s->patches + 12 * p, s->r + 3 * q, s->patches + 12 * q, s->box, s->sqr_rcut, s->kf_cosmax);
								int v_new = ppo_pair(r_new, p_new, s->r + 3 * q, s->patches + 12 * q, s->box, s->sqr_rcut, s->kf_cosmax);
								if(v_old == 1) b_old++;
								if(v_new == 1) b_new++;
								if(v_new == -1) ov_new = 1;
							}
						}
						if(ov_new) {
							st->overlapped++;
							continue;
						}
						int dE = b_old - b_new; /* energy = -bonds */
						if(dE < 0 || u_acc < exp_table[dE < 32 ? dE : 31]) {
							memcpy(s->r + 3 * p, r_new, sizeof(r_new));
							memcpy(s->m + 9 * p, m_new, sizeof(m_new));
							memcpy(s->patches + 12 * p, p_new, sizeof(p_new));
							*energy_bonds += b_new - b_old;
							st->accepted++;
							st->w_accepted += Wfix;
						}
					}
				}
	}
	free(start);
	free(items);
	free(cell_of);
	s->cells_valid = 0;
	cells_build(s); /* keep the reference-grid lists consistent with the new positions */
}
