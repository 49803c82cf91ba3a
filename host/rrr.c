#include "rrr.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double dot(const double *a, const double *b) {
	return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

static void unit(double *v) {
	double n = sqrt(dot(v, v));
	v[0] /= n;
	v[1] /= n;
	v[2] /= n;
}

/* rows 0/1 of the file -> orthonormal 3x3 (system.c:225-233, utils.c:91-113) */
static void build_orientation(const double *a, const double *b, double *m) {
	double *v1 = m, *v2 = m + 3, *v3 = m + 6;
	memcpy(v1, a, 3 * sizeof(double));
	memcpy(v2, b, 3 * sizeof(double));
	unit(v1);
	unit(v2);
	v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
	v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
	v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
	double n1 = dot(v1, v1);
	double c = dot(v2, v1) / n1;
	for(int k = 0; k < 3; k++) v2[k] -= c * v1[k];
	double c31 = dot(v3, v1), c32 = dot(v3, v2), n2 = dot(v2, v2);
	double f1 = c31 / n1, f2 = c32 / n2;
	for(int k = 0; k < 3; k++) v3[k] -= f1 * v1[k] + f2 * v2[k];
	unit(v1);
	unit(v2);
	unit(v3);
}

static const char kRrbMagic[8] = {'P', 'P', 'R', 'R', 'B', '1', 0, 0};

static int rrb_read(FILE *f, long long *step, int *N, double box[3], double **r_out, double **m_out, int *found) {
	long long head[2];
	if(fread(head, sizeof(long long), 2, f) != 2 || fread(box, sizeof(double), 3, f) != 3 || head[1] < 0 || head[1] > 2147483647LL) return -2;
	const size_t n = (size_t) head[1], cap = n > 0 ? n : 1;
	double *r = malloc(cap * 3 * sizeof(double)), *m = malloc(cap * 9 * sizeof(double));
	if(r == NULL || m == NULL) {
		free(r);
		free(m);
		return -2;
	}
	const size_t got_r = fread(r, 3 * sizeof(double), n, f), got_m = fread(m, 9 * sizeof(double), n, f);
	if(found) *found = (int) (got_m < got_r ? got_m : got_r);
	if(got_r != n || got_m != n) {
		free(r);
		free(m);
		return -3;
	}
	if(step) *step = head[0];
	*N = (int) n;
	*r_out = r;
	*m_out = m;
	return 0;
}

int pph_rrb_write(const char *path, long long step, int N, const double box[3], const double *r, const double *m) {
	FILE *f = fopen(path, "wb");
	if(f == NULL) return -1;
	const long long head[2] = {step, N};
	int ok = fwrite(kRrbMagic, 1, 8, f) == 8 && fwrite(head, sizeof(long long), 2, f) == 2 && fwrite(box, sizeof(double), 3, f) == 3;
	ok = ok && fwrite(r, 3 * sizeof(double), (size_t) N, f) == (size_t) N && fwrite(m, 9 * sizeof(double), (size_t) N, f) == (size_t) N;
	return (fclose(f) == 0 && ok) ? 0 : -1;
}

int pph_rrr_read(const char *path, long long *step, int *N, double box[3], double **r_out, double **m_out, int *found) {
	FILE *f = fopen(path, "r");
	if(f == NULL) return -1;
	{ /* binary side format? */
		char magic[8];
		if(fread(magic, 1, 8, f) == 8 && memcmp(magic, kRrbMagic, 8) == 0) {
			const int rc = rrb_read(f, step, N, box, r_out, m_out, found);
			fclose(f);
			return rc;
		}
		rewind(f);
	}
	long long st;
	int n;
	if(fscanf(f, "%lld %d %lf %lf %lf", &st, &n, box, box + 1, box + 2) != 5 || n < 0) {
		fclose(f);
		return -2;
	}
	size_t cap = (size_t) (n > 0 ? n : 1);
	double *r = malloc(cap * 3 * sizeof(double)), *m = malloc(cap * 9 * sizeof(double));
	int i = 0;
	for(;;) {
		double a[3], b[3], p[3];
		int got = fscanf(f, "%lf %lf %lf %lf %lf %lf %lf %lf %lf", a, a + 1, a + 2, b, b + 1, b + 2, p, p + 1, p + 2);
		if(got != 9) break;
		if(i < n) {
			memcpy(r + 3 * (size_t) i, p, sizeof(p));
			build_orientation(a, b, m + 9 * (size_t) i);
		}
		i++;
	}
	fclose(f);
	if(found) *found = i;
	if(i != n) {
		free(r);
		free(m);
		return -3;
	}
	if(step) *step = st;
	*N = n;
	*r_out = r;
	*m_out = m;
	return 0;
}

int pph_rrr_write(const char *path, long long step, int N, const double box[3], const double *r, const double *m) {
	FILE *f = fopen(path, "w");
	if(f == NULL) return -1;
	fprintf(f, "%lld %d %.12lf %.12lf %.12lf\n", step, N, box[0], box[1], box[2]);
	for(int i = 0; i < N; i++) {
		const double *o = m + 9 * (size_t) i, *p = r + 3 * (size_t) i;
		fprintf(f, "%lf %lf %lf\n", o[0], o[1], o[2]);
		fprintf(f, "%lf %lf %lf\n", o[3], o[4], o[5]);
		fprintf(f, "%.12lf %.12lf %.12lf\n", p[0], p[1], p[2]);
	}
	return fclose(f) == 0 ? 0 : -1;
}
