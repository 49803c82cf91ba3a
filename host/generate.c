/* generate.c -- patchy_generate: initial configurations at scale.
 *
 * The reference's generator (generator.c:43-84) inserts particles one by one at random positions and
 * retries on overlap: serial, minutes at N = 4M, and it refuses densities above 0.7
 * (generator.c:99-102).  This program builds the same kind of state -- N hard spheres of diameter 1 at
 * number density rho in a cubic box of side cbrt(N / rho), no overlap, random orientations -- in two steps:
 *   1. a simple-cubic lattice of spacing a >= 1 filling the box, randomly thinned to N sites, every
 *      particle displaced inside its site by less than (a - 1) / 2 per axis: overlap-free by construction,
 *      any density below 1, O(N);
 *   2. optionally, `melt` sweeps of the GPU engine at a temperature where bonds do not matter, i.e.
 *      hard-sphere Monte Carlo, which turns the thinned lattice into a fluid (lattice melt).
 * Output: the reference's text format (generator.c:73-81 -- header "0 N Lx Ly Lz", then per particle
 * the first two orientation rows and the position), or the binary side format of rrr.h when the
 * output name ends in ".rrb".
 *
 * usage: patchy_generate N density [seed [output [melt_sweeps [device]]]]
 */
#include "../include/pp_b200.h"
#include "rrr.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* splitmix64: small, seedable, good enough to place particles */
static uint64_t rng_state;
static uint64_t rng_next(void) {
	uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
static double rng_unit(void) { return ((double) (rng_next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); } /* (0, 1) */
static uint64_t rng_below(uint64_t n) { return (uint64_t) (rng_unit() * (double) n) % n; }

/* uniform random rotation: unit quaternion from four normal deviates */
static void random_orientation(double *m) {
	double q[4], norm = 0.;
	for(int k = 0; k < 4; k += 2) {
		const double u = rng_unit(), v = rng_unit(), rad = sqrt(-2. * log(u));
		q[k] = rad * cos(2. * M_PI * v);
		q[k + 1] = rad * sin(2. * M_PI * v);
	}
	for(int k = 0; k < 4; k++) norm += q[k] * q[k];
	norm = sqrt(norm);
	const double w = q[0] / norm, x = q[1] / norm, y = q[2] / norm, z = q[3] / norm;
	m[0] = 1 - 2 * (y * y + z * z);
	m[1] = 2 * (x * y - z * w);
	m[2] = 2 * (x * z + y * w);
	m[3] = 2 * (x * y + z * w);
	m[4] = 1 - 2 * (x * x + z * z);
	m[5] = 2 * (y * z - x * w);
	m[6] = 2 * (x * z - y * w);
	m[7] = 2 * (y * z + x * w);
	m[8] = 1 - 2 * (x * x + y * y);
}

static int ends_with(const char *s, const char *tail) {
	const size_t a = strlen(s), b = strlen(tail);
	return a >= b && strcmp(s + a - b, tail) == 0;
}

int main(int argc, char *argv[]) {
	if(argc < 3) {
		fprintf(stderr, "Usage is %s N density [seed [output [melt_sweeps [device]]]]\n", argv[0]);
		return 1;
	}
	const long long N = atoll(argv[1]);
	const double density = atof(argv[2]);
	const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 10) : 1;
	const char *name = argc > 4 ? argv[4] : "generated.rrr";
	const long long melt = argc > 5 ? atoll(argv[5]) : 0;
	const int device = argc > 6 ? atoi(argv[6]) : 0;
	if(N < 1 || N > 2000000000LL || !(density > 0.)) {
		fprintf(stderr, "N must be in [1, 2e9] and the density positive\n");
		return 1;
	}
	const double L = cbrt((double) N / density);
	long long n_side = (long long) floor(L / 1.0000001);
	if(n_side < 1 || n_side * n_side * n_side < N) {
		fprintf(stderr, "The lattice-thinning generator cannot produce configurations with density %g: a simple-cubic lattice of spacing 1 holds %lld particles in a box of side %g\n",
				density, n_side > 0 ? n_side * n_side * n_side : 0LL, L);
		return 1;
	}
	const double a = L / (double) n_side, jitter = 0.5 * (a - 1.) * 0.999;
	const long long sites = n_side * n_side * n_side;
	if(sites > 2147483647LL) {
		fprintf(stderr, "box too large for this generator (%lld lattice sites)\n", sites);
		return 1;
	}
	rng_state = seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull;

	double *r = malloc(sizeof(double) * 3 * (size_t) N), *m = malloc(sizeof(double) * 9 * (size_t) N);
	int *site = malloc(sizeof(int) * (size_t) sites);
	if(r == NULL || m == NULL || site == NULL) {
		fprintf(stderr, "out of memory\n");
		return 1;
	}
	/* N distinct sites: partial Fisher-Yates shuffle */
	for(long long i = 0; i < sites; i++) site[i] = (int) i;
	for(long long i = 0; i < N; i++) {
		const long long j = i + (long long) rng_below((uint64_t) (sites - i));
		const int t = site[i];
		site[i] = site[j];
		site[j] = t;
	}
	for(long long i = 0; i < N; i++) {
		const long long s = site[i], ix = s / (n_side * n_side), iy = (s / n_side) % n_side, iz = s % n_side;
		r[3 * i] = ((double) ix + 0.5) * a + (2. * rng_unit() - 1.) * jitter;
		r[3 * i + 1] = ((double) iy + 0.5) * a + (2. * rng_unit() - 1.) * jitter;
		r[3 * i + 2] = ((double) iz + 0.5) * a + (2. * rng_unit() - 1.) * jitter;
		random_orientation(m + 9 * i);
	}
	free(site);
	const double box[3] = {L, L, L};

	if(melt > 0) { /* hard-sphere Monte Carlo on the GPU: bonds do not matter at this temperature */
		pp_params prm;
		pp_default_params(&prm);
		prm.ensemble = PP_NVT;
		prm.temperature = 1e9;
		prm.disp_max = 0.3;
		prm.theta_max = 0.5;
		prm.seed = seed;
		prm.device = device;
		pp_handle *h = NULL;
		if(pp_create(&prm, &h) || h == NULL) {
			fprintf(stderr, "pp_create failed: %s\n", h ? pp_last_error(h) : "no handle");
			return 1;
		}
		long long bonds;
		int overlap = 0, ok = 0;
		if(pp_upload(h, (int) N, box, r, m) || pp_sweep(h, melt) || pp_check_energy(h, &bonds, &overlap, &ok) || pp_download(h, r, m)) {
			fprintf(stderr, "melt failed: %s\n", pp_last_error(h));
			return 1;
		}
		if(overlap || !ok) {
			fprintf(stderr, "melt produced an inconsistent configuration\n");
			return 1;
		}
		pp_observables o;
		pp_get_observables(h, &o, 0);
		fprintf(stderr, "Melted for %lld sweeps: %lld of %lld attempts accepted\n", melt, o.accepted[PP_ROTO_TRASL], o.tries[PP_ROTO_TRASL]);
		pp_destroy(h);
	}

	const int rc = ends_with(name, ".rrb") ? pph_rrb_write(name, 0, (int) N, box, r, m) : pph_rrr_write(name, 0, (int) N, box, r, m);
	if(rc != 0) {
		fprintf(stderr, "File '%s' is not writable\n", name);
		return 1;
	}
	fprintf(stderr, "Generation done. The new configuration has been written to the file '%s'\n", name);
	free(r);
	free(m);
	return 0;
}
