/*
 * patchy_b200 -- C host of the B200 Monte Carlo engine.
 *
 * Keeps the reference program's contract for the roto-translation path (Dynamics = 0, Ensemble = 0
 * or 5): one positional argument = input file with the reference's keys, .rrr configurations in and
 * out, energy.dat / acceptance.dat / density.dat / boxshape.dat and the stdout line in the
 * reference's formats, Print_every / Save_every / Restart_step_counter semantics and graceful stop on
 * SIGTERM / SIGABRT / SIGINT / SIGUSR2 (main.c:26-143, output.c:19-211, system.c:58-252).
 * Every number comes from libpatchy_b200.so (include/pp_b200.h); there is no CPU fallback.
 *
 * Extra optional keys: GPU_device, Lanes_per_cell, Shift_every, Lx_reference_sign.
 */
#include "../include/pp_b200.h"
#include "input.h"
#include "rrr.h"

#include <signal.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef long long llint;

static volatile sig_atomic_t stop_requested = 0;

static void on_signal(int sig) {
	/* the next time the default handler runs, so the program stays killable (main.c:15-24) */
	signal(sig, SIG_DFL);
	static const char msg[] = "# Caught SIGNAL; setting stop = 1\n";
	if(write(2, msg, sizeof(msg) - 1) < 0) {}
	stop_requested = 1;
}

typedef struct {
	FILE *log, *energy, *density, *acc, *boxshape;
	llint start_from, print_every, save_every;
	char conf_folder[512], conf_last[512];
	int print_bonds; /* Print_bonds / Bonds_folder (output.c:88-99) */
	char bonds_folder[512];
	int *bond_counts, *bond_neighbours;
} host_output;

#define PP_HOST_GC 1  /* Ensemble values the host itself drives (defs.h:22-23) */
#define PP_HOST_SUS 3
#define PP_HOST_MAX_BONDS 16 /* neighbours listed per particle; four patches of half-angle acos(0.98) hold far fewer */

static void log_msg(host_output *out, const char *fmt, ...) {
	va_list ap;
	if(out->log != NULL && out->log != stderr) {
		va_start(ap, fmt);
		vfprintf(out->log, fmt, ap);
		va_end(ap);
		fflush(out->log);
	}
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
}

/* output_exit (output.c:374-388): message to the log and stderr, exit status 1 */
static void die(host_output *out, const char *fmt, ...) {
	va_list ap;
	if(out != NULL && out->log != NULL && out->log != stderr) {
		va_start(ap, fmt);
		vfprintf(out->log, fmt, ap);
		va_end(ap);
		fflush(out->log);
	}
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
	exit(1);
}

#define PP_CALL(out, h, call) \
	do { \
		if((call) != 0) die(out, "%s\n", pp_group_last_error(h)); \
	} while(0)

static FILE *open_or_die(host_output *out, const char *name, const char *mode, const char *what) {
	FILE *f = fopen(name, mode);
	if(f == NULL) die(out, "%s file '%s' is not writable\n", what, name);
	return f;
}

static void save_configuration(host_output *out, pp_group *h, int N, llint step, const char *name, double *r, double *orient) {
	pp_observables obs;
	PP_CALL(out, h, pp_group_download(h, r, orient));
	PP_CALL(out, h, pp_group_get_observables(h, &obs, 0));
	if(pph_rrr_write(name, step, N, obs.box, r, orient) != 0) die(out, "File '%s' is not writable\n", name);
}

/* output_print (output.c:119-211) */
/* output_print_bonds (output.c:213-268): N, then two lines per particle -- "index n_bonds" and the indices of the bonded
 * neighbours, each followed by a blank (possibly an empty line).  The reference builds the second line with an
 * overlapping sprintf (output.c:250, undefined behaviour); here it is simply the list, cell by cell. */
static void print_bonds(host_output *out, pp_group *h, int N, llint step) {
	char name[1100];
	snprintf(name, sizeof(name), "%s/bonds_%lld.dat", out->bonds_folder, step);
	if(pp_group_bond_lists(h, PP_HOST_MAX_BONDS, out->bond_counts, out->bond_neighbours)) {
		fprintf(stderr, "%s\n", pp_group_last_error(h));
		exit(1);
	}
	FILE *f = fopen(name, "w");
	if(f == NULL) {
		fprintf(stderr, "Bonds file '%s' is not writable\n", name);
		exit(1);
	}
	fprintf(f, "%d\n", N);
	for(int i = 0; i < N; i++) {
		fprintf(f, "%d %d\n", i, out->bond_counts[i]);
		const int n = out->bond_counts[i] < PP_HOST_MAX_BONDS ? out->bond_counts[i] : PP_HOST_MAX_BONDS;
		for(int k = 0; k < n; k++) fprintf(f, "%d ", out->bond_neighbours[(size_t) i * PP_HOST_MAX_BONDS + k]);
		fprintf(f, "\n");
	}
	fclose(f);
}

static void print_observables(host_output *out, pp_group *h, const pp_params *prm, int ensemble, llint step) {
	pp_observables o;
	PP_CALL(out, h, pp_group_get_observables(h, &o, 1)); /* counters are reset after each print (output.c:189) */
	const double V = o.volume; /* System.V, output.c:131 */
	const double E = (o.N > 0) ? o.energy / o.N : 0.;
	fprintf(out->energy, "%lld %lf\n", step, E);
	fflush(out->energy);
	if(ensemble != PP_NVT) {
		fprintf(out->density, "%lld %lf %d\n", step, o.N / V, o.N);
		fflush(out->density);
	}
	if(prm->Lx_move) {
		fprintf(out->boxshape, "%lld %lf %lf %lf\n", step, o.box[0], o.box[1], o.box[2]);
		fflush(out->boxshape);
	}
	/* accepted / tries per move type; the roto-translation column is the engine's estimator of the same
	 * ratio (cell-confined proposals are geometry, not physics: DESIGN.md "acceptance estimator") */
	const double zero = 0.;
	const double acc_roto = (o.tries[PP_ROTO_TRASL] > 0) ? o.acceptance_roto : zero / zero;
	fprintf(out->acc, "%lld %e", step, acc_roto);
	if(ensemble == PP_HOST_GC || ensemble == PP_HOST_SUS) { /* output.c:166-170 */
		fprintf(out->acc, " %e", o.accepted[PP_ADD] / (double) o.tries[PP_ADD]);
		fprintf(out->acc, " %e", o.accepted[PP_REMOVE] / (double) o.tries[PP_REMOVE]);
	}
	if(prm->ensemble == PP_NPT) fprintf(out->acc, " %e", o.accepted[PP_VOLUME] / (double) o.tries[PP_VOLUME]);
	if(prm->Lx_move) fprintf(out->acc, " %e", o.accepted[PP_LX] / (double) o.tries[PP_LX]);
	fprintf(out->acc, "\n");
	fflush(out->acc);
	printf("%lld %lf %lf", step, E, o.energy);
	if(ensemble != PP_NVT) printf(" %lf", o.N / V);
	printf("\n");
	fflush(stdout);
}

int main(int argc, char *argv[]) {
	host_output out;
	memset(&out, 0, sizeof(out));
	out.log = stderr;
	if(argc < 2) die(&out, "Usage is %s input\n", argv[0]);

	signal(SIGTERM, on_signal);
	signal(SIGABRT, on_signal);
	signal(SIGINT, on_signal);
	signal(SIGUSR2, on_signal);

	pph_input *in = pph_input_load(argv[1]);
	if(in == NULL) {
		fprintf(stderr, "Input file not found\n");
		return 1;
	}

	/* ---- output_init (output.c:19-117) ----------------------------------------------------- */
	int restart_step_counter = 1;
	pph_get_int(in, "Restart_step_counter", &restart_step_counter, 0);
	const char *mode = restart_step_counter ? "w" : "a";
	char name[1100], conf_name[512];
	if(pph_get_string(in, "Log_file", name, sizeof(name), 0) && strcmp(name, "none") != 0) {
		out.log = fopen(name, "a");
		if(out.log == NULL) {
			out.log = stderr;
			die(&out, "Log file '%s' is not writable\n", name);
		}
	}
	pph_get_string(in, "Initial_conditions_file", conf_name, sizeof(conf_name), 1);
	pph_get_llint(in, "Print_every", &out.print_every, 1);
	pph_get_llint(in, "Save_every", &out.save_every, 1);
	if(out.print_every < 1 || out.save_every < 1) die(&out, "Print_every and Save_every must be positive\n");
	snprintf(name, sizeof(name), "energy.dat");
	pph_get_string(in, "Energy_file", name, sizeof(name), 0);
	out.energy = open_or_die(&out, name, mode, "Energy");
	snprintf(name, sizeof(name), "acceptance.dat");
	pph_get_string(in, "Acceptance_file", name, sizeof(name), 0);
	out.acc = open_or_die(&out, name, mode, "Acceptance");
	pph_get_string(in, "Configuration_folder", out.conf_folder, sizeof(out.conf_folder), 1);
	if(access(out.conf_folder, W_OK) != 0)
		die(&out, "Cannot create files in directory '%s': please make sure that the directory exists and it is writable\n", out.conf_folder);
	snprintf(out.conf_last, sizeof(out.conf_last), "last.rrr");
	pph_get_string(in, "Configuration_last", out.conf_last, sizeof(out.conf_last), 0);

	/* ---- system_init (system.c:58-252) ----------------------------------------------------- */
	pp_params prm;
	pp_default_params(&prm);
	int dynamics = 0, seed = 0, itmp = 0;
	pph_get_int(in, "Dynamics", &dynamics, 1);
	pph_get_int(in, "Ensemble", &prm.ensemble, 1);
	pph_get_double(in, "Disp_max", &prm.disp_max, 1);
	pph_get_double(in, "Theta_max", &prm.theta_max, 1);
	pph_get_double(in, "Temperature", &prm.temperature, 1);
	if(!pph_get_int(in, "Seed", &seed, 0)) {
		seed = (int) time(NULL);
		log_msg(&out, "Using seed %d\n", seed);
	}
	prm.seed = (uint64_t) (unsigned int) seed;
	pph_get_double(in, "KF_delta", &prm.kf_delta, 1);
	pph_get_double(in, "KF_cosmax", &prm.kf_cosmax, 1);
	log_msg(&out, "Patch parameters: cosmax = %lf, delta = %lf\n", prm.kf_cosmax, prm.kf_delta);
	if(dynamics != 0) die(&out, "Dynamics %d not supported on the GPU path (only 0, roto-translations; VMMC/AVB stay on the CPU reference)\n", dynamics);
	pph_get_int(in, "Print_bonds", &out.print_bonds, 0);
	if(out.print_bonds) { /* output.c:88-99 */
		if(prm.ensemble != PP_NVT) die(&out, "The printing of bonds is unavailable in non-canonical ensembles");
		pph_get_string(in, "Bonds_folder", out.bonds_folder, sizeof(out.bonds_folder), 1);
		if(access(out.bonds_folder, W_OK) != 0)
			die(&out, "Cannot create files in directory '%s': please make sure that the directory exists and it is writable\n", out.bonds_folder);
	}
	/* Ensemble = 1 (GC) / 3 (SUS), system.c:90-103: insertions and removals (MC_add_remove, MC.c:342-448) are driven from here,
	 * pp_gc_step before every sweep; the engine itself runs its NVT sweeps */
	const int ensemble = prm.ensemble;
	double activity = 0.;
	int n_min = 0, n_max = 0;
	llint *sus_hist = NULL;
	char sus_folder[512] = "";
	if(ensemble == PP_HOST_GC || ensemble == PP_HOST_SUS) {
		pph_get_double(in, "Activity", &activity, 1);
		if(ensemble == PP_HOST_GC) pph_get_int(in, "GC_N_max", &n_max, 1);
		else {
			pph_get_int(in, "Umbrella_sampling_min", &n_min, 1);
			pph_get_int(in, "Umbrella_sampling_max", &n_max, 1);
			pph_get_string(in, "Umbrella_sampling_folder", sus_folder, sizeof(sus_folder), 1);
			if(access(sus_folder, W_OK) != 0)
				die(&out, "Cannot create files in directory '%s': please make sure that the directory exists and it is writable\n", sus_folder);
		}
		prm.ensemble = PP_NVT;
	}
	else if(prm.ensemble != PP_NVT && prm.ensemble != PP_NPT)
		die(&out, "Ensemble %d not supported on the GPU path (0 = NVT, 1 = GC, 3 = SUS, 5 = NPT; BSUS stays on the CPU reference)\n", prm.ensemble);
	if(ensemble != PP_NVT) {
		snprintf(name, sizeof(name), "density.dat");
		pph_get_string(in, "Density_file", name, sizeof(name), 0);
		out.density = open_or_die(&out, name, mode, "Density");
	}

	llint conf_step = 0;
	int N = 0, found = 0;
	double box[3], *r = NULL, *orient = NULL;
	int rc = pph_rrr_read(conf_name, &conf_step, &N, box, &r, &orient, &found);
	if(rc == -1) die(&out, "Initial_conditions_file '%s' is not readable\n", conf_name);
	if(rc == -2) die(&out, "The initial configuration file '%s' is empty or its headers are malformed\n", conf_name);
	if(rc == -3) die(&out, "Number of particles found in configuration (%d) is different from the value found in the header (%d)\n", found, N);
	out.start_from = restart_step_counter ? 0 : conf_step;
	if(ensemble == PP_HOST_GC || ensemble == PP_HOST_SUS) {
		if(ensemble == PP_HOST_SUS) { /* system.c:100-102 */
			if(N < n_min) die(&out, "Number of particles %d is smaller than Umbrella_sampling_min (%d)\n", N, n_min);
			if(N > n_max) die(&out, "Number of particles %d is larger than Umbrella_sampling_max (%d)\n", N, n_max);
			sus_hist = calloc((size_t) (n_max - n_min + 1), sizeof(llint));
			if(sus_hist == NULL) die(&out, "out of memory\n");
		}
		if(n_max < N) die(&out, "The initial configuration holds %d particles, more than the maximum of %d\n", N, n_max);
		double *r2 = realloc(r, sizeof(double) * 3 * (size_t) (n_max + 1)), *o2 = realloc(orient, sizeof(double) * 9 * (size_t) (n_max + 1));
		if(r2 == NULL || o2 == NULL) die(&out, "out of memory\n");
		r = r2;
		orient = o2;
	}

	if(prm.ensemble == PP_NPT) {
		pph_get_double(in, "rescale_factor_max", &prm.rescale_factor_max, 1);
		pph_get_double(in, "P", &prm.P, 1);
	}
	else {
		pph_get_int(in, "Lx_move", &prm.Lx_move, 0);
		if(prm.Lx_move) {
			if(box[1] != box[2]) die(&out, "Lx_move = 1 requires Ly = Lz\n");
			pph_get_double(in, "Lx_change_max", &prm.Lx_change_max, 1);
			pph_get_double(in, "Lyz_min", &prm.Lyz_min, 1);
			pph_get_double(in, "Lyz_max", &prm.Lyz_max, 1);
			if(prm.Lyz_min >= box[1]) die(&out, "Lyz_min should be smaller than box size\n");
			if(prm.Lyz_max <= box[1]) die(&out, "Lyz_max should be larger than box size\n");
			log_msg(&out, "Ly and Lz are allowed to vary between %lf and %lf\n", prm.Lyz_min, prm.Lyz_max);
			out.boxshape = open_or_die(&out, "boxshape.dat", mode, "boxshape.dat");
		}
	}
	if(pph_get_int(in, "GPU_device", &itmp, 0)) prm.device = itmp;
	if(pph_get_int(in, "Lanes_per_cell", &itmp, 0)) prm.lanes_per_cell = itmp;
	if(pph_get_int(in, "Shift_every", &itmp, 0)) prm.shift_every = itmp;
	if(pph_get_int(in, "Lx_reference_sign", &itmp, 0)) prm.Lx_reference_sign = itmp;
	if(pph_get_int(in, "X_layers_multiple", &itmp, 0)) prm.x_layers_multiple = itmp; /* 2 x GPUs: every slab the same number of cell layers */

	/* GPUs = n: the system is split into n x-slabs, slab k on device GPU_device + k (GPU_devices = a b c ... names them
	 * explicitly; a device may be listed several times).  One host thread drives all of them (pp_group). */
	int n_gpus = 1, devices[64];
	if(pph_get_int(in, "GPUs", &itmp, 0)) n_gpus = itmp;
	if(n_gpus < 1 || n_gpus > 64) die(&out, "GPUs = %d out of range (1 .. 64)\n", n_gpus);
	for(int k = 0; k < n_gpus; k++) devices[k] = prm.device + k;
	{
		char list[512];
		if(pph_get_string(in, "GPU_devices", list, sizeof(list), 0)) {
			int k = 0;
			for(char *tok = strtok(list, " ,\t"); tok != NULL && k < n_gpus; tok = strtok(NULL, " ,\t")) devices[k++] = atoi(tok);
			if(k != n_gpus) die(&out, "GPU_devices lists %d devices, GPUs = %d\n", k, n_gpus);
		}
	}
	pp_group *h = NULL;
	if(pp_group_create(&prm, n_gpus, devices, &h) != 0) die(&out, "%s\n", h ? pp_group_last_error(h) : "pp_group_create failed");
	if(n_gpus > 1) log_msg(&out, "Running on %d x-slabs\n", n_gpus);
	pp_handle *engine0 = NULL;
	if(ensemble == PP_HOST_GC || ensemble == PP_HOST_SUS) {
		if(n_gpus != 1) die(&out, "Ensemble %d runs on a single engine (GPUs = 1)\n", ensemble);
		if(prm.Lx_move) die(&out, "Lx_move is not available with Ensemble %d on the GPU path\n", ensemble);
		PP_CALL(&out, h, pp_group_engine(h, 0, &engine0));
		PP_CALL(&out, h, pp_reserve(engine0, n_max)); /* System.particles holds N_max of them (system.c:199) */
	}
	PP_CALL(&out, h, pp_group_upload(h, N, box, r, orient));
	{
		pp_observables o;
		PP_CALL(&out, h, pp_group_get_observables(h, &o, 0));
		log_msg(&out, "Cells per side: (%d, %d, %d), total: %d\n", o.n_side[0], o.n_side[1], o.n_side[2], o.n_side[0] * o.n_side[1] * o.n_side[2]);
	}

	/* ---- initial energy and overlap check (main.c:55-64) ------------------------------------- */
	llint bonds = 0;
	int overlap = 0;
	PP_CALL(&out, h, pp_group_total_bonds(h, &bonds, &overlap));
	if(overlap) die(&out, "Initial configuration contains an overlap, aborting\n");
	log_msg(&out, "Initial energy: %lf\n", (N > 0) ? -(double) bonds / N : 0.);

	if(out.print_bonds) {
		out.bond_counts = malloc(sizeof(int) * (size_t) (N > 0 ? N : 1));
		out.bond_neighbours = malloc(sizeof(int) * (size_t) (N > 0 ? N : 1) * PP_HOST_MAX_BONDS);
		if(out.bond_counts == NULL || out.bond_neighbours == NULL) die(&out, "out of memory\n");
	}

	llint steps = 0;
	pph_get_llint(in, "Steps", &steps, 1);
	steps += out.start_from;

	/* ---- main loop (main.c:78-120) -------------------------------------------------------------- */
	llint step = out.start_from;
	while(step < steps && !stop_requested) {
		if(engine0 != NULL) { /* the particle number moves */
			pp_observables o;
			PP_CALL(&out, h, pp_group_get_observables(h, &o, 0));
			N = o.N;
		}
		if(step % out.print_every == 0) {
			print_observables(&out, h, &prm, ensemble, step);
			if(out.print_bonds) print_bonds(&out, h, N, step);
			save_configuration(&out, h, N, step, out.conf_last, r, orient);
			if(step != 0) { /* MC_check_energy + cells_check (main.c:85-87) */
				llint scratch;
				int ov, ok;
				PP_CALL(&out, h, pp_group_check_energy(h, &scratch, &ov, &ok));
				if(ov) die(&out, "\nComputing energy from scratch resulted in an overlap, aborting\n");
				if(!ok) die(&out, "\nEnergy check failed: %s\n", pp_group_last_error(h));
				if(sus_hist != NULL) { /* output_and_reset_sus (output.c:343-356) */
					snprintf(name, sizeof(name), "%s/sus-%lld.dat", sus_folder, step);
					FILE *f = fopen(name, "w");
					if(f == NULL) die(&out, "SUS file '%s' is not writable\n", name);
					for(int i = 0; i < n_max - n_min + 1; i++) {
						if(sus_hist[i] > 0) fprintf(f, "%d %lld\n", i + n_min, sus_hist[i]);
						sus_hist[i] = 0;
					}
					fclose(f);
				}
			}
		}
		if(step > 0 && step % out.save_every == 0) {
			snprintf(name, sizeof(name), "%s/conf_%lld.rrr", out.conf_folder, step);
			save_configuration(&out, h, N, step, name, r, orient);
			save_configuration(&out, h, N, step, out.conf_last, r, orient);
		}
		/* run up to the next print / save point, in chunks short enough to notice a signal */
		llint next = steps;
		const llint np = (step / out.print_every + 1) * out.print_every, ns = (step / out.save_every + 1) * out.save_every;
		if(np < next) next = np;
		if(ns < next) next = ns;
		llint chunk = next - step;
		if(chunk > 64) chunk = 64;
		if(engine0 != NULL) { /* do_GC / do_SUS (MC.c:31-58): the step's insertion / removal trials, then its particle moves */
			for(llint k = 0; k < chunk; k++) {
				int full = 0;
				PP_CALL(&out, h, pp_gc_step(engine0, activity, n_min, n_max, (long long *) sus_hist, NULL, &full));
				if(full && ensemble == PP_HOST_GC)
					die(&out, "The system contains the maximum number of particles set in the input file (%d), increase GC_N_max\n", n_max);
				PP_CALL(&out, h, pp_group_sweep(h, 1));
			}
		}
		else PP_CALL(&out, h, pp_group_sweep(h, chunk));
		PP_CALL(&out, h, pp_group_sync(h));
		step += chunk;
	}

	/* ---- last configuration and last line (main.c:125-132) ------------------------------------ */
	if(engine0 != NULL) {
		pp_observables o;
		PP_CALL(&out, h, pp_group_get_observables(h, &o, 0));
		N = o.N;
	}
	save_configuration(&out, h, N, step, out.conf_last, r, orient);
	{
		pp_observables o;
		PP_CALL(&out, h, pp_group_get_observables(h, &o, 0));
		printf("%lld %lf %lf", step, (o.N > 0) ? o.energy / o.N : 0., o.energy);
		if(ensemble != PP_NVT) printf(" %lf", o.N / o.volume);
		printf("\n");
	}

	free(sus_hist);
	pp_group_destroy(h);
	free(r);
	free(orient);
	free(out.bond_counts);
	free(out.bond_neighbours);
	fclose(out.energy);
	fclose(out.acc);
	if(out.density) fclose(out.density);
	if(out.boxshape) fclose(out.boxshape);
	if(out.log != stderr) fclose(out.log);
	pph_input_free(in);
	return 0;
}
