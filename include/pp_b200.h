/*
 * pp_b200.h -- C ABI of the B200-native Monte Carlo engine for the PatchyParticles hot path.
 *
 * This is the drop-in boundary: plain C, opaque handle, host pointers and sizes only.  Every entry
 * point names the reference interface it replaces (paths relative to the PatchyParticles tree).
 * The reference's plugin points are the two function pointers System.do_ensemble / do_dynamics
 * (defs.h:77-78) bound in MC_init (MC.c:85-131) and called once per MC step from main.c:119; the
 * GPU engine replaces the whole do_ensemble body (sweep granularity).  INTEGRATION.md shows the
 * binding a PatchyParticles maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pp_last_error() gives the message
 *     (the reference's convention is log-and-exit(1), output.c:374-388: the host maps non-zero
 *     to exactly that).
 *   - host arrays are caller-owned and copied; particle order on the ABI is always the caller's
 *     original order (the engine keeps particles cell-sorted internally).
 *   - positions r[N][3] and orientation matrices orient[N][3][3] are FP64, row-major, with the
 *     meaning of PatchyParticle.r / .orientation (defs.h:47-56).
 *   - a handle is not thread-safe; one handle drives one GPU (one slab of a multi-GPU run).
 *   - there is no CPU fallback: without a CUDA device pp_create fails.
 */
#ifndef PP_B200_H_
#define PP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PP_ABI_VERSION 2

/* move-type slots of System.tries[] / accepted[] (defs.h:11-19) */
#define PP_N_MOVES 8
#define PP_ROTO_TRASL 0
#define PP_VOLUME 6
#define PP_LX 7

/* ensembles (defs.h:21-25); only NVT and NPT are on the GPU path */
#define PP_NVT 0
#define PP_NPT 5

typedef struct pp_handle pp_handle;

/* Input-file keys of the hot path (system.c:61-196), same meaning and units. */
typedef struct pp_params {
	int ensemble;               /* Ensemble: PP_NVT or PP_NPT                       system.c:62 */
	double temperature;         /* Temperature                                      system.c:66 */
	double disp_max;            /* Disp_max                                         system.c:64 */
	double theta_max;           /* Theta_max                                        system.c:65 */
	double kf_delta;            /* KF_delta                                         system.c:83 */
	double kf_cosmax;           /* KF_cosmax                                        system.c:84 */
	double P;                   /* P (NPT)                                          system.c:168 */
	double rescale_factor_max;  /* rescale_factor_max (NPT)                         system.c:167 */
	int Lx_move;                /* Lx_move (NVT only, as in the reference)          system.c:171 */
	double Lx_change_max;       /* Lx_change_max                                    system.c:177 */
	double Lyz_min, Lyz_max;    /* Lyz_min / Lyz_max (absolute lengths)             system.c:178-179 */
	int Lx_reference_sign;      /* 1: reproduce exp(+dE/T) of MC.c:562; 0: Metropolis exp(-dE/T) */
	uint64_t seed;              /* Seed                                             system.c:71-75 */
	int device;                 /* CUDA device ordinal */
	int lanes_per_cell;         /* kernel choice of the sweep.  0 = automatic: by system size, either the tile kernel alone or the
	                               hybrid routing (sparsely populated cells on the thread-per-cell kernel, crowded ones on its
	                               8-lane configuration, concurrently); 1 = thread-per-cell kernel wherever a cell fits it;
	                               4, 8, 16, 32 = tile kernel of that width (for the crowded cells if cells_max_occupancy > 0) */
	int cell_lanes;             /* lanes per cell of the thread-per-cell kernel for the sparsely populated cells: 0 = 1, or 1, 2, 4, 8 */
	int cells_max_occupancy;    /* hybrid routing: cells with at most this many particles go to the thread-per-cell kernel;
	                               -1 = automatic (3 for N >= 800k, 2 for N >= 400k, 1 for N >= 150k, else 0), 0 = tile kernel only */
	int shift_every;            /* redraw the random grid shift (and rebuild the cells) every this many sweeps; 0 = default (4) */
} pp_params;

typedef struct pp_observables {
	long long step;                      /* sweeps done since pp_upload */
	int N;
	double box[3];                       /* System.box                                  defs.h:60 */
	double energy;                       /* System.energy = -(bonded pairs)             defs.h:65 */
	long long tries[PP_N_MOVES];         /* System.tries   (int in the reference)       defs.h:95 */
	long long accepted[PP_N_MOVES];      /* System.accepted                             defs.h:96 */
	/* checkerboard bookkeeping (DESIGN.md "acceptance estimator") */
	long long confined;                  /* proposals rejected for leaving their cell (no energy evaluated) */
	long long overlapped;                /* proposals rejected by a hard-core overlap */
	double acceptance_roto;              /* estimator of the reference's accepted/tries for ROTO_TRASL */
	long long pair_checks;               /* pair tests executed by the sweep kernel */
	int n_side[3];                       /* current (even) cell grid */
	double volume;                       /* System.V (system.c:198, MC.c:500): what density.dat divides by, output.c:131 */
} pp_observables;

void pp_default_params(pp_params *p);
int pp_abi_version(void);

/* system_init + MC_init (system.c:58-252, MC.c:85-131): allocate an engine for one GPU. */
int pp_create(const pp_params *params, pp_handle **out);
int pp_destroy(pp_handle *h);                           /* system_free, system.c:254-276 */
const char *pp_last_error(const pp_handle *h);          /* output_exit message, output.c:374-388 */

/* The configuration read from Initial_conditions_file (system.c:213-245) goes to the GPU; builds the
 * cell structure (cells_init + cells_fill, cells.c:17-63) as a counting sort.  Resets step/counters.
 * The running energy is NOT computed here (main.c:55-61 does that explicitly: call pp_total_bonds). */
int pp_upload(pp_handle *h, int N, const double *box, const double *r, const double *orient);
/* particles[i].r / .orientation as output_save reads them (output.c:302-317), original order */
int pp_download(pp_handle *h, double *r, double *orient);

/* 0.5 * sum_i MC_energy(i) with the overlap flag (main.c:55-61, MC.c:488-495, 556-560, 769-776).
 * bonds = number of bonded pairs (energy = -bonds).  Also (re)sets the running energy. */
int pp_total_bonds(pp_handle *h, long long *bonds, int *overlap);
/* per-particle bonded-neighbour counts, what Print_bonds reports (output.c:213-268) */
int pp_bond_counts(pp_handle *h, int *counts);
/* the two lines output_print_bonds writes per particle (output.c:222-264): counts[i] = number of bonded neighbours of
 * particle i, neighbours[i * max_per_particle + k] = their indices (first max_per_particle of them, -1 beyond the
 * count), listed cell by cell in the reference's 27-cell order and by index inside a cell (the reference follows its
 * linked lists and builds the line with an overlapping sprintf, output.c:250; compare as sets).  neighbours may be NULL. */
int pp_bond_lists(pp_handle *h, int max_per_particle, int *counts, int *neighbours);

/* do_ensemble x nsteps (main.c:119 -> do_NVT / do_NPT, MC.c:19-29,60-68): each step is N
 * roto-translation attempts (MC_move_rototranslate, MC.c:311-340) as one 8-colour checkerboard
 * sweep, plus Binomial(N, 1/N) volume moves (NPT, MC.c:450-513) or Lx moves (NVT + Lx_move,
 * MC.c:515-585).  Asynchronous: returns after enqueueing unless a box move needs a host decision. */
int pp_sweep(pp_handle *h, long long nsteps);
int pp_sync(pp_handle *h);

/* what output_print reads from System (output.c:119-196); reset != 0 zeroes tries/accepted the way
 * utils_reset_acceptance_counters does after each print (output.c:189). */
int pp_get_observables(pp_handle *h, pp_observables *obs, int reset);

/* MC_check_energy + cells_check (MC.c:766-784, cells.c:65-76): recompute the bond count from
 * scratch, re-orthonormalise every orientation (gram_schmidt, utils.c:91-113), verify the cell
 * population.  *ok = 0 if the running energy drifted by more than 1e-5 relative or an overlap
 * exists; the running energy is then replaced by the recomputed one, as in the reference. */
int pp_check_energy(pp_handle *h, long long *bonds_scratch, int *overlap, int *ok);

/* Test hook: MC.c:328-335 on a list of proposals.  commit == 0: all against the frozen current
 * configuration; commit != 0: sequential move list (accepted proposals are applied).
 * dE[k] = reference deltaE (E_new = 0 when the new position overlaps), overlap[k], accept[k]. */
int pp_eval_proposals(pp_handle *h, int n, const int *idx, const double *disp, const double *orient_new,
		const double *u, int commit, int *dE, int *overlap, int *accept);

/* Test hooks: one MC_change_volume (MC.c:450-513) / MC_change_Lx (MC.c:515-585) with the uniforms
 * the reference would draw, in its draw order.  *accepted: 1 yes, 0 no, -1 early reject (Lx only). */
int pp_volume_move(pp_handle *h, double u_dir, double u_vol, double u_acc, int *accepted);
int pp_Lx_move(pp_handle *h, double u_lx, double u_acc, int *accepted);

/* Plumbing for measurement: run on a caller-owned CUDA stream (cudaStream_t as void*), per-kernel
 * CUDA-event timing, launch counting. */
int pp_set_stream(pp_handle *h, void *cuda_stream);
int pp_set_profiling(pp_handle *h, int enable);
/* kernel classes for pp_kernel_stats */
#define PP_K_CELLS 0     /* K1: cell build (count, scan, scatter, order, gather) */
#define PP_K_SWEEP 1     /* K2: checkerboard sweep colour passes */
#define PP_K_BONDS 2     /* K3: full bond count */
#define PP_K_RESCALE 3   /* K4: position rescale */
#define PP_K_PROPOSALS 4 /* K5: proposal evaluator */
#define PP_K_ORTHO 5     /* K6: re-orthonormalise */
#define PP_K_RNG 6       /* K2a: proposal generation (Philox, rotation matrices) */
#define PP_K_HALO 7      /* slab decomposition: halo refresh and migration over peer memory (includes waiting for the neighbour) */
#define PP_K_CLASSES 8
int pp_kernel_stats(pp_handle *h, int kclass, long long *launches, double *ms_total);
int pp_reset_kernel_stats(pp_handle *h);

/* ---- x-slab decomposition across the GPUs of one box (one engine per GPU) -------------------------
 * The reference has no decomposition (single thread, main.c:26-143); this is the scale-out axis of
 * the engine for large N.  Each engine owns whole pairs of x layers of the global cell grid plus one
 * halo layer per side.  pp_slab_setup is called before pp_upload (every engine receives the whole
 * configuration and keeps its layers).  pp_total_bonds / pp_get_observables / pp_check_energy then
 * return this engine's contribution (a cross-boundary pair is counted by the engine on its left);
 * the driver sums over engines.  One sweep is driven as
 *     pp_slab_sweep_begin(&exchange)
 *     if exchange:  send pp_slab_outbox(side) to the neighbour on that side, receive into
 *                   pp_slab_inbox(side, count), then pp_slab_sweep_rebuild()
 *     pp_slab_pass_block(0);  first owned layer -> left neighbour's right halo   (pp_slab_halo views)
 *     pp_slab_pass_block(1);  last owned layer  -> right neighbour's left halo
 *     pp_slab_sweep_end()
 * Records are 13 doubles (x, y, z, id, orientation[9]); halo views are contiguous slices of the
 * engine's position (double4), orientation (9 doubles) and fixed-point shadow (uint2) arrays.  All
 * pointers returned here are DEVICE pointers.  NVT only. */
int pp_slab_setup(pp_handle *h, int rank, int world);
int pp_slab_range(pp_handle *h, int *x0, int *x1, int *nx, int *n_local);
int pp_slab_sweep_begin(pp_handle *h, int *exchange);
int pp_slab_outbox(pp_handle *h, int side, void **dev, int *count);
int pp_slab_inbox(pp_handle *h, int side, int count, void **dev);
int pp_slab_sweep_rebuild(pp_handle *h);
int pp_slab_pass_block(pp_handle *h, int cx);
int pp_slab_halo(pp_handle *h, int side, int recv, void **pos, void **ori, void **posf, int *count);
/* Halo refresh over peer memory, for engines on the GPUs of one NVLink / NVSwitch node (replaces the
 * library transfers of the pp_slab_halo slices).  Every engine exports a mailbox block as a 64-byte
 * cudaIpcMemHandle_t after pp_upload, the driver hands each engine the handles of its left and right
 * neighbour, and pp_slab_refresh(side) then (1) stores the engine's first (side 0) / last (side 1) owned
 * layer into the neighbour's mailbox and publishes a sequence number there, (2) waits on the device for the
 * matching number in its own mailbox and moves it into its right (side 0) / left (side 1) halo layer.  Nothing
 * synchronises with the host; a neighbour that never shows up is reported by pp_sync after 20 s. */
int pp_slab_peer_export(pp_handle *h, void *ipc_handle_64);
int pp_slab_peer_connect(pp_handle *h, const void *left_handle_64, const void *right_handle_64);
int pp_slab_refresh(pp_handle *h, int side);
int pp_slab_peer_close(pp_handle *h);
/* start of a shift period over peer memory: pp_slab_sweep_begin + record exchange + pp_slab_sweep_rebuild in one call */
int pp_slab_migrate(pp_handle *h, int *migrated);
/* whole sweeps over peer memory until max_sweeps are done or the next grid shift needs pp_slab_sweep_begin */
int pp_slab_sweeps(pp_handle *h, int max_sweeps, int *done);
int pp_slab_sweep_end(pp_handle *h);
int pp_slab_download(pp_handle *h, int *n_owned, int *ids, double *r, double *orient);

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H_ */
