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

#define PP_ABI_VERSION 4

/* move-type slots of System.tries[] / accepted[] (defs.h:11-19) */
#define PP_N_MOVES 8
#define PP_ROTO_TRASL 0
#define PP_ADD 2
#define PP_REMOVE 3
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
	int lanes_per_cell;         /* kernel choice of the sweep.  0 = automatic: by system size, either the tile kernel alone or one fused
	                               launch per colour pass in which sparsely populated cells run on one thread per cell, the middle
	                               range on 4 lanes and the crowded cells on 8 lanes per cell; 1 = one thread per cell wherever a
	                               cell fits it; 4, 8, 16, 32 = tile kernel of that width for every cell */
	int cell_lanes;             /* 2, 4, 8 (or -1: one thread): that many lanes per cell for the sparsely populated range (the three ranges then run as
	                               three launches); 0 or 1: one thread per cell inside the fused launch */
	int cells_max_occupancy;    /* cells with at most this many particles run on one thread per cell;
	                               -1 = automatic (3 for N >= 800k, 2 for N >= 400k, 1 for N >= 100k, else 0), 0 = tile kernel only */
	int shift_every;            /* redraw the random grid shift (and rebuild the cells) every this many sweeps; 0 = default (4) */
	int cells_mid_occupancy;    /* cells with more particles than cells_max_occupancy and at most this many run on 4 lanes per cell, the
	                               more crowded ones on 8; -1 = automatic */
	int x_layers_multiple;      /* the number of cell layers along x is rounded down to a multiple of this (even) number; 0 = 2 (any even
	                               grid).  2 n for a run in n x-slabs gives every slab the same number of layers; a single engine given the
	                               same value follows the same trajectory */
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
	unsigned long long w_tries, w_accepted; /* the sums behind acceptance_roto (fixed point 2^-24): additive over the engines of a decomposition */
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

/* ---- grand-canonical insertion / removal: MC_add_remove, MC.c:342-448, as do_GC / do_SUS call it
 * (MC.c:31-58).  The engine is created with Ensemble = 0; the host alternates pp_gc_step and pp_sweep.  Single engine only. */
typedef struct pp_gc_result {
	int kind;        /* PP_ADD, PP_REMOVE, or -1: returned before a try (full system on insertion, N == n_min on removal) */
	int accepted;
	int overlap;     /* System.overlap as MC_energy left it */
	int bonds;       /* -MC_energy of the inserted / removed particle (0 with an overlap, MC.c:296-299) */
	int index;       /* index of the new particle (N) or of the one picked for removal */
	int n_uniforms;  /* uniforms consumed */
	int N;           /* particles after the move */
} pp_gc_result;
/* make room for up to n_max particles (GC_N_max / Umbrella_sampling_max, system.c:94-99, 199); call before pp_upload */
int pp_reserve(pp_handle *h, int n_max);
/* -MC_energy(syst, particles + id) (MC.c:272-309): bonded neighbours of one particle and the overlap flag, on the GPU */
int pp_particle_bonds(pp_handle *h, int id, int *bonds, int *overlap);
/* the same for a particle that is not there yet (MC.c:350-361): position r[3], orientation matrix orient[9] */
int pp_probe_bonds(pp_handle *h, const double *r, const double *orient, int *bonds, int *overlap);
/* Test hook: one MC_add_remove with the uniforms the reference would draw, in its order (MC.c:344, 352-354 + random_orientation,
 * 364 / 383, 387) */
int pp_add_remove(pp_handle *h, double activity, int n_min, int n_max, const double *u, int n_u, pp_gc_result *res);
/* the insertion / removal trials of one MC step: Binomial(n_max, 0.01) of them (MC.c:33-36, 48-52) from the engine's own stream.
 * sus_hist (may be NULL): SUS_hist[N - n_min]++ after every trial (MC.c:51); *full = 1 if an insertion met N == n_max (fatal
 * under GC, MC.c:346). */
int pp_gc_step(pp_handle *h, double activity, int n_min, int n_max, long long *sus_hist, int *n_trials, int *full);

/* ---- neighbour queries of the cluster moves; the moves themselves stay on the host.  A batch of n particles
 * per call, answered from the cell-sorted state with the reference's pair arithmetic; single engine only. */
/* _set_neighbours (avb.c:43-75): neighbours[k * 4 + patch] = index of the neighbour bonded through that patch of particle idx[k]
 * (MC_interact's onp, MC.c:249-261; -1: free; where two neighbours share a patch the last one in the reference's cell order and
 * ascending index inside a cell stays), num_neighbours[k] = number of bonded neighbours */
int pp_patch_neighbours(pp_handle *h, int n, const int *idx, int *neighbours, int *num_neighbours);
/* the geometric part of _populate_possible_links (vmmc.c:148-229): the particles within the interaction range of idx[k]
 * (dist2 < kf_sqr_rcut, vmmc.c:178), first max_per_particle of them in the same order, -1 beyond counts[k] */
int pp_range_neighbours(pp_handle *h, int n, const int *idx, int max_per_particle, int *counts, int *neighbours);

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
 * engine's position (double4), orientation (9 doubles) and fixed-point shadow (uint4) arrays.  All
 * pointers returned here are DEVICE pointers. */
int pp_slab_setup(pp_handle *h, int rank, int world);
int pp_slab_range(pp_handle *h, int *x0, int *x1, int *nx, int *n_local);
int pp_slab_sweep_begin(pp_handle *h, int *exchange);
int pp_slab_outbox(pp_handle *h, int side, void **dev, int *count);
int pp_slab_inbox(pp_handle *h, int side, int count, void **dev);
int pp_slab_sweep_rebuild(pp_handle *h);
int pp_slab_pass_block(pp_handle *h, int cx);
int pp_slab_halo(pp_handle *h, int side, int recv, void **pos, void **ori, void **posf, int *count);
/* ---- the data path of the decomposition: stores into the neighbouring engine's memory, no mailbox copy -----------
 * Engines on the GPUs of one NVLink / NVSwitch node (or several engines on one GPU).  Every engine keeps its whole
 * device state in one allocation and describes it with a pp_peer_info record, a CUDA IPC handle + offsets; the driver hands
 * each engine the descriptions of its left and right neighbour (pp_slab_peer_connect).  From then on
 *   - the sweep kernels themselves keep the neighbours' halo layers current: the thread that commits an accepted move
 *     of a particle in the engine's first (last) owned layer also stores the new position, orientation and shadow into
 *     the left (right) neighbour's halo copy over NVLink -- no push / pull kernels, nothing to overlap;
 *   - neighbouring engines only exchange a phase number (a release store into the neighbour's control words after the
 *     rebuild and after each block of four colour passes, an acquire spin in a one-thread kernel before the block that
 *     needs the neighbour's data); nothing synchronises with the host, a neighbour that never shows up is reported by
 *     pp_sync after 20 s;
 *   - at a grid shift the migrating particles and the fresh halo layers travel as 13-double records stored into the
 *     neighbours' migration mailboxes (pp_slab_migrate), one host synchronisation per shift period.
 * pp_upload_subset gives an engine only the particles of the layers it holds (owned + halo). */
typedef struct pp_peer_info {
	unsigned char ipc[64];              /* cudaIpcMemHandle_t of the engine's arena */
	unsigned long long off_pos[3];      /* byte offsets inside the arena: position buffers (double4) */
	unsigned long long off_ori[2];      /* orientation buffers (9 doubles per particle) */
	unsigned long long off_posf;        /* fixed-point shadows (uint4: box fractions x, y, z + slot) */
	unsigned long long off_mig[2];      /* migration mailboxes: records from the left / right neighbour */
	unsigned long long off_ctrl;        /* control words: phases completed by the left / right neighbour, migration flags and counts */
	long long cap_records;              /* records per mailbox */
} pp_peer_info;
int pp_slab_peer_export(pp_handle *h, pp_peer_info *info);
int pp_slab_peer_connect(pp_handle *h, const pp_peer_info *left, const pp_peer_info *right);
int pp_slab_peer_close(pp_handle *h);
/* particles with their global ids (ascending ids are not required); N_global = particles of the whole system.
 * pp_slab_select marks the particles engine `rank` of `world` has to receive for a given box (layer of each particle
 * on the unshifted grid, exactly as the engine computes it). */
int pp_upload_subset(pp_handle *h, int N_global, int n_local, const int *ids, const double *box, const double *r, const double *orient);
/* every engine is handed the WHOLE configuration and takes the particles of its own layers: from pinned host memory the
 * selection and the gather run on the GPU (one kernel reading over PCIe, no host-side work, no staging copy); from
 * pageable memory they run on the host.  Without a decomposition this is pp_upload. */
int pp_upload_select(pp_handle *h, int N, const double *box, const double *r, const double *orient);
int pp_slab_select(const pp_params *params, int rank, int world, int N, const double *box, const double *r, unsigned char *mine);
/* start of a shift period: classify on the new grid, store the outgoing records into the neighbours' mailboxes, take in
 * what they stored here, rebuild the cells.  *migrated = 0 when the current period is still running. */
int pp_slab_migrate(pp_handle *h, int *migrated);
/* whole sweeps (two blocks of four colour passes, each after its neighbour's phase) until max_sweeps are done or the
 * next grid shift needs pp_slab_migrate */
int pp_slab_sweeps(pp_handle *h, int max_sweeps, int *done);
/* box moves under the decomposition (MC_change_volume / MC_change_Lx, MC.c:450-585): every engine rescales and counts
 * its own layers (pp_slab_box_trial), the driver sums {bonds, overlap} over the engines and every engine takes the
 * same decision from the same uniforms (pp_slab_box_finish).  kind: PP_VOLUME or PP_LX; u: the move's uniforms in the
 * reference's draw order.  *status: 0 evaluated; -1 rejected before any work by the Lyz window (MC.c:527-530), -2 because
 * the trial box cannot hold two cells per side; 2 the trial box needs another cell grid (nothing evaluated: the driver
 * gathers the configuration, scales it and uploads it again in the trial box). */
int pp_slab_box_trial(pp_handle *h, int kind, const double *u, long long bonds_running_total, long long *bonds_local, int *overlap_local, int *status);
int pp_slab_box_finish(pp_handle *h, long long bonds_total, int overlap_any, int *accepted);
int pp_slab_sweep_end(pp_handle *h);
int pp_slab_download(pp_handle *h, int *n_owned, int *ids, double *r, double *orient);


/* ---- several engines driven from one host thread (the C host's `GPUs = n` key) --------------------------------------
 * A group is one simulation in n x-slabs, slab k on CUDA device devices[k]; devices may repeat (several slabs on one
 * GPU run through the same peer stores and orderings, which is how the decomposition is tested on a single GPU).
 * Same calls and meaning as the single-engine API; orderings between slabs are CUDA events, no process group needed.
 * n = 1 is a plain engine. */
typedef struct pp_group pp_group;
int pp_group_create(const pp_params *params, int n, const int *devices, pp_group **out);
int pp_group_destroy(pp_group *g);
const char *pp_group_last_error(const pp_group *g);
int pp_group_upload(pp_group *g, int N, const double *box, const double *r, const double *orient);
int pp_group_download(pp_group *g, double *r, double *orient);
int pp_group_total_bonds(pp_group *g, long long *bonds, int *overlap);
int pp_group_bond_lists(pp_group *g, int max_per_particle, int *counts, int *neighbours);
int pp_group_sweep(pp_group *g, long long nsteps);
int pp_group_sync(pp_group *g);
int pp_group_get_observables(pp_group *g, pp_observables *obs, int reset);
int pp_group_check_energy(pp_group *g, long long *bonds_scratch, int *overlap, int *ok);
int pp_group_volume_move(pp_group *g, double u_dir, double u_vol, double u_acc, int *accepted);
int pp_group_Lx_move(pp_group *g, double u_lx, double u_acc, int *accepted);
int pp_group_engine(pp_group *g, int k, pp_handle **h); /* slab k's engine (kernel statistics, tests) */

#ifdef __cplusplus
}
#endif
#endif /* PP_B200_H_ */
