// pp_group.cu -- one simulation in n x-slabs driven by ONE host thread (include/pp_b200.h, pp_group_*): the C host's
// `GPUs = n` path.  Slab k lives on CUDA device devices[k]; devices may repeat, so the whole decomposition -- mirror
// stores into the neighbour's halo, migration mailboxes, phase ordering, box moves decided on the sums over the engines --
// also runs, and is tested, on a single GPU.  Orderings between the engines' streams are CUDA events (pp_slab_api.cu);
// the group issues every phase for all its engines before the next one, which is what makes those events valid.
// n = 1 is a plain engine.  Same random streams as the single engine: a group follows its trajectory bit for bit.
#include "pp_internal.h"

using namespace pp;
using namespace ppi;

struct pp_group {
	int n;
	pp_handle **h;
	pp_params prm;
	char err[512];
	int N;
	uint64_t box_draws;
	long long box_tries[PP_N_MOVES], box_accepted[PP_N_MOVES];
};

namespace {

int gfail(pp_group *g, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g->err, sizeof(g->err), fmt, ap);
	va_end(ap);
	return 1;
}

// an engine's error becomes the group's
int efail(pp_group *g, int k) {
	snprintf(g->err, sizeof(g->err), "slab %d: %s", k, g->h[k]->err);
	return 1;
}

#define EACH(g, k) for(int k = 0; k < (g)->n; k++)
#define ENGINE_OK(g, k, call) \
	do {                      \
		if(call) return efail(g, k); \
	} while(0)

int connect_local(pp_group *g) {
	const int n = g->n;
	EACH(g, k) {
		pp_handle *h = g->h[k];
		peer_release(h);
		pp_handle *nb[2] = {g->h[(k + n - 1) % n], g->h[(k + 1) % n]};
		cudaSetDevice(h->device);
		for(int s = 0; s < 2; s++) {
			if(nb[s]->device != h->device) {
				int can = 0;
				cudaDeviceCanAccessPeer(&can, h->device, nb[s]->device);
				if(!can) return gfail(g, "device %d cannot access device %d's memory", h->device, nb[s]->device);
				cudaError_t e = cudaDeviceEnablePeerAccess(nb[s]->device, 0);
				if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return gfail(g, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
				cudaGetLastError();
			}
			h->peer_handle[s] = nb[s];
			h->peer_arena[s] = nb[s]->arena;
			h->peer_layout[s] = nb[s]->layout;
		}
		h->peer_same = n == 2;
		h->peer_ipc = false;
		h->peer_connected = true;
		h->phase = 0;
		h->mig_epoch = 0;
	}
	EACH(g, k) ENGINE_OK(g, k, slab_build_mirror(g->h[k]));
	return 0;
}

// distributes a configuration over the engines: each receives the particles of the layers it holds (owned + halo)
int distribute(pp_group *g, int N, const double *box, const double *r, const double *orient, bool keep_books) {
	if(g->n == 1) {
		ENGINE_OK(g, 0, upload_impl(g->h[0], N, N, nullptr, box, r, orient, keep_books));
		return 0;
	}
	unsigned char *mine = (unsigned char *) malloc((size_t) N + 1);
	int *ids = (int *) malloc(sizeof(int) * ((size_t) N + 1));
	double *rs = (double *) malloc(sizeof(double) * 3 * ((size_t) N + 1)), *os = (double *) malloc(sizeof(double) * 9 * ((size_t) N + 1));
	int rc = 0;
	if(!mine || !ids || !rs || !os) rc = gfail(g, "out of host memory");
	for(int k = 0; k < g->n && rc == 0; k++) {
		if(pp_slab_select(&g->prm, k, g->n, N, box, r, mine)) {
			rc = gfail(g, "the box (%g, %g, %g) cannot be split into %d slabs of at least two cell layers", box[0], box[1], box[2], g->n);
			break;
		}
		int m = 0;
		for(int i = 0; i < N; i++)
			if(mine[i]) {
				ids[m] = i;
				memcpy(rs + 3ll * m, r + 3ll * i, 3 * sizeof(double));
				memcpy(os + 9ll * m, orient + 9ll * i, 9 * sizeof(double));
				m++;
			}
		if(upload_impl(g->h[k], N, m, ids, box, rs, os, keep_books)) rc = efail(g, k);
	}
	free(mine);
	free(ids);
	free(rs);
	free(os);
	if(rc) return rc;
	return connect_local(g);
}

int gather(pp_group *g, double *r, double *orient) {
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_download(g->h[0], r, orient));
		return 0;
	}
	const size_t cap = (size_t) g->N + 1;
	int *ids = (int *) malloc(sizeof(int) * cap);
	double *rs = (double *) malloc(sizeof(double) * 3 * cap), *os = (double *) malloc(sizeof(double) * 9 * cap);
	unsigned char *seen = (unsigned char *) calloc(cap, 1);
	int rc = 0;
	if(!ids || !rs || !os || !seen) rc = gfail(g, "out of host memory");
	for(int k = 0; k < g->n && rc == 0; k++) {
		int m = 0;
		if(pp_slab_download(g->h[k], &m, ids, rs, os)) {
			rc = efail(g, k);
			break;
		}
		for(int j = 0; j < m; j++) {
			const int i = ids[j];
			if(i < 0 || i >= g->N || seen[i]) {
				rc = gfail(g, "particle %d is owned by no slab or by two", i);
				break;
			}
			seen[i] = 1;
			memcpy(r + 3ll * i, rs + 3ll * j, 3 * sizeof(double));
			memcpy(orient + 9ll * i, os + 9ll * j, 9 * sizeof(double));
		}
	}
	for(int i = 0; i < g->N && rc == 0; i++)
		if(!seen[i]) rc = gfail(g, "particle %d is owned by no slab", i);
	free(ids);
	free(rs);
	free(os);
	free(seen);
	return rc;
}

int sum_running(pp_group *g, long long *total) {
	*total = 0;
	EACH(g, k) {
		ENGINE_OK(g, k, fetch_counters(g->h[k]));
		*total += running_bonds(g->h[k]);
	}
	return 0;
}

int recount(pp_group *g, long long *bonds, int *overlap) {
	*bonds = 0;
	*overlap = 0;
	EACH(g, k) {
		long long b = 0;
		int ov = 0;
		ENGINE_OK(g, k, pp_total_bonds(g->h[k], &b, &ov));
		*bonds += b;
		*overlap |= ov;
	}
	return 0;
}

// a trial box that needs another cell grid changes the partition into slabs: the configuration is gathered, scaled on
// the host with the device's arithmetic (one IEEE multiplication / division per coordinate, MC.c:474-478, 505-508) and
// distributed again; the books (step, counters, draws) are kept
int rescale_and_redistribute(pp_group *g, const double *box, const double *f, bool divide) {
	const size_t N = (size_t) g->N;
	double *r = (double *) malloc(sizeof(double) * 3 * (N + 1)), *o = (double *) malloc(sizeof(double) * 9 * (N + 1));
	if(!r || !o) {
		free(r);
		free(o);
		return gfail(g, "out of host memory");
	}
	int rc = gather(g, r, o);
	if(rc == 0) {
		for(size_t i = 0; i < N; i++)
			for(int k = 0; k < 3; k++) r[3 * i + k] = divide ? r[3 * i + k] / f[k] : r[3 * i + k] * f[k];
		rc = distribute(g, g->N, box, r, o, true);
	}
	free(r);
	free(o);
	return rc;
}

int group_box_move(pp_group *g, int kind, const double *u, int *accepted) {
	const int n = g->n;
	long long old_total = 0;
	if(sum_running(g, &old_total)) return 1;
	long long new_total = 0;
	int overlap = 0, status = 0;
	EACH(g, k) {
		long long b = 0;
		int ov = 0, st = 0;
		ENGINE_OK(g, k, pp_slab_box_trial(g->h[k], kind, u, old_total, &b, &ov, &st));
		if(k > 0 && st != status) return gfail(g, "slab engines disagree on a box move (status %d / %d)", status, st);
		status = st;
		new_total += b;
		overlap |= ov;
	}
	const int slot = kind == PP_VOLUME ? PP_VOLUME : PP_LX;
	g->box_tries[slot]++;
	if(status < 0) {
		*accepted = status == -1 ? -1 : 0;
		return 0;
	}
	if(status == 2 && n > 1) {
		const BoxMove bm = g->h[0]->bm; // (every engine proposed the same move)
		const double V_old = g->h[0]->V;
		if(rescale_and_redistribute(g, bm.nb, bm.f, false)) return 1;
		if(recount(g, &new_total, &overlap)) return 1;
		g->h[0]->bm = bm;
		const bool acc = box_decide(g->h[0], new_total, overlap, old_total, g->N);
		if(acc) {
			EACH(g, k) g->h[k]->V = kind == PP_VOLUME ? bm.final_V : V_old;
		}
		else {
			if(rescale_and_redistribute(g, bm.old_box, bm.f, true)) return 1;
			long long b;
			int ov;
			if(recount(g, &b, &ov)) return 1;
			EACH(g, k) g->h[k]->V = V_old;
		}
		EACH(g, k) {
			g->h[k]->bm.pending = false;
			g->h[k]->sweep_grid = false;
		}
		if(acc) g->box_accepted[slot]++;
		*accepted = acc ? 1 : 0;
		return 0;
	}
	int acc = 0;
	EACH(g, k) {
		int a = 0;
		ENGINE_OK(g, k, pp_slab_box_finish(g->h[k], new_total, overlap, &a));
		if(k > 0 && a != acc) return gfail(g, "slab engines disagree on a box move");
		acc = a;
	}
	if(acc && n > 1) EACH(g, k) ENGINE_OK(g, k, slab_build_mirror(g->h[k])); // everybody switched buffers
	if(acc) g->box_accepted[slot]++;
	*accepted = acc;
	return 0;
}

// the engines' grid no longer is the even grid of the box (an expansion made room for more cells): a single engine picks
// the new grid up at its next cell build; a decomposition has to be split again
int regrid_if_needed(pp_group *g) {
	int ns[3];
	pp_handle *h0 = g->h[0];
	if(even_grid(h0->box, h0->r_cut, ns, h0->prm.x_layers_multiple)) return gfail(g, "box too small for the checkerboard");
	if(ns[0] == h0->n_side[0] && ns[1] == h0->n_side[1] && ns[2] == h0->n_side[2]) return 0;
	const double one[3] = {1., 1., 1.};
	const double V = h0->V;
	double box[3] = {h0->box[0], h0->box[1], h0->box[2]};
	long long before = 0, after = 0;
	int ov = 0;
	if(sum_running(g, &before)) return 1;
	if(rescale_and_redistribute(g, box, one, false)) return 1;
	if(recount(g, &after, &ov)) return 1;
	if(after != before || ov) return gfail(g, "re-partitioning changed the bond count (%lld -> %lld, overlap %d)", before, after, ov);
	EACH(g, k) g->h[k]->V = V;
	return 0;
}

} // namespace

extern "C" {

int pp_group_create(const pp_params *params, int n, const int *devices, pp_group **out) {
	if(params == nullptr || out == nullptr || n < 1) return 1;
	pp_group *g = (pp_group *) calloc(1, sizeof(pp_group));
	if(g == nullptr) return 1;
	*out = g;
	g->n = n;
	g->prm = *params;
	g->h = (pp_handle **) calloc((size_t) n, sizeof(pp_handle *));
	if(g->h == nullptr) return gfail(g, "out of host memory");
	EACH(g, k) {
		pp_params p = *params;
		p.device = devices ? devices[k] : params->device;
		if(pp_create(&p, &g->h[k])) {
			if(g->h[k]) return efail(g, k);
			return gfail(g, "pp_create failed for slab %d", k);
		}
		if(n > 1) ENGINE_OK(g, k, pp_slab_setup(g->h[k], k, n));
	}
	return 0;
}

int pp_group_destroy(pp_group *g) {
	if(g == nullptr) return 0;
	if(g->h) {
		EACH(g, k) if(g->h[k]) {
			cudaSetDevice(g->h[k]->device);
			cudaStreamSynchronize(g->h[k]->stream);
		}
		EACH(g, k) pp_destroy(g->h[k]);
		free(g->h);
	}
	free(g);
	return 0;
}

const char *pp_group_last_error(const pp_group *g) {
	return g ? g->err : "null group";
}

int pp_group_engine(pp_group *g, int k, pp_handle **h) {
	if(g == nullptr || h == nullptr || k < 0 || k >= g->n) return 1;
	*h = g->h[k];
	return 0;
}

int pp_group_upload(pp_group *g, int N, const double *box, const double *r, const double *orient) {
	if(g == nullptr || N < 0 || box == nullptr) return 1;
	g->N = N;
	g->box_draws = 0;
	memset(g->box_tries, 0, sizeof(g->box_tries));
	memset(g->box_accepted, 0, sizeof(g->box_accepted));
	return distribute(g, N, box, r, orient, false);
}

int pp_group_download(pp_group *g, double *r, double *orient) {
	if(g == nullptr) return 1;
	return gather(g, r, orient);
}

int pp_group_total_bonds(pp_group *g, long long *bonds, int *overlap) {
	if(g == nullptr) return 1;
	long long b = 0;
	int ov = 0;
	if(recount(g, &b, &ov)) return 1;
	if(bonds) *bonds = b;
	if(overlap) *overlap = ov;
	return 0;
}

int pp_group_bond_lists(pp_group *g, int max_per_particle, int *counts, int *neighbours) {
	if(g == nullptr) return 1;
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_bond_lists(g->h[0], max_per_particle, counts, neighbours));
		return 0;
	}
	// the per-particle lists are an output call, not a hot one: gather and count on the first engine alone
	const size_t N = (size_t) g->N;
	double *r = (double *) malloc(sizeof(double) * 3 * (N + 1)), *o = (double *) malloc(sizeof(double) * 9 * (N + 1));
	int rc = (!r || !o) ? gfail(g, "out of host memory") : gather(g, r, o);
	pp_handle *single = nullptr;
	if(rc == 0) {
		pp_params p = g->prm;
		p.device = g->h[0]->device;
		if(pp_create(&p, &single) || pp_upload(single, g->N, g->h[0]->box, r, o) || pp_bond_lists(single, max_per_particle, counts, neighbours))
			rc = gfail(g, "bond lists: %s", single ? single->err : "pp_create failed");
	}
	pp_destroy(single);
	free(r);
	free(o);
	return rc;
}

int pp_group_sync(pp_group *g) {
	if(g == nullptr) return 1;
	EACH(g, k) ENGINE_OK(g, k, pp_sync(g->h[k]));
	return 0;
}

int pp_group_sweep(pp_group *g, long long nsteps) {
	if(g == nullptr) return 1;
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_sweep(g->h[0], nsteps));
		return 0;
	}
	if(g->N == 0) {
		EACH(g, k) g->h[k]->step += nsteps;
		return 0;
	}
	for(long long s = 0; s < nsteps; s++) {
		if(g->prm.ensemble == PP_NPT || g->prm.Lx_move) { // Binomial(N, 1/N) box moves per step (MC.c:24-27, 62-63), the single engine's stream
			const int kind = g->prm.ensemble == PP_NPT ? PP_VOLUME : PP_LX;
			const int m = binomial_one_over_N(g->N, box_uniform(g->prm.seed, g->box_draws++));
			for(int j = 0; j < m; j++) {
				double u[3] = {0., 0., 0.};
				for(int c = 0; c < (kind == PP_VOLUME ? 3 : 2); c++) u[c] = box_uniform(g->prm.seed, g->box_draws++);
				int acc;
				if(group_box_move(g, kind, u, &acc)) return 1;
			}
		}
		if(slab_period_starts(g->h[0])) {
			if(regrid_if_needed(g)) return 1;
			EACH(g, k) ENGINE_OK(g, k, slab_rebuild_begin(g->h[k]));
			EACH(g, k) ENGINE_OK(g, k, slab_rebuild_end(g->h[k]));
			EACH(g, k) ENGINE_OK(g, k, slab_build_mirror(g->h[k]));
			EACH(g, k) ENGINE_OK(g, k, slab_signal(g->h[k], true));
		}
		for(int cx = 0; cx < 2; cx++) EACH(g, k) ENGINE_OK(g, k, slab_block(g->h[k], cx));
		EACH(g, k) g->h[k]->step++;
	}
	return 0;
}

int pp_group_get_observables(pp_group *g, pp_observables *obs, int reset) {
	if(g == nullptr || obs == nullptr) return 1;
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_get_observables(g->h[0], obs, reset));
		return 0;
	}
	pp_observables sum;
	memset(&sum, 0, sizeof(sum));
	EACH(g, k) {
		pp_observables o;
		ENGINE_OK(g, k, pp_get_observables(g->h[k], &o, reset));
		if(k == 0) sum = o;
		else {
			sum.energy += o.energy;
			sum.tries[PP_ROTO_TRASL] += o.tries[PP_ROTO_TRASL];
			sum.accepted[PP_ROTO_TRASL] += o.accepted[PP_ROTO_TRASL];
			sum.confined += o.confined;
			sum.overlapped += o.overlapped;
			sum.w_tries += o.w_tries;
			sum.w_accepted += o.w_accepted;
			sum.pair_checks += o.pair_checks;
		}
	}
	for(int m = 1; m < PP_N_MOVES; m++) { // box moves are counted once, by the group
		sum.tries[m] = g->box_tries[m];
		sum.accepted[m] = g->box_accepted[m];
	}
	sum.N = g->N;
	sum.acceptance_roto = sum.w_tries > 0 ? (double) sum.w_accepted / (double) sum.w_tries : NAN;
	if(reset) {
		memset(g->box_tries, 0, sizeof(g->box_tries));
		memset(g->box_accepted, 0, sizeof(g->box_accepted));
	}
	*obs = sum;
	return 0;
}

int pp_group_check_energy(pp_group *g, long long *bonds_scratch, int *overlap, int *ok) {
	if(g == nullptr) return 1;
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_check_energy(g->h[0], bonds_scratch, overlap, ok));
		return 0;
	}
	// only the sums over the engines are comparable (MC_check_energy, MC.c:766-784)
	long long running = 0, b_total = 0;
	int ov_any = 0, good = 1;
	if(sum_running(g, &running)) return 1;
	EACH(g, k) {
		long long b = 0;
		int ov = 0, okk = 0;
		ENGINE_OK(g, k, pp_check_energy(g->h[k], &b, &ov, &okk));
		b_total += b;
		ov_any |= ov;
		if(!okk) good = 0;
	}
	const double E_run = -(double) running, E_new = -(double) b_total;
	if(ov_any) good = 0;
	if(fabs(E_run) > 1e-5 && fabs((E_new - E_run) / E_run) > 1e-5) good = 0; // MC.c:779
	if(bonds_scratch) *bonds_scratch = b_total;
	if(overlap) *overlap = ov_any;
	if(ok) *ok = good;
	if(!good) gfail(g, "energy check failed: running bonds %lld, recomputed %lld, overlap %d", running, b_total, ov_any);
	return 0;
}

int pp_group_volume_move(pp_group *g, double u_dir, double u_vol, double u_acc, int *accepted) {
	if(g == nullptr) return 1;
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_volume_move(g->h[0], u_dir, u_vol, u_acc, accepted));
		return 0;
	}
	const double u[3] = {u_dir, u_vol, u_acc};
	int acc = 0;
	if(group_box_move(g, PP_VOLUME, u, &acc)) return 1;
	if(accepted) *accepted = acc;
	return 0;
}

int pp_group_Lx_move(pp_group *g, double u_lx, double u_acc, int *accepted) {
	if(g == nullptr) return 1;
	if(g->n == 1) {
		ENGINE_OK(g, 0, pp_Lx_move(g->h[0], u_lx, u_acc, accepted));
		return 0;
	}
	const double u[3] = {u_lx, u_acc, 0.};
	int acc = 0;
	if(group_box_move(g, PP_LX, u, &acc)) return 1;
	if(accepted) *accepted = acc;
	return 0;
}

} // extern "C"
