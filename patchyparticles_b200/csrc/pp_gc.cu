// pp_gc.cu -- grand-canonical insertion / removal: MC_add_remove (MC.c:342-448) as the
// reference's do_GC / do_SUS drivers call it (MC.c:31-58), with the one quantity that costs anything -- MC_energy of the
// inserted or removed particle (MC.c:272-309) -- evaluated on the GPU against the cell-sorted state.  Host-driven: a trial
// is one probe launch and one 8-byte read-back; an accepted trial edits the particle arrays in place (one small kernel)
// and marks the cells stale, so the next sweep or count re-sorts.
//
// Index convention of the reference kept: a new particle gets index N; when particle p is removed the LAST particle
// (index N - 1) takes index p (MC.c:414-438).
#include "pp_internal.h"

using namespace pp;
using namespace ppi;

namespace pp {

// bonds and overlap flag of one particle against everything else: an existing one (id >= 0, itself skipped) or a ghost
// with the given position and matrix (id < 0).  One tile of 32 lanes splits the 27-cell neighbourhood.
__global__ void __launch_bounds__(32) k_probe_particle(const __grid_constant__ SimConst sc, int id, const double *__restrict__ ghost /* r[3], orient[9] */,
		const int *__restrict__ cell_start, const double4 *__restrict__ pos, const double *__restrict__ ori, const int *__restrict__ slot_of,
		int *__restrict__ out /* bonds, overlap */) {
	const int lane = threadIdx.x;
	int self = -1;
	Vec3 p;
	double m[9], pp[12];
	if(id >= 0) {
		self = slot_of[id];
		const double4 p4 = pos[self];
		p = {p4.x, p4.y, p4.z};
#pragma unroll
		for(int k = 0; k < 9; k++) m[k] = ori[9ll * self + k];
	}
	else {
		p = {ghost[0], ghost[1], ghost[2]};
#pragma unroll
		for(int k = 0; k < 9; k++) m[k] = ghost[3 + k];
	}
	patches_from(sc.base, m, pp); // patches[i] = orientation . base_patches[i]  (MC.c:356-358)
	int ix, iy, iz;
	cell_index(sc, p.x, p.y, p.z, ix, iy, iz); // the cell of the current position, not a cached one (MC.c:277)
	const int nx = sc.n_side[0], ny = sc.n_side[1], nz = sc.n_side[2];
	int bonds = 0, overlap = 0;
	// the distinct cells of the neighbourhood (an axis with two cells has two, not three)
	const int wx = nx < 3 ? nx : 3, wy = ny < 3 ? ny : 3, wz = nz < 3 ? nz : 3;
	for(int c = 0; c < wx * wy * wz; c++) {
		const int dx = c / (wy * wz), dy = (c / wz) % wy, dz = c % wz;
		const int cx = (ix + dx - 1 + nx) % nx, cy = (iy + dy - 1 + ny) % ny, cz = (iz + dz - 1 + nz) % nz;
		const int cell = (cx * ny + cy) * nz + cz;
		for(int s = cell_start[cell] + lane; s < cell_start[cell + 1]; s += 32) {
			if(s == self) continue;
			const int v = pair_test(sc, p, pp, pos[s], ori + 9ll * s);
			bonds += (v == 1);
			overlap |= (v == -1);
		}
	}
#pragma unroll
	for(int o = 16; o > 0; o >>= 1) {
		bonds += __shfl_xor_sync(0xffffffffu, bonds, o);
		overlap |= __shfl_xor_sync(0xffffffffu, overlap, o);
	}
	if(lane == 0) {
		out[0] = bonds;
		out[1] = overlap;
	}
}

// a new particle with index N at slot N (MC.c:347-375); the next cell build sorts it in
__global__ void k_insert_particle(int N, const double *__restrict__ ghost, double4 *pos, double *ori, int *pid, int *slot_of) {
	if(threadIdx.x == 0) {
		pos[N] = make_double4(ghost[0], ghost[1], ghost[2], (double) N);
		pid[N] = N;
		slot_of[N] = N;
	}
	if(threadIdx.x < 9) ori[9ll * N + threadIdx.x] = ghost[3 + threadIdx.x];
}

// particle p leaves, the last one (index N - 1) takes its index (MC.c:389-438); the slots stay dense: whoever sits in
// the last slot moves into the one that falls vacant
__global__ void k_remove_particle(int N, int p, double4 *pos, double *ori, int *pid, int *slot_of) {
	if(threadIdx.x != 0) return;
	const int q = N - 1;
	int vacant = slot_of[p];
	if(p != q) {
		const int sq = slot_of[q];
		const double4 s = pos[sq];
		pos[vacant] = make_double4(s.x, s.y, s.z, (double) p);
		for(int k = 0; k < 9; k++) ori[9ll * vacant + k] = ori[9ll * sq + k];
		pid[vacant] = p; // slot_of[p] stays: the former q lives there now
		vacant = sq;
	}
	const int last = N - 1;
	if(vacant != last) {
		const int r = pid[last];
		pos[vacant] = pos[last];
		for(int k = 0; k < 9; k++) ori[9ll * vacant + k] = ori[9ll * last + k];
		pid[vacant] = r;
		slot_of[r] = vacant;
	}
}

} // namespace pp

namespace {

// the uniforms of one MC_add_remove: a recorded list (tests: the reference's own drand48 stream) or the engine's stream
struct Uniforms {
	const double *list;
	int n, used;
	pp_handle *h;
	bool exhausted;
	double next() {
		if(list != nullptr) {
			if(used >= n) {
				exhausted = true;
				return 0.75;
			}
			return list[used++];
		}
		used++;
		return box_uniform(h->prm.seed ^ 0x9E3779B97F4A7C15ull, h->box_draws++);
	}
};

inline double dot3h(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// random_vector_on_sphere (utils.c:115-131): Marsaglia, pairs drawn until inside the unit disc
void sphere_point(Uniforms &u, double *res) {
	double ran1, ran2, ransq;
	do {
		ran1 = 1. - 2. * u.next();
		ran2 = 1. - 2. * u.next();
		ransq = ran1 * ran1 + ran2 * ran2;
	} while(ransq >= 1. && !u.exhausted);
	const double ranh = 2. * sqrt(1. - ransq);
	res[0] = ran1 * ranh;
	res[1] = ran2 * ranh;
	res[2] = 1. - 2. * ransq;
}

void normalise_h(double *v) {
	const double norm = sqrt(dot3h(v, v));
	v[0] /= norm;
	v[1] /= norm;
	v[2] /= norm;
}

// random_orientation (utils.c:133-169): three points on the sphere, Gram-Schmidt (utils.c:91-113), rows (z, x, y), the
// first two swapped when the determinant is negative
void random_orientation_h(Uniforms &u, double *m) {
	double x[3], y[3], z[3];
	sphere_point(u, x);
	sphere_point(u, y);
	sphere_point(u, z);
	const double x_n2 = dot3h(x, x);
	const double b1 = dot3h(y, x) / x_n2;
	for(int k = 0; k < 3; k++) y[k] -= b1 * x[k];
	const double z_x = dot3h(z, x), z_y = dot3h(z, y), y_n2 = dot3h(y, y);
	const double c1 = z_x / x_n2, c2 = z_y / y_n2;
	for(int k = 0; k < 3; k++) z[k] -= c1 * x[k] + c2 * y[k];
	normalise_h(x);
	normalise_h(y);
	normalise_h(z);
	for(int k = 0; k < 3; k++) {
		m[k] = z[k];
		m[3 + k] = x[k];
		m[6 + k] = y[k];
	}
	double det = 0.;
	det += m[0] * m[4] * m[8];
	det += m[1] * m[5] * m[6];
	det += m[2] * m[3] * m[7];
	det -= m[2] * m[4] * m[6];
	det -= m[0] * m[5] * m[7];
	det -= m[1] * m[3] * m[8];
	if(det < 0)
		for(int k = 0; k < 3; k++) {
			m[k] = x[k];
			m[3 + k] = z[k];
		}
}

int gc_ready(pp_handle *h, const char *who) {
	if(h->slab_world > 1) return fail(h, "%s: insertion / removal runs on a single engine", who);
	if(h->prm.ensemble != PP_NVT || h->prm.Lx_move) return fail(h, "%s: the engine must be created with Ensemble = 0 and without Lx_move", who);
	if(h->cap_N == 0) return fail(h, "%s: upload a configuration first (pp_reserve before it makes room for insertions)", who);
	CUDA_OK(h, cudaSetDevice(h->device));
	return 0;
}

// number of successes of n Bernoulli(0.01) draws, by inversion of one uniform per block of 50000 (0.99^n underflows beyond)
int binomial_one_percent(int n, Uniforms &u) {
	int k = 0;
	while(n > 0) {
		const int m = n < 50000 ? n : 50000;
		const double p = 0.01, q = 0.99, x = u.next();
		double pk = pow(q, (double) m), cdf = pk;
		int j = 0;
		while(x >= cdf && j < m) {
			pk *= (double) (m - j) / (double) (j + 1) * (p / q);
			cdf += pk;
			j++;
		}
		k += j;
		n -= m;
	}
	return k;
}

// MC_energy of one particle (MC.c:272-309) as a bond count: id >= 0 an existing particle, id < 0 the ghost in d_stage_r
int probe(pp_handle *h, int id, int *bonds, int *overlap) {
	*bonds = 0;
	*overlap = 0;
	if(h->N == 0) return 0;
	if(ensure_cells_valid(h)) return 1;
	const SimConst sc = make_const(h, h->box, h->shift);
	if(h->d_gc_out == nullptr) CUDA_OK(h, cudaMalloc(&h->d_gc_out, 2 * sizeof(int)));
	int *d_out = h->d_gc_out;
	prof_begin(h);
	k_probe_particle<<<1, 32, 0, h->stream>>>(sc, id, h->d_stage_r, h->cb.cell_start, h->pos[h->cp], h->ori[h->co], h->slot_of, d_out);
	prof_end(h, PP_K_PROPOSALS, 1);
	int out[2];
	CUDA_OK(h, cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	*bonds = out[0];
	*overlap = out[1];
	return 0;
}

int stage_ghost(pp_handle *h, const double *r, const double *m) {
	double g[12];
	memcpy(g, r, sizeof(double) * 3);
	memcpy(g + 3, m, sizeof(double) * 9);
	CUDA_OK(h, cudaMemcpyAsync(h->d_stage_r, g, sizeof(g), cudaMemcpyHostToDevice, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream)); // (g is on the stack)
	return 0;
}

// one MC_add_remove (MC.c:342-448) with the uniforms of `u`, in the reference's draw order
int add_remove(pp_handle *h, double z, int n_min, int n_max, Uniforms &u, pp_gc_result *res) {
	memset(res, 0, sizeof(*res));
	res->kind = -1;
	res->index = -1;
	const double T = h->prm.temperature;
	if(u.next() < 0.5) { // try to add a particle
		if(h->N == n_max) {
			res->N = h->N;
			res->n_uniforms = u.used;
			return 0; // (do_GC stops the run here, MC.c:346; the host decides)
		}
		if((size_t) h->N + 1 > h->cap_N || (size_t) h->N + 1 > h->cap_ids) return fail(h, "insertion beyond the reserved capacity: call pp_reserve(n_max) before pp_upload");
		res->kind = PP_ADD;
		h->tries[PP_ADD]++;
		double r[3], m[9];
		for(int k = 0; k < 3; k++) r[k] = u.next() * h->box[k];
		random_orientation_h(u, m);
		if(stage_ghost(h, r, m)) return 1;
		int bonds, overlap;
		if(probe(h, -1, &bonds, &overlap)) return 1;
		if(overlap) bonds = 0; // MC_energy returns 0 at the first overlap (MC.c:296-299)
		const double delta_E = -(double) bonds;
		const double acc = exp(-delta_E / T) * z * h->V / (h->N + 1.);
		res->bonds = bonds;
		res->overlap = overlap;
		res->index = h->N;
		if(!overlap && u.next() < acc) { // the uniform is drawn only without an overlap (MC.c:364)
			k_insert_particle<<<1, 32, 0, h->stream>>>(h->N, h->d_stage_r, h->pos[h->cp], h->ori[h->co], h->pid[h->co], h->slot_of);
			CUDA_OK(h, cudaGetLastError());
			h->N++;
			h->bonds_base += bonds; // energy += delta_E
			h->accepted[PP_ADD]++;
			h->cells_valid = false;
			h->sweep_grid = false;
			res->accepted = 1;
		}
	}
	else { // try to remove a particle
		if(h->N == n_min) {
			res->N = h->N;
			res->n_uniforms = u.used;
			return 0;
		}
		res->kind = PP_REMOVE;
		h->tries[PP_REMOVE]++;
		const int p = (int) (u.next() * h->N);
		int bonds, overlap;
		if(probe(h, p, &bonds, &overlap)) return 1;
		if(overlap) bonds = 0;
		const double delta_E = (double) bonds; // -MC_energy
		const double acc = exp(-delta_E / T) * h->N / (h->V * z);
		res->bonds = bonds;
		res->overlap = overlap;
		res->index = p;
		if(u.next() < acc) {
			k_remove_particle<<<1, 32, 0, h->stream>>>(h->N, p, h->pos[h->cp], h->ori[h->co], h->pid[h->co], h->slot_of);
			CUDA_OK(h, cudaGetLastError());
			h->N--;
			h->bonds_base -= bonds;
			h->accepted[PP_REMOVE]++;
			h->cells_valid = false;
			h->sweep_grid = false;
			res->accepted = 1;
		}
	}
	res->N = h->N;
	res->n_uniforms = u.used;
	return 0;
}

} // namespace

extern "C" {

int pp_reserve(pp_handle *h, int n_max) {
	if(h == nullptr) return 1;
	if(n_max < 0) return fail(h, "pp_reserve: negative capacity");
	h->reserve_N = n_max;
	return 0;
}

int pp_particle_bonds(pp_handle *h, int id, int *bonds, int *overlap) {
	if(h == nullptr || bonds == nullptr || overlap == nullptr) return 1;
	if(gc_ready(h, "pp_particle_bonds")) return 1;
	if(id < 0 || id >= h->N) return fail(h, "pp_particle_bonds: particle index %d out of range", id);
	return probe(h, id, bonds, overlap);
}

int pp_probe_bonds(pp_handle *h, const double *r, const double *orient, int *bonds, int *overlap) {
	if(h == nullptr || r == nullptr || orient == nullptr || bonds == nullptr || overlap == nullptr) return 1;
	if(gc_ready(h, "pp_probe_bonds")) return 1;
	if(stage_ghost(h, r, orient)) return 1;
	return probe(h, -1, bonds, overlap);
}

int pp_add_remove(pp_handle *h, double activity, int n_min, int n_max, const double *u, int n_u, pp_gc_result *res) {
	if(h == nullptr || u == nullptr || res == nullptr) return 1;
	if(gc_ready(h, "pp_add_remove")) return 1;
	Uniforms src = {u, n_u, 0, h, false};
	if(add_remove(h, activity, n_min, n_max, src, res)) return 1;
	if(src.exhausted) return fail(h, "pp_add_remove: the move needed more than the %d uniforms supplied", n_u);
	return 0;
}

int pp_gc_step(pp_handle *h, double activity, int n_min, int n_max, long long *sus_hist, int *n_trials, int *full) {
	if(h == nullptr) return 1;
	if(gc_ready(h, "pp_gc_step")) return 1;
	// do_GC / do_SUS (MC.c:31-58): N_max iterations per step, each an insertion / removal trial with probability 0.01;
	// here the Binomial(N_max, 0.01) trials of a step run at the sweep boundary
	Uniforms src = {nullptr, 0, 0, h, false};
	const int k = binomial_one_percent(n_max, src);
	if(full) *full = 0;
	for(int t = 0; t < k; t++) {
		pp_gc_result res;
		if(add_remove(h, activity, n_min, n_max, src, &res)) return 1;
		if(res.kind == -1 && res.N == n_max && full) *full = 1; // an insertion into a full system (fatal under GC, MC.c:346)
		if(sus_hist != nullptr && h->N >= n_min && h->N <= n_max) sus_hist[h->N - n_min]++; // MC.c:51
	}
	if(n_trials) *n_trials = k;
	return 0;
}

} // extern "C"
