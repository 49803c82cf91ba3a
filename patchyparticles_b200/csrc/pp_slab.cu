// pp_slab.cu -- kernels of the x-slab decomposition.
//
// Each engine (one per GPU) holds the particles of a contiguous range of x layers of the global
// cell grid plus one halo layer on each side.  Cells are stored x-major, so a layer is a contiguous
// range of particle slots: the in-period halo refresh is a plain copy of three array slices
// (positions, orientations, FP32 shadow) and needs no kernel.  The kernels here serve the start of
// a shift period, when the grid origin moves and particles change owner:
//   k_slab_classify  walks the owned particles with the NEW grid and splits them into "kept" (new
//                    layer still held here, as owner or as halo copy), "to the left neighbour"
//                    (its new last owned layer or its new right halo) and "to the right
//                    neighbour"; boundary-layer particles and migrants are both kept and sent;
//   k_slab_unpack    appends the records received from a neighbour to the kept particles;
//   k_slab_download  compacts the owned particles for output.
// A record is 13 doubles: x, y, z, id, orientation[9].
//
// In-period halo refresh: none.  The sweep kernels mirror every accepted move of a boundary-layer particle into the
// neighbour's halo copy themselves (HaloMirror, pp_common.cuh); neighbouring engines only exchange phase numbers
// (k_phase_signal / k_phase_wait below).  Migration records travel through mailboxes in the neighbour's memory
// (k_records_push / k_records_pull).
#include "pp_kernels.cuh"

namespace pp {

__global__ void __launch_bounds__(256) k_slab_classify(const __grid_constant__ SimConst sc_new, int slot_begin, int slot_end, int x0, int n_owned_layers,
		const double4 *__restrict__ pos, const double *__restrict__ ori, const int *__restrict__ pid, double4 *__restrict__ keep_pos,
		double *__restrict__ keep_ori, int *__restrict__ keep_pid, double *__restrict__ out_left, double *__restrict__ out_right, int cap_records,
		int *__restrict__ counts /* keep, left, right, error */) {
	const int s = slot_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= slot_end) return;
	const double4 p = pos[s];
	const int nx = sc_new.n_side[0];
	int rel = cell_coord(p.x, sc_new.shift[0], sc_new.inv_box[0], nx) - x0; // new global layer relative to the first owned one
	rel += (rel < 0) ? nx : 0;
	// kept: everything that lies in a layer held here under the new grid -- the owned layers and the two halo
	// layers (a particle migrating to a neighbour stays behind as a halo copy)
	const bool keep = (rel <= n_owned_layers) || (rel == nx - 1);
	const bool left = (rel == 0) || (rel == nx - 1);
	const bool right = (rel == n_owned_layers - 1) || (rel == n_owned_layers);
	if(!keep && !left && !right) {
		atomicAdd(counts + 3, 1); // moved by more than one layer: impossible while the shift changes by less than a cell
		return;
	}
	double m[9];
#pragma unroll
	for(int k = 0; k < 9; k++) m[k] = ori[9ll * s + k];
	if(keep) {
		const int d = atomicAdd(counts + 0, 1);
		keep_pos[d] = p;
		keep_pid[d] = pid[s];
#pragma unroll
		for(int k = 0; k < 9; k++) keep_ori[9ll * d + k] = m[k];
	}
	if(left) {
		const int d = atomicAdd(counts + 1, 1);
		if(d < cap_records) {
			double *rec = out_left + 13ll * d;
			rec[0] = p.x;
			rec[1] = p.y;
			rec[2] = p.z;
			rec[3] = p.w;
#pragma unroll
			for(int k = 0; k < 9; k++) rec[4 + k] = m[k];
		}
	}
	if(right) {
		const int d = atomicAdd(counts + 2, 1);
		if(d < cap_records) {
			double *rec = out_right + 13ll * d;
			rec[0] = p.x;
			rec[1] = p.y;
			rec[2] = p.z;
			rec[3] = p.w;
#pragma unroll
			for(int k = 0; k < 9; k++) rec[4 + k] = m[k];
		}
	}
}

__global__ void __launch_bounds__(256) k_slab_unpack(int n, const double *__restrict__ records, int offset, double4 *__restrict__ pos,
		double *__restrict__ ori, int *__restrict__ pid) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= n) return;
	const double *rec = records + 13ll * i;
	const int d = offset + i;
	pos[d] = make_double4(rec[0], rec[1], rec[2], rec[3]);
	pid[d] = (int) rec[3];
#pragma unroll
	for(int k = 0; k < 9; k++) ori[9ll * d + k] = rec[4 + k];
}

__global__ void __launch_bounds__(256) k_slab_download(int slot_begin, int slot_end, const double4 *__restrict__ pos, const double *__restrict__ ori,
		const int *__restrict__ pid, int *__restrict__ ids, double *__restrict__ r, double *__restrict__ orient) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int s = slot_begin + i;
	if(s >= slot_end) return;
	const double4 p = pos[s];
	ids[i] = pid[s];
	r[3ll * i] = p.x;
	r[3ll * i + 1] = p.y;
	r[3ll * i + 2] = p.z;
#pragma unroll
	for(int k = 0; k < 9; k++) orient[9ll * i + k] = ori[9ll * s + k];
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// Phase numbers between neighbouring engines.  After a rebuild, a block of four colour passes or a box move an engine
// stores the number of phases it has completed into the control words of both neighbours (release, system scope; the
// kernels before it on the stream, with their stores into the neighbour's halo, are complete).  Before a phase that
// reads what a neighbour wrote, or writes where a neighbour may still read, a one-thread kernel spins on the engine's
// own control word (acquire) until the neighbour has completed as many phases as the engine itself.
__global__ void k_phase_signal(unsigned long long *peer_a, unsigned long long *peer_b, unsigned long long value, int *layout_dst, int layout_val,
		const int *status) {
	if(*status) return; // a wait timed out earlier: do not let the neighbours run on
	if(layout_dst) *layout_dst = layout_val; // first slot of this engine's right halo, for the right neighbour's mirror stores
	__threadfence_system();
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_a), "l"(value) : "memory");
	if(peer_b != peer_a) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_b), "l"(value) : "memory");
}

__global__ void k_phase_wait(const unsigned long long *flag_a, const unsigned long long *flag_b, unsigned long long need, int *status) {
	const unsigned long long t0 = global_timer_ns();
	for(int k = 0; k < 2; k++) {
		const unsigned long long *flag = k == 0 ? flag_a : flag_b;
		if(flag == nullptr) continue;
		for(;;) {
			unsigned long long v;
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
			if(v >= need) break;
			if(global_timer_ns() - t0 > 20000000000ull) { // 20 s: the neighbour is gone; report instead of hanging
				atomicExch(status, 1);
				return;
			}
			__nanosleep(100);
		}
	}
}

// migration records at a grid shift, over the same peer mailboxes: the sender's count travels with the data
__global__ void __launch_bounds__(256) k_records_push(const double *__restrict__ outbox, const int *__restrict__ counts, int which, int cap,
		double *mbox /* peer memory */, int *peer_count, unsigned long long *flag /* peer memory */, unsigned long long epoch, unsigned int *done) {
	int n = counts[which];
	n = n < cap ? n : cap; // an overflow is reported by the host from the same counts
	const long long stride = (long long) gridDim.x * blockDim.x, i0 = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	for(long long i = i0; i < 13ll * n; i += stride) mbox[i] = outbox[i];
	__threadfence_system();
	__syncthreads();
	if(threadIdx.x == 0 && atomicAdd(done, 1u) == gridDim.x - 1) {
		*done = 0;
		*peer_count = n;
		__threadfence_system();
		asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
	}
}

__global__ void __launch_bounds__(256) k_records_pull(const double *__restrict__ mbox, const int *my_count, const unsigned long long *flag,
		unsigned long long epoch, int cap, double *__restrict__ inbox, int *__restrict__ in_count, int *status) {
	if(threadIdx.x == 0) {
		const unsigned long long t0 = global_timer_ns();
		for(;;) {
			unsigned long long v;
			asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
			if(v >= epoch) break;
			if(global_timer_ns() - t0 > 20000000000ull) {
				atomicExch(status, 1);
				break;
			}
			__nanosleep(200);
		}
	}
	__syncthreads();
	int n = __ldcg(my_count);
	n = n < 0 ? 0 : (n < cap ? n : cap);
	const long long stride = (long long) gridDim.x * blockDim.x, i0 = (long long) blockIdx.x * blockDim.x + threadIdx.x;
	for(long long i = i0; i < 13ll * n; i += stride) inbox[i] = __ldcg(mbox + i);
	if(blockIdx.x == 0 && threadIdx.x == 0) *in_count = n;
}

// send / recv transport: a halo layer arrives as a copy of the owner's slices, shadows included; their .w must name the
// slots of THIS engine's arrays
__global__ void __launch_bounds__(256) k_shadow_slots(uint4 *__restrict__ posf, int begin, int end) {
	const int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
	if(s < end) posf[s].w = (unsigned) s;
}

int launch_shadow_slots(uint4 *posf, int begin, int end, cudaStream_t st) {
	if(end <= begin) return 0;
	k_shadow_slots<<<(end - begin + 255) / 256, 256, 0, st>>>(posf, begin, end);
	return 1;
}

struct GatherIdx {
	int idx[8];
};
__global__ void k_gather_ints(const int *__restrict__ src, GatherIdx g, int n, int *__restrict__ out) {
	if(threadIdx.x < n) out[threadIdx.x] = src[g.idx[threadIdx.x]];
}

int launch_gather_ints(const int *src, const int *idx, int n, int *out, cudaStream_t st) {
	GatherIdx g;
	for(int k = 0; k < 8; k++) g.idx[k] = k < n ? idx[k] : 0;
	k_gather_ints<<<1, 32, 0, st>>>(src, g, n, out);
	return 1;
}

int launch_records_push(const double *outbox, const int *counts, int which, int cap, double *peer_box, int *peer_count, unsigned long long *peer_flag,
		unsigned long long epoch, unsigned int *done, cudaStream_t st) {
	k_records_push<<<148, 256, 0, st>>>(outbox, counts, which, cap, peer_box, peer_count, peer_flag, epoch, done);
	return 1;
}

int launch_records_pull(const double *mbox, const int *my_count, const unsigned long long *flag, unsigned long long epoch, int cap, double *inbox,
		int *in_count, int *status, cudaStream_t st) {
	k_records_pull<<<148, 256, 0, st>>>(mbox, my_count, flag, epoch, cap, inbox, in_count, status);
	return 1;
}

int launch_phase_signal(unsigned long long *peer_a, unsigned long long *peer_b, unsigned long long value, int *layout_dst, int layout_val, const int *status,
		cudaStream_t st) {
	k_phase_signal<<<1, 1, 0, st>>>(peer_a, peer_b, value, layout_dst, layout_val, status);
	return 1;
}

int launch_phase_wait(const unsigned long long *flag_a, const unsigned long long *flag_b, unsigned long long need, int *status, cudaStream_t st) {
	k_phase_wait<<<1, 1, 0, st>>>(flag_a, flag_b, need, status);
	return 1;
}

int launch_slab_classify(const SimConst &sc_new, int slot_begin, int slot_end, int x0, int n_owned_layers, const StateView &cur, const StateView &keep,
		double *out_left, double *out_right, int cap_records, int *counts, cudaStream_t st) {
	cudaMemsetAsync(counts, 0, 4 * sizeof(int), st);
	const int n = slot_end - slot_begin;
	if(n <= 0) return 0;
	k_slab_classify<<<(n + 255) / 256, 256, 0, st>>>(sc_new, slot_begin, slot_end, x0, n_owned_layers, cur.pos, cur.ori, cur.pid, keep.pos, keep.ori,
			keep.pid, out_left, out_right, cap_records, counts);
	return 1;
}

int launch_slab_unpack(int n, const double *records, int offset, const StateView &dst, cudaStream_t st) {
	if(n <= 0) return 0;
	k_slab_unpack<<<(n + 255) / 256, 256, 0, st>>>(n, records, offset, dst.pos, dst.ori, dst.pid);
	return 1;
}

int launch_slab_download(int slot_begin, int slot_end, const StateView &cur, int *ids, double *r, double *orient, cudaStream_t st) {
	const int n = slot_end - slot_begin;
	if(n <= 0) return 0;
	k_slab_download<<<(n + 255) / 256, 256, 0, st>>>(slot_begin, slot_end, cur.pos, cur.ori, cur.pid, ids, r, orient);
	return 1;
}

} // namespace pp
