// pp_slab_api.cu -- host side of the x-slab decomposition: one engine per slab, each owning whole
// pairs of x layers of the global cell grid plus one halo layer per side.  The reference has no decomposition (single
// thread, main.c:26-143); this is the scale-out axis of the engine.
//
// Every engine runs the same sequence of PHASES -- a rebuild at the start of a shift period, then per sweep a block of
// the four even-x colour passes (B0) and a block of the four odd-x passes (B1), plus box moves at sweep boundaries --
// and counts the phases it has completed.  The data path has no transfer step: the sweep kernels mirror accepted moves
// of the boundary layers straight into the neighbour's halo copy (HaloMirror).  What is left to order is
//   B0       reads the left halo (written by the left neighbour's B1) and writes the left neighbour's right halo (read by
//            its B1)                                                    -> the left neighbour has completed my phase count
//   B1       the same with the right neighbour and its B0               -> the right neighbour has completed my phase count
//   rebuild  re-sorts the arrays both neighbours mirror into            -> both
//   box move rescales / counts the halo layers the neighbours write     -> both
// Between processes (one engine per GPU, CUDA IPC) the phase count is a word in the neighbour's memory: release store
// after the phase, acquire spin in a one-thread kernel before it.  Inside one process (pp_group) the same orderings are
// CUDA events, which also lets several slabs share one GPU without spinning kernels.
#include "pp_internal.h"

using namespace pp;
using namespace ppi;

namespace {

unsigned long long *ctrl_words(char *arena, const pp_peer_info &L) {
	return reinterpret_cast<unsigned long long *>(arena + L.off_ctrl);
}
int *ctrl_ints(char *arena, const pp_peer_info &L) {
	return reinterpret_cast<int *>(arena + L.off_ctrl + 64);
}
// control block: u64 [0] phases completed by the left neighbour, [1] by the right neighbour, [2] / [3] migration epochs
// from the left / right; int (at byte 64) [0] / [1] migration record counts from the left / right, [2] first slot of the
// left neighbour's right halo

int draw_pending_shift(pp_handle *h) {
	const uint64_t epoch = (uint64_t) (h->step / h->prm.shift_every);
	sweep_shift(h->prm.seed, epoch, h->box, h->n_side, h->pending_shift);
	return 0;
}

int classify(pp_handle *h) {
	if(ensure_cells_valid(h)) return 1;
	draw_pending_shift(h);
	const SimConst sc_new = make_const(h, h->box, h->pending_shift);
	const int tp = (h->cp + 2) % 3, to = h->co ^ 1;
	prof_begin(h);
	const int nl = launch_slab_classify(sc_new, h->layer_slots[1], h->layer_slots[4], h->x0, h->x1 - h->x0, view(h, h->cp, h->co), view(h, tp, to),
			h->outbox[0], h->outbox[1], h->cap_records, h->d_slab_counts, h->stream);
	prof_end(h, PP_K_CELLS, nl);
	return 0;
}

// unpack the received records behind the kept particles and sort the merged set on the new grid
int merge_and_sort(pp_handle *h) {
	const int tp = (h->cp + 2) % 3, to = h->co ^ 1;
	const int n_new = h->keep_count + h->in_count[0] + h->in_count[1];
	if((size_t) n_new > h->cap_N) return fail(h, "slab: %d local particles exceed the capacity %zu", n_new, h->cap_N);
	launch_slab_unpack(h->in_count[0], h->inbox[0], h->keep_count, view(h, tp, to), h->stream);
	launch_slab_unpack(h->in_count[1], h->inbox[1], h->keep_count + h->in_count[0], view(h, tp, to), h->stream);
	// the previous current buffers are free to be the destination
	h->N = n_new;
	if(sort_into(h, h->pending_shift, view(h, tp, to), view(h, h->cp, h->co))) return 1;
	h->cells_valid = true;
	h->sweep_grid = true;
	h->slab_pending = false;
	h->rebuilt_step = h->step;
	if(slab_after_sort(h)) return 1;
	if(h->N != n_new) return fail(h, "slab: %d of %d particles fell outside the held layers after the exchange", n_new - h->N, n_new);
	return 0;
}

int read_counts(pp_handle *h, int n) {
	int *c = h->h_slab_ints;
	CUDA_OK(h, cudaMemcpyAsync(c, h->d_slab_counts, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	if(c[3]) return fail(h, "%d particles moved by more than one x layer at a grid shift", c[3]);
	if(c[1] > h->cap_records || c[2] > h->cap_records) return fail(h, "slab outbox overflow (%d, %d records, capacity %d)", c[1], c[2], h->cap_records);
	h->keep_count = c[0];
	h->out_count[0] = c[1];
	h->out_count[1] = c[2];
	return 0;
}

} // namespace

namespace ppi {

bool slab_period_starts(const pp_handle *h) {
	if(!(h->sweep_grid && h->cells_valid)) return true;
	return h->step % h->prm.shift_every == 0 && h->rebuilt_step != h->step;
}

int slab_wait(pp_handle *h, int sides) {
	if(!h->peer_connected) return fail(h, "slab phase without a peer connection");
	if(h->phase == 0) return 0; // nothing has run since the connection was made
	if(h->peer_ipc) {
		unsigned long long *mine = ctrl_words(h->arena, h->layout);
		launch_phase_wait((sides & kSideLeft) ? mine + 0 : nullptr, (sides & kSideRight) ? mine + 1 : nullptr, (unsigned long long) h->phase,
				h->d_peer_status, h->stream);
		h->launches[PP_K_HALO]++;
		h->total_launches++;
	}
	else {
		// the neighbour recorded the event of its phase count `phase` on ev_phase[phase & 1] (the group driver issues the
		// phases of all its engines in lock step, so that record has been issued)
		if(sides & kSideLeft) CUDA_OK(h, cudaStreamWaitEvent(h->stream, h->peer_handle[0]->ev_phase[h->phase & 1], 0));
		if((sides & kSideRight) && !(h->peer_same && (sides & kSideLeft))) CUDA_OK(h, cudaStreamWaitEvent(h->stream, h->peer_handle[1]->ev_phase[h->phase & 1], 0));
	}
	return 0;
}

int slab_signal(pp_handle *h, bool layout) {
	h->phase++;
	if(h->peer_ipc) {
		// I am the right neighbour of my left neighbour: its word [1]; the left neighbour of my right neighbour: its word [0]
		unsigned long long *to_left = ctrl_words(h->peer_arena[0], h->peer_layout[0]) + 1;
		unsigned long long *to_right = ctrl_words(h->peer_arena[1], h->peer_layout[1]) + 0;
		int *layout_dst = layout ? ctrl_ints(h->peer_arena[1], h->peer_layout[1]) + 2 : nullptr;
		launch_phase_signal(to_left, to_right, (unsigned long long) h->phase, layout_dst, h->layer_slots[4], h->d_peer_status, h->stream);
		h->launches[PP_K_HALO]++;
		h->total_launches++;
	}
	else CUDA_OK(h, cudaEventRecord(h->ev_phase[h->phase & 1], h->stream));
	return 0;
}

// where this engine's accepted boundary moves go: the neighbours' CURRENT buffers (every engine of a decomposition goes
// through the same sequence of buffer switches, so the indices are the engine's own)
int slab_build_mirror(pp_handle *h) {
	HaloMirror &hm = h->hm;
	memset(&hm, 0, sizeof(hm));
	if(!h->peer_connected) return 0;
	for(int side = 0; side < 2; side++) {
		char *a = h->peer_arena[side];
		const pp_peer_info &L = h->peer_layout[side];
		double4 *pos = reinterpret_cast<double4 *>(a + L.off_pos[h->cp]);
		double *ori = reinterpret_cast<double *>(a + L.off_ori[h->co]);
		uint4 *posf = reinterpret_cast<uint4 *>(a + L.off_posf);
		if(side == 0) {
			hm.pos_l = pos;
			hm.ori_l = ori;
			hm.posf_l = posf;
		}
		else {
			hm.pos_r = pos;
			hm.ori_r = ori;
			hm.posf_r = posf;
		}
	}
	hm.ls1 = h->layer_slots[1];
	hm.ls2 = h->layer_slots[2];
	hm.ls3 = h->layer_slots[3];
	if(h->peer_ipc) hm.off_l_ptr = ctrl_ints(h->arena, h->layout) + 2; // published by the left neighbour with its phase number
	else {
		const pp_handle *left = h->peer_handle[0];
		if(left->cp != h->cp || left->co != h->co) return fail(h, "slab engines out of step (buffers %d/%d here, %d/%d on the left)", h->cp, h->co, left->cp, left->co);
		hm.off_l = left->layer_slots[4];
	}
	return 0;
}

// start of a shift period, first half: wait until nobody writes into this engine's arrays any more, classify the owned
// particles on the new grid and store the outgoing records (migrants + fresh halo layers) into the neighbours' mailboxes
int slab_rebuild_begin(pp_handle *h) {
	CUDA_OK(h, cudaSetDevice(h->device));
	if(slab_wait(h, kSideLeft | kSideRight)) return 1;
	if(classify(h)) return 1;
	prof_begin(h);
	int nl = 0;
	const unsigned long long me = ++h->mig_epoch;
	// to the left neighbour: its from-right mailbox (1), flag [3], count [1]; to the right neighbour: its from-left mailbox (0), flag [2], count [0]
	for(int side = 0; side < 2; side++) {
		char *a = h->peer_arena[side];
		const pp_peer_info &L = h->peer_layout[side];
		const int box = side == 0 ? 1 : 0;
		const int cap = (int) (L.cap_records < h->cap_records ? L.cap_records : h->cap_records);
		nl += launch_records_push(h->outbox[side], h->d_slab_counts, 1 + side, cap, reinterpret_cast<double *>(a + L.off_mig[box]), ctrl_ints(a, L) + box,
				ctrl_words(a, L) + 2 + box, me, h->d_push_done, h->stream);
	}
	prof_end(h, PP_K_HALO, nl);
	if(!h->peer_ipc) CUDA_OK(h, cudaEventRecord(h->ev_mig, h->stream));
	h->slab_pending = true;
	return 0;
}

// second half: take in what the neighbours stored here, re-sort, read the new layout back (the one host synchronisation
// of a shift period)
int slab_rebuild_end(pp_handle *h) {
	CUDA_OK(h, cudaSetDevice(h->device));
	if(!h->slab_pending) return fail(h, "slab_rebuild_end without slab_rebuild_begin");
	if(!h->peer_ipc) { // the neighbours' pushes are complete (their flags are set: the pulls below do not spin)
		CUDA_OK(h, cudaStreamWaitEvent(h->stream, h->peer_handle[0]->ev_mig, 0));
		if(!h->peer_same) CUDA_OK(h, cudaStreamWaitEvent(h->stream, h->peer_handle[1]->ev_mig, 0));
	}
	prof_begin(h);
	int nl = 0;
	for(int side = 0; side < 2; side++)
		nl += launch_records_pull(reinterpret_cast<const double *>(h->arena + h->layout.off_mig[side]), ctrl_ints(h->arena, h->layout) + side,
				ctrl_words(h->arena, h->layout) + 2 + side, h->mig_epoch, h->cap_records, h->inbox[side], h->d_slab_counts + 4 + side, h->d_peer_status,
				h->stream);
	prof_end(h, PP_K_HALO, nl);
	if(read_counts(h, 6)) return 1;
	h->in_count[0] = h->h_slab_ints[4];
	h->in_count[1] = h->h_slab_ints[5];
	if(merge_and_sort(h)) return 1;
	return 0;
}

int slab_block(pp_handle *h, int cx) {
	CUDA_OK(h, cudaSetDevice(h->device));
	if(!h->sweep_grid || !h->cells_valid) return fail(h, "slab pass block before the rebuild of its shift period");
	const SimConst sc = make_const(h, h->box, h->shift);
	if(cx == 0) { // (the proposals do not depend on the neighbours: before the wait)
		prof_begin(h);
		const int nr = launch_proposals_for_sweep(sc, h->cb, h->d_proposals, h->stream);
		prof_end(h, PP_K_RNG, nr);
	}
	if(h->peer_connected && slab_wait(h, cx == 0 ? kSideLeft : kSideRight)) return 1;
	if(!h->peer_connected) { // send / recv transport: the driver has just copied the neighbour's layer into the halo this block reads
		const int b = cx == 0 ? h->layer_slots[0] : h->layer_slots[4], e = cx == 0 ? h->layer_slots[1] : h->layer_slots[5];
		const int nf = launch_shadow_slots(h->cb.posf, b, e, h->stream);
		h->launches[PP_K_HALO] += nf;
		h->total_launches += nf;
	}
	prof_begin(h);
	const int nl = launch_passes(h, sc, 4 * cx, 4);
	prof_end(h, PP_K_SWEEP, nl);
	CUDA_OK(h, cudaGetLastError());
	if(h->peer_connected && slab_signal(h, false)) return 1;
	return 0;
}

} // namespace ppi

// ---------------------------------------------------------------------------------------------------
extern "C" {

int pp_slab_setup(pp_handle *h, int rank, int world) {
	if(h == nullptr) return 1;
	if(world < 1 || rank < 0 || rank >= world) return fail(h, "bad slab rank %d / world %d", rank, world);
	h->slab_rank = rank;
	h->slab_world = world;
	return 0;
}

int pp_slab_range(pp_handle *h, int *x0, int *x1, int *nx, int *n_local) {
	if(h == nullptr) return 1;
	if(x0) *x0 = (h->slab_world > 1) ? h->x0 : 0;
	if(x1) *x1 = (h->slab_world > 1) ? h->x1 : h->n_side[0];
	if(nx) *nx = h->n_side[0];
	if(n_local) *n_local = h->N;
	return 0;
}

// ---- transport by the driver (send / recv of device buffers): the pass blocks run without the halo mirror ---------------
int pp_slab_sweep_begin(pp_handle *h, int *exchange) {
	if(h == nullptr || exchange == nullptr) return 1;
	if(h->slab_world <= 1) return fail(h, "pp_slab_sweep_begin without a slab decomposition");
	CUDA_OK(h, cudaSetDevice(h->device));
	*exchange = 0;
	if(!slab_period_starts(h)) return 0;
	// new shift period: same stream as the single-engine path, so every engine draws the same shift
	if(classify(h)) return 1;
	if(read_counts(h, 4)) return 1;
	h->in_count[0] = h->in_count[1] = 0;
	h->slab_pending = true;
	*exchange = 1;
	return 0;
}

int pp_slab_outbox(pp_handle *h, int side, void **dev, int *count) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(dev) *dev = h->outbox[side];
	if(count) *count = h->out_count[side];
	return 0;
}

int pp_slab_inbox(pp_handle *h, int side, int count, void **dev) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(count < 0 || count > h->cap_records) return fail(h, "slab inbox overflow (%d records, capacity %d)", count, h->cap_records);
	h->in_count[side] = count;
	if(dev) *dev = h->inbox[side];
	return 0;
}

int pp_slab_sweep_rebuild(pp_handle *h) {
	if(h == nullptr) return 1;
	if(!h->slab_pending) return fail(h, "pp_slab_sweep_rebuild without a pending exchange");
	return merge_and_sort(h);
}

int pp_slab_pass_block(pp_handle *h, int cx) {
	if(h == nullptr || cx < 0 || cx > 1) return 1;
	if(h->slab_world <= 1) return fail(h, "pp_slab_pass_block without a slab decomposition");
	if(h->peer_connected) return fail(h, "pp_slab_pass_block is the call of the send / recv transport; with a peer connection use pp_slab_sweeps");
	return slab_block(h, cx);
}

int pp_slab_halo(pp_handle *h, int side, int recv, void **pos, void **ori, void **posf, int *count) {
	if(h == nullptr || side < 0 || side > 1) return 1;
	if(h->slab_world <= 1 || !h->cells_valid) return fail(h, "pp_slab_halo: no valid slab layout");
	// send: my first (side 0) / last (side 1) owned layer; recv: my left (side 0) / right (side 1) halo layer
	int b, e;
	if(recv) {
		b = side == 0 ? h->layer_slots[0] : h->layer_slots[4];
		e = side == 0 ? h->layer_slots[1] : h->layer_slots[5];
	}
	else {
		b = side == 0 ? h->layer_slots[1] : h->layer_slots[3];
		e = side == 0 ? h->layer_slots[2] : h->layer_slots[4];
	}
	if(pos) *pos = h->pos[h->cp] + b;
	if(ori) *ori = h->ori[h->co] + 9ll * b;
	if(posf) *posf = h->cb.posf + b;
	if(count) *count = e - b;
	return 0;
}

int pp_slab_sweep_end(pp_handle *h) {
	if(h == nullptr) return 1;
	h->step++;
	return 0;
}

// ---- stores into the neighbour's memory ---------------------------------------------------------------------------------
int pp_slab_peer_export(pp_handle *h, pp_peer_info *info) {
	if(h == nullptr || info == nullptr) return 1;
	if(h->slab_world <= 1 || !h->cells_valid) return fail(h, "pp_slab_peer_export: upload a configuration in slab mode first");
	CUDA_OK(h, cudaSetDevice(h->device));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "pp_peer_info carries the IPC handle as 64 bytes");
	cudaIpcMemHandle_t mh;
	CUDA_OK(h, cudaIpcGetMemHandle(&mh, h->arena));
	*info = h->layout;
	memcpy(info->ipc, &mh, 64);
	return 0;
}

int pp_slab_peer_connect(pp_handle *h, const pp_peer_info *left, const pp_peer_info *right) {
	if(h == nullptr || left == nullptr || right == nullptr) return 1;
	if(h->slab_world <= 1 || !h->cells_valid) return fail(h, "pp_slab_peer_connect: upload a configuration in slab mode first");
	CUDA_OK(h, cudaSetDevice(h->device));
	peer_release(h);
	cudaIpcMemHandle_t l, r;
	memcpy(&l, left->ipc, 64);
	memcpy(&r, right->ipc, 64);
	h->peer_same = memcmp(&l, &r, 64) == 0; // two engines: both neighbours are the same one
	CUDA_OK(h, cudaIpcOpenMemHandle((void **) &h->peer_arena[0], l, cudaIpcMemLazyEnablePeerAccess));
	if(h->peer_same) h->peer_arena[1] = h->peer_arena[0];
	else CUDA_OK(h, cudaIpcOpenMemHandle((void **) &h->peer_arena[1], r, cudaIpcMemLazyEnablePeerAccess));
	h->peer_layout[0] = *left;
	h->peer_layout[1] = *right;
	h->peer_ipc = true;
	h->peer_connected = true;
	h->phase = 0;
	h->mig_epoch = 0;
	CUDA_OK(h, cudaMemset(h->d_peer_status, 0, sizeof(int)));
	// my right neighbour needs the first slot of my right halo for its mirror stores: the first signal publishes it
	if(slab_build_mirror(h)) return 1;
	if(slab_signal(h, true)) return 1; // phase 1 = "connected, layout published"
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

int pp_slab_peer_close(pp_handle *h) {
	if(h == nullptr) return 1;
	if(h->stream) cudaStreamSynchronize(h->stream);
	peer_release(h);
	return 0;
}

int pp_slab_migrate(pp_handle *h, int *migrated) {
	if(h == nullptr || migrated == nullptr) return 1;
	if(!h->peer_connected || !h->peer_ipc) return fail(h, "pp_slab_migrate needs the peer connection (pp_slab_peer_connect)");
	*migrated = 0;
	if(!slab_period_starts(h)) return 0;
	if(slab_rebuild_begin(h)) return 1;
	if(slab_rebuild_end(h)) return 1;
	if(slab_build_mirror(h)) return 1;
	if(slab_signal(h, true)) return 1;
	*migrated = 1;
	return 0;
}

int pp_slab_sweeps(pp_handle *h, int max_sweeps, int *done) {
	if(h == nullptr || done == nullptr) return 1;
	*done = 0;
	if(!h->peer_connected || !h->peer_ipc) return fail(h, "pp_slab_sweeps needs the peer connection (pp_slab_peer_connect)");
	while(*done < max_sweeps) {
		if(slab_period_starts(h)) { // a new shift period starts here: pp_slab_migrate first
			if(*done == 0) return fail(h, "pp_slab_sweeps at the start of a shift period: call pp_slab_migrate first");
			break;
		}
		for(int cx = 0; cx < 2; cx++)
			if(slab_block(h, cx)) return 1;
		h->step++;
		(*done)++;
	}
	return 0;
}

int pp_slab_box_trial(pp_handle *h, int kind, const double *u, long long bonds_running_total, long long *bonds_local, int *overlap_local, int *status) {
	if(h == nullptr || u == nullptr || status == nullptr) return 1;
	CUDA_OK(h, cudaSetDevice(h->device));
	if(fetch_counters(h)) return 1;
	if(box_propose(h, kind, u, status)) return 1;
	if(*status < 0) return 0;
	h->bm.bonds_old_total = bonds_running_total;
	// the halo layers are about to be rescaled and counted: the neighbours must be done writing them
	if(h->peer_connected && slab_wait(h, kSideLeft | kSideRight)) return 1;
	long long b = 0;
	int ov = 0;
	if(box_trial_local(h, &b, &ov, status)) return 1;
	h->bm.bonds_local_new = b;
	if(bonds_local) *bonds_local = b;
	if(overlap_local) *overlap_local = ov;
	return 0;
}

int pp_slab_box_finish(pp_handle *h, long long bonds_total, int overlap_any, int *accepted) {
	if(h == nullptr) return 1;
	CUDA_OK(h, cudaSetDevice(h->device));
	const bool acc = box_decide(h, bonds_total, overlap_any, h->bm.bonds_old_total, h->N_global);
	if(box_finish(h, acc, h->bm.bonds_local_new)) return 1;
	if(h->peer_connected && slab_signal(h, false)) return 1; // the box move is a phase of its own
	if(accepted) *accepted = acc ? 1 : 0;
	return 0;
}

int pp_slab_download(pp_handle *h, int *n_owned, int *ids, double *r, double *orient) {
	if(h == nullptr) return 1;
	if(h->slab_world <= 1) return fail(h, "pp_slab_download without a slab decomposition");
	CUDA_OK(h, cudaSetDevice(h->device));
	if(ensure_cells_valid(h)) return 1;
	const int b = h->layer_slots[1], e = h->layer_slots[4], n = e - b;
	if(n_owned) *n_owned = n;
	if(ids == nullptr || n == 0) return 0;
	int *d_ids = h->cb.cell_of; // scratch
	launch_slab_download(b, e, view(h, h->cp, h->co), d_ids, h->d_stage_r, h->d_stage_o, h->stream);
	CUDA_OK(h, cudaMemcpyAsync(ids, d_ids, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(r, h->d_stage_r, sizeof(double) * 3 * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaMemcpyAsync(orient, h->d_stage_o, sizeof(double) * 9 * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
	CUDA_OK(h, cudaStreamSynchronize(h->stream));
	return 0;
}

} // extern "C"
