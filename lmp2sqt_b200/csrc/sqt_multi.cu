// sqt_multi.cu -- several GPUs behind ONE caller thread (include/sqt.h, sqt_multi_*; SURVEY.md section 8e "frame fan-out").
//
// The reference's frame loop (lmp2sqt.c:292-320) is single threaded: ReadDump, EvalSpacetimeCorr, next frame.  The drop-in
// keeps that shape: the caller fills one pinned SoA frame and pushes it; here the frame is split into G contiguous particle
// chunks (dump order, the same bounds as the one-process-per-GPU launchers use) and every chunk goes to its GPU's engine
// handle.  One worker thread per GPU issues that handle's copies and launches -- NCCL wants the ranks of one process on
// separate threads (or in group calls), and G threads also hide the launch latency of G kernel sequences.  The only
// exchanges stay the two NCCL allreduces of the single-handle path.
#include "../../include/sqt.h"
#include "sqt_internal.h"

#include <condition_variable>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>

enum { TASK_NONE = 0, TASK_CREATE, TASK_COMM, TASK_PUSH, TASK_SYNC, TASK_DESTROY, TASK_EXIT };

struct sqt_multi {
	int G, ncomp;
	sqt_params p;
	std::vector<sqt_handle *> h;
	std::vector<std::thread> th;
	std::vector<int> rc;
	std::mutex mu;
	std::condition_variable cv_go, cv_done;
	long long gen;
	int task, ndone;
	// arguments of the current task
	unsigned char comm_id[SQT_COMM_ID_BYTES];
	const double *soa;
	int n_atoms;
	double xlo, xhi;
	int stage_slot_of_push;
	// pinned whole-frame staging, double buffered; ev[s][g] = rank g's copies out of buffer s are done
	double *h_stage[2];
	int stage_cap, stage_next;
	std::vector<cudaEvent_t> ev[2];
	bool ev_pending[2];
	char err[512];
};

static char g_multi_err[512] = "";

static int mfail(sqt_multi *m, int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(m ? m->err : g_multi_err, 512, fmt, ap);
	va_end(ap);
	return code;
}

static void shard(int n, int g, int G, int *lo, int *hi)
{
	*lo = (int)(((long long)n * g) / G);
	*hi = (int)(((long long)n * (g + 1)) / G);
}

static int run_task(sqt_multi *m, int g, int task)
{
	switch (task) {
	case TASK_CREATE: {
		sqt_params q = m->p;
		q.device = m->p.device + g;
		q.n_gpus = 0;
		int rc = sqt_create(&q, &m->h[g]);
		if (rc == SQT_OK)
			rc = cudaEventCreateWithFlags(&m->ev[0][g], cudaEventDisableTiming) == cudaSuccess &&
			     cudaEventCreateWithFlags(&m->ev[1][g], cudaEventDisableTiming) == cudaSuccess ? SQT_OK : SQT_ERR_CUDA;
		return rc;
	}
	case TASK_COMM: return sqt_comm_init(m->h[g], m->comm_id, g, m->G);
	case TASK_PUSH: {
		int lo, hi;
		shard(m->n_atoms, g, m->G, &lo, &hi);
		const size_t n = (size_t)m->n_atoms;
		const double *s = m->soa;
		const int cv = (m->p.flag_V || m->p.flag_s) ? 3 : -1, cm = m->p.flag_M ? (cv >= 0 ? 6 : 3) : -1; // component order of the frame
		const double *v = cv >= 0 ? s + cv * n + lo : nullptr, *u = cm >= 0 ? s + cm * n + lo : nullptr;
		int rc = sqt_push_frame_xyz_vm(m->h[g], s + lo, s + n + lo, s + 2 * n + lo, v, v ? v + n : nullptr, v ? v + 2 * n : nullptr,
		                               u, u ? u + n : nullptr, u ? u + 2 * n : nullptr, hi - lo, m->n_atoms, m->xlo, m->xhi);
		if (rc >= 0 && m->stage_slot_of_push >= 0) {
			cudaSetDevice(m->p.device + g);
			if (cudaEventRecord(m->ev[m->stage_slot_of_push][g], sqt_internal_copy_stream(m->h[g])) != cudaSuccess) rc = SQT_ERR_CUDA;
		}
		return rc;
	}
	case TASK_SYNC: return sqt_sync(m->h[g]);
	case TASK_DESTROY:
		if (m->h[g]) {
			cudaSetDevice(m->p.device + g);
			for (int s = 0; s < 2; ++s) if (m->ev[s][g]) cudaEventDestroy(m->ev[s][g]);
			sqt_destroy(m->h[g]);
			m->h[g] = nullptr;
		}
		return SQT_OK;
	}
	return SQT_OK;
}

static void worker(sqt_multi *m, int g)
{
	long long seen = 0;
	for (;;) {
		int task;
		{
			std::unique_lock<std::mutex> lk(m->mu);
			m->cv_go.wait(lk, [&] { return m->gen != seen; });
			seen = m->gen;
			task = m->task;
		}
		if (task == TASK_EXIT) return;
		const int rc = run_task(m, g, task);
		{
			std::lock_guard<std::mutex> lk(m->mu);
			m->rc[g] = rc;
			if (++m->ndone == m->G) m->cv_done.notify_one();
		}
	}
}

// runs `task` on every rank (concurrently when G > 1); returns the first negative rc, else rank 0's
static int dispatch(sqt_multi *m, int task)
{
	if (m->G == 1) {
		m->rc[0] = run_task(m, 0, task);
	} else {
		std::unique_lock<std::mutex> lk(m->mu);
		m->task = task; m->ndone = 0; ++m->gen;
		m->cv_go.notify_all();
		m->cv_done.wait(lk, [&] { return m->ndone == m->G; });
	}
	for (int g = 0; g < m->G; ++g)
		if (m->rc[g] < 0) {
			const char *e = m->h[g] ? sqt_last_error(m->h[g]) : sqt_last_error(nullptr);
			mfail(m, m->rc[g], "GPU %d: %s", m->p.device + g, e && *e ? e : "failed");
			return m->rc[g];
		}
	return m->rc[0];
}

extern "C" const char *sqt_multi_last_error(const sqt_multi *m) { return m ? m->err : g_multi_err; }
extern "C" int sqt_multi_n_gpus(const sqt_multi *m) { return m ? m->G : 0; }
extern "C" sqt_handle *sqt_multi_handle(sqt_multi *m, int rank) { return (m && rank >= 0 && rank < m->G) ? m->h[rank] : nullptr; }

extern "C" int sqt_multi_create(const sqt_params *p, sqt_multi **out)
{
	if (!p || !out) return mfail(nullptr, SQT_ERR_ARG, "sqt_multi_create: null argument");
	const int G = p->n_gpus < 1 ? 1 : p->n_gpus;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
		return mfail(nullptr, SQT_ERR_CUDA, "sqt_multi_create: no usable CUDA device; this engine has no CPU fallback");
	if (p->device < 0 || p->device + G > ndev)
		return mfail(nullptr, SQT_ERR_ARG, "sqt_multi_create: devices %d..%d requested, %d present", p->device, p->device + G - 1, ndev);
	sqt_multi *m = new sqt_multi();
	m->G = G; m->p = *p; m->ncomp = 3 + ((p->flag_V || p->flag_s) ? 3 : 0) + (p->flag_M ? 3 : 0);
	m->h.assign(G, nullptr); m->rc.assign(G, 0);
	m->ev[0].assign(G, nullptr); m->ev[1].assign(G, nullptr);
	m->ev_pending[0] = m->ev_pending[1] = false;
	m->gen = 0; m->task = TASK_NONE; m->ndone = 0;
	m->h_stage[0] = m->h_stage[1] = nullptr; m->stage_cap = 0; m->stage_next = 0; m->stage_slot_of_push = -1;
	m->err[0] = 0;
	if (G > 1)
		for (int g = 0; g < G; ++g) m->th.emplace_back(worker, m, g);
	int rc = dispatch(m, TASK_CREATE);
	if (rc == SQT_OK && G > 1) {
		rc = sqt_comm_unique_id(m->comm_id);
		if (rc) mfail(m, rc, "%s", sqt_last_error(nullptr));
		else rc = dispatch(m, TASK_COMM); // every rank joins from its own thread
	}
	if (rc != SQT_OK) {
		char keep[512];
		snprintf(keep, sizeof(keep), "%s", m->err);
		sqt_multi_destroy(m);
		snprintf(g_multi_err, sizeof(g_multi_err), "%s", keep);
		return rc;
	}
	*out = m;
	return SQT_OK;
}

extern "C" void sqt_multi_destroy(sqt_multi *m)
{
	if (!m) return;
	dispatch(m, TASK_DESTROY);
	if (m->G > 1) {
		{
			std::lock_guard<std::mutex> lk(m->mu);
			m->task = TASK_EXIT; ++m->gen;
		}
		m->cv_go.notify_all();
		for (auto &t : m->th) t.join();
	}
	cudaSetDevice(m->p.device);
	for (int s = 0; s < 2; ++s) if (m->h_stage[s]) cudaFreeHost(m->h_stage[s]);
	delete m;
}

extern "C" int sqt_multi_set_limit(sqt_multi *m, int limitCorrAv)
{
	if (!m) return SQT_ERR_ARG;
	for (auto hh : m->h) sqt_set_limit(hh, limitCorrAv);
	return SQT_OK;
}

extern "C" int sqt_multi_set_types(sqt_multi *m, const int *types, int n_atoms)
{
	if (!m || !types || n_atoms < 1) return SQT_ERR_ARG;
	for (int g = 0; g < m->G; ++g) { // the same contiguous chunks the frames are split into
		int lo, hi;
		shard(n_atoms, g, m->G, &lo, &hi);
		const int rc = sqt_set_types(m->h[g], types + lo, hi - lo);
		if (rc) return mfail(m, rc, "GPU %d: %s", m->p.device + g, sqt_last_error(m->h[g]));
	}
	return SQT_OK;
}

extern "C" int sqt_multi_begin_file(sqt_multi *m)
{
	if (!m) return SQT_ERR_ARG;
	for (auto hh : m->h) sqt_begin_file(hh);
	return SQT_OK;
}

static int wait_stage(sqt_multi *m, int s)
{
	if (!m->ev_pending[s]) return SQT_OK;
	for (int g = 0; g < m->G; ++g) {
		cudaSetDevice(m->p.device + g);
		if (cudaEventSynchronize(m->ev[s][g]) != cudaSuccess) return mfail(m, SQT_ERR_CUDA, "stage_acquire: event sync failed on GPU %d", m->p.device + g);
	}
	m->ev_pending[s] = false;
	return SQT_OK;
}

extern "C" double *sqt_multi_stage_acquire(sqt_multi *m, int n_atoms)
{
	if (!m || n_atoms < 1) return nullptr;
	if (n_atoms > m->stage_cap) {
		for (int s = 0; s < 2; ++s) {
			if (wait_stage(m, s)) return nullptr;
			if (m->h_stage[s]) { cudaFreeHost(m->h_stage[s]); m->h_stage[s] = nullptr; }
			// portable: page-locked for every device's context, so each GPU's chunk is a true asynchronous DMA
			if (cudaHostAlloc(&m->h_stage[s], sizeof(double) * m->ncomp * (size_t)n_atoms, cudaHostAllocPortable) != cudaSuccess) {
				mfail(m, SQT_ERR_NOMEM, "cannot pin %zu bytes of host memory", sizeof(double) * m->ncomp * (size_t)n_atoms);
				return nullptr;
			}
		}
		m->stage_cap = n_atoms;
	}
	const int s = m->stage_next;
	m->stage_next ^= 1;
	if (wait_stage(m, s)) return nullptr; // the copies that last read this buffer must be done
	return m->h_stage[s];
}

extern "C" int sqt_multi_push_frame(sqt_multi *m, const double *soa, int n_atoms, double xlo, double xhi)
{
	if (!m || !soa || n_atoms < m->G) return m ? mfail(m, SQT_ERR_ARG, "push_frame: bad arguments (at least one atom per GPU)") : SQT_ERR_ARG;
	m->soa = soa; m->n_atoms = n_atoms; m->xlo = xlo; m->xhi = xhi;
	m->stage_slot_of_push = soa == m->h_stage[0] ? 0 : (soa == m->h_stage[1] ? 1 : -1);
	const int rc = dispatch(m, TASK_PUSH);
	if (rc >= 0 && m->stage_slot_of_push >= 0) m->ev_pending[m->stage_slot_of_push] = true;
	return rc;
}

extern "C" int sqt_multi_sync(sqt_multi *m)
{
	if (!m) return SQT_ERR_ARG;
	return dispatch(m, TASK_SYNC);
}

// rank 0's sums go to the caller; the other ranks' (identical) copies are fetched too, to keep their latch state in step
static int fetch_all(sqt_multi *m, sqt_results *out, bool finish)
{
	int rc = dispatch(m, TASK_SYNC); // every rank has taken part in all allreduces before rank 0's sums are read
	if (rc < 0) return rc;
	for (int g = 1; g < m->G; ++g) {
		sqt_results tmp;
		if (sqt_results_alloc(&tmp, out->nCSpatial, out->nCTime)) return mfail(m, SQT_ERR_NOMEM, "out of host memory");
		rc = finish ? sqt_finish(m->h[g], &tmp) : sqt_fetch_results(m->h[g], &tmp);
		sqt_results_free(&tmp);
		if (rc < 0) return mfail(m, rc, "GPU %d: %s", m->p.device + g, sqt_last_error(m->h[g]));
	}
	rc = finish ? sqt_finish(m->h[0], out) : sqt_fetch_results(m->h[0], out);
	if (rc < 0) mfail(m, rc, "GPU %d: %s", m->p.device, sqt_last_error(m->h[0]));
	return rc;
}

extern "C" int sqt_multi_fetch_results(sqt_multi *m, sqt_results *out) { return (m && out) ? fetch_all(m, out, false) : SQT_ERR_ARG; }
extern "C" int sqt_multi_finish(sqt_multi *m, sqt_results *out) { return (m && out) ? fetch_all(m, out, true) : SQT_ERR_ARG; }

extern "C" long long sqt_multi_kernel_launches(const sqt_multi *m)
{
	long long n = 0;
	if (m) for (auto hh : m->h) n += sqt_kernel_launches(hh);
	return n;
}
