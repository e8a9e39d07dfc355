// sqt_comm.cpp -- NCCL shim for the particle-sharded path (one process / handle per GPU).
//
// The only data-path exchanges are:
//   * per frame: sum of the partial rho(k) vectors and of the partial sum of positions,
//     2*3*nCSpatial + 3 doubles (consumed by collective_ring_kernel only);
//   * per completed buffer: sum of the buffer's self block before it is squared into StdDev.
// Both are tiny, latency-bound allreduces over NVLink; they run on a side stream so that they
// overlap the next frame's ring kernel.
//
// NCCL is resolved at run time with dlopen so that (a) the single-GPU library has no NCCL
// dependency at all and (b) inside a torch process we bind to the very libnccl.so.2 torch
// already loaded (RTLD_NOLOAD first) instead of pulling in a second copy.
#include "sqt_internal.h"
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
	void *lib;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int nccl_load(char *err, int errlen)
{
	if (g_nccl.lib) return 0;
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	void *lib = nullptr;
	for (int pass = 0; pass < 2 && !lib; ++pass)
		for (int i = 0; i < 2 && !lib; ++i)
			lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL | (pass == 0 ? RTLD_NOLOAD : 0));
	if (!lib) { snprintf(err, errlen, "cannot load libnccl.so.2: %s", dlerror()); return -1; }
#define SYM(field, name) \
	*(void **)(&g_nccl.field) = dlsym(lib, name); \
	if (!g_nccl.field) { snprintf(err, errlen, "libnccl lacks %s", name); return -1; }
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(AllReduce, "ncclAllReduce");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	g_nccl.lib = lib;
	return 0;
}

struct SqtComm { ncclComm_t comm; int rank, nranks; };

int sqt_comm_get_unique_id(void *out128, char *err, int errlen)
{
	if (nccl_load(err, errlen)) return -1;
	ncclUniqueId id;
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r) { snprintf(err, errlen, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r)); return -1; }
	memcpy(out128, &id, sizeof(id));
	return 0;
}

int sqt_comm_create(SqtComm **out, const void *id128, int rank, int nranks, char *err, int errlen)
{
	if (nccl_load(err, errlen)) return -1;
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	SqtComm *c = new SqtComm();
	c->rank = rank; c->nranks = nranks;
	ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
	if (r) { snprintf(err, errlen, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); delete c; return -1; }
	*out = c;
	return 0;
}

int sqt_comm_allreduce_f64(SqtComm *c, double *buf, size_t count, cudaStream_t st, char *err, int errlen)
{
	ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->comm, st);
	if (r) { snprintf(err, errlen, "ncclAllReduce: %s", g_nccl.GetErrorString(r)); return -1; }
	return 0;
}

void sqt_comm_destroy(SqtComm *c)
{
	if (!c) return;
	if (g_nccl.lib && c->comm) g_nccl.CommDestroy(c->comm);
	delete c;
}
