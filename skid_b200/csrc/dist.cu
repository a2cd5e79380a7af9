// Multi-GPU exchange points of the sharded pipeline, issued by the library itself with NCCL on
// the context's own stream: one communicator per context (one context per GPU; one host thread or one
// process per context).  NCCL is bound with dlopen at the first skidgpu_comm_* call, so the library has no
// link-time dependency on it (single-GPU users and the CPU-side ABI checks never load it).
//
// Exchanges (all stream-ordered, none synchronises the host):
//   sk_reduce      all-reduce (min / max / sum) of small agreement values and of the f64 density partials
//   sk_allgather   equal-sized blocks, in place (fBall2 of the sharded kNN queries, mover positions)
//   sk_allgatherv  variable-sized blocks (pieces of a distributed sort)
// Without a communicator the legacy reduce callback (skidgpu_set_reduce_cb, a test shim) serves sk_reduce
// and the gathers fall back to zero-fill + sum through it.
#include "ctx.cuh"
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>

namespace {
struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

template <class F> void bind(F &f, const char *name)
{
	f = (F)dlsym(g_nccl.h, name);
	if (!f) throw SkidError(std::string("NCCL symbol missing: ") + name);
}

NcclApi &nccl()
{
	std::lock_guard<std::mutex> lk(g_nccl_mu);
	if (g_nccl.h) return g_nccl;
	// a process that already carries an NCCL (e.g. the one bundled with torch) gets that one: same soname
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (!h) throw SkidError(std::string("cannot load libnccl.so.2: ") + dlerror());
	g_nccl.h = h;
	bind(g_nccl.GetUniqueId, "ncclGetUniqueId");
	bind(g_nccl.CommInitRank, "ncclCommInitRank");
	bind(g_nccl.CommDestroy, "ncclCommDestroy");
	bind(g_nccl.AllReduce, "ncclAllReduce");
	bind(g_nccl.AllGather, "ncclAllGather");
	bind(g_nccl.Broadcast, "ncclBroadcast");
	bind(g_nccl.GroupStart, "ncclGroupStart");
	bind(g_nccl.GroupEnd, "ncclGroupEnd");
	bind(g_nccl.GetErrorString, "ncclGetErrorString");
	return g_nccl;
}

#define NK(call)                                                                                       \
	do {                                                                                           \
		ncclResult_t r_ = (call);                                                              \
		if (r_ != ncclSuccess) {                                                               \
			char b_[384];                                                                  \
			snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
			throw SkidError(b_);                                                           \
		}                                                                                      \
	} while (0)

ncclDataType_t nccl_type(int dtype)
{
	switch (dtype) {
	case SK_I32: return ncclInt32;
	case SK_U8: return ncclUint8;
	case SK_F32: return ncclFloat32;
	case SK_F64: return ncclFloat64;
	}
	throw SkidError("sk_reduce: bad dtype");
}
size_t type_size(int dtype) { return dtype == SK_U8 ? 1 : (dtype == SK_F64 ? 8 : 4); }
} // namespace

void dist_unique_id(void *id128)
{
	static_assert(sizeof(ncclUniqueId) == SKIDGPU_UNIQUE_ID_BYTES, "ncclUniqueId size");
	NcclApi &a = nccl();
	NK(a.GetUniqueId((ncclUniqueId *)id128));
}

void dist_comm_init(skidgpu_ctx &c, const void *id128, int rank, int nranks)
{
	if (nranks < 1 || rank < 0 || rank >= nranks) throw SkidError("skidgpu_comm_init: bad rank/nranks");
	dist_comm_destroy(c);
	c.rank = rank;
	c.nranks = nranks;
	if (nranks == 1) return;
	NcclApi &a = nccl();
	ncclUniqueId id;
	memcpy(&id, id128, sizeof id);
	ncclComm_t comm = nullptr;
	NK(a.CommInitRank(&comm, nranks, id, rank));
	c.comm = comm;
}

void dist_comm_destroy(skidgpu_ctx &c)
{
	if (c.comm) {
		cudaStreamSynchronize(c.stream);
		g_nccl.CommDestroy((ncclComm_t)c.comm);
		c.comm = nullptr;
	}
}

void sk_reduce(skidgpu_ctx &c, void *dev, long long count, int dtype, int op)
{
	if (c.nranks <= 1 || count <= 0) return;
	c.commBytes += (long long)type_size(dtype) * count;
	++c.commCalls;
	if (c.comm) {
		const ncclRedOp_t rop = op == SK_MIN ? ncclMin : (op == SK_MAX ? ncclMax : ncclSum);
		NK(g_nccl.AllReduce(dev, dev, (size_t)count, nccl_type(dtype), rop, (ncclComm_t)c.comm, c.stream));
		return;
	}
	if (!c.reduceCb) throw SkidError("nranks > 1 but neither a communicator (skidgpu_comm_init) nor a reduce callback is set");
	if (c.reduceCb(c.reduceUser, dev, count, dtype, op) != 0) throw SkidError("reduce callback failed");
}

// buf holds nranks blocks of `per` elements; this rank's block (index rank) is filled, the others are
// received.  In place.
void sk_allgather(skidgpu_ctx &c, void *buf, long long per, int dtype)
{
	if (c.nranks <= 1 || per <= 0) return;
	const size_t es = type_size(dtype);
	c.commBytes += (long long)es * per * c.nranks;
	++c.commCalls;
	if (c.comm) {
		NK(g_nccl.AllGather((char *)buf + es * (size_t)per * c.rank, buf, (size_t)per, nccl_type(dtype), (ncclComm_t)c.comm,
		                    c.stream));
		return;
	}
	// callback shim: zero what the others own, then sum
	if (c.rank > 0) CK(cudaMemsetAsync(buf, 0, es * (size_t)per * c.rank, c.stream));
	if (c.rank < c.nranks - 1)
		CK(cudaMemsetAsync((char *)buf + es * (size_t)per * (c.rank + 1), 0, es * (size_t)per * (c.nranks - 1 - c.rank), c.stream));
	c.commBytes -= (long long)es * per * c.nranks;
	--c.commCalls;
	sk_reduce(c, buf, per * c.nranks, dtype == SK_U8 ? SK_U8 : dtype, dtype == SK_U8 ? SK_MAX : SK_SUM);
}

// Variable block sizes: rank r contributes counts[r] elements, placed at offs[r] of recv (host arrays, known
// to every rank).  send = this rank's block (may alias recv + offs[rank]).
void sk_allgatherv(skidgpu_ctx &c, const void *send, void *recv, const long long *counts, const long long *offs, int dtype)
{
	if (c.nranks <= 1) return;
	const size_t es = type_size(dtype);
	if (!c.comm) throw SkidError("sk_allgatherv needs a communicator (skidgpu_comm_init)");
	NK(g_nccl.GroupStart());
	for (int r = 0; r < c.nranks; ++r) {
		if (counts[r] <= 0) continue;
		c.commBytes += (long long)es * counts[r];
		NK(g_nccl.Broadcast(r == c.rank ? send : (const void *)((char *)recv + es * (size_t)offs[r]), (char *)recv + es * (size_t)offs[r],
		                    (size_t)counts[r], nccl_type(dtype), r, (ncclComm_t)c.comm, c.stream));
	}
	NK(g_nccl.GroupEnd());
	++c.commCalls;
}
