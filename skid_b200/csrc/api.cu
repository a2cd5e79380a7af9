// extern "C" entry points of libskidgpu (include/skidgpu.h).  No exceptions cross the ABI.
#include "ctx.cuh"
#include <algorithm>

thread_local cudaStream_t g_skid_stream = 0;

#define API_BEGIN(ctx)                                                                                 \
	if (!(ctx)) return SKIDGPU_ERR;                                                                \
	try {                                                                                          \
		CK(cudaSetDevice((ctx)->device));                                                      \
		g_skid_stream = (ctx)->stream;
#define API_END(ctx)                                                                                   \
	}                                                                                              \
	catch (const std::exception &e)                                                                \
	{                                                                                              \
		(ctx)->err = e.what();                                                                 \
		return SKIDGPU_ERR;                                                                    \
	}                                                                                              \
	return SKIDGPU_OK;

static std::string g_create_err;

extern "C" int skidgpu_create(skidgpu_ctx **pctx, int device, const float fPeriod[3], const float fCenter[3],
                              int bPeriodic, int bDiag)
{
	if (!pctx) return SKIDGPU_ERR;
	*pctx = nullptr;
	skidgpu_ctx *c = nullptr;
	try {
		int nDev = 0;
		cudaError_t e = cudaGetDeviceCount(&nDev);
		if (e != cudaSuccess || nDev <= 0)
			throw SkidError(std::string("skidgpu_create: no usable CUDA device (") + cudaGetErrorString(e) +
			                "); this library has no CPU fallback");
		if (device < 0 || device >= nDev) throw SkidError("skidgpu_create: bad device ordinal");
		CK(cudaSetDevice(device));
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		if (prop.major < 10)
			throw SkidError("skidgpu_create: device is not sm_100 (Blackwell); this library is built for sm_100a only");
		c = new skidgpu_ctx();
		c->device = device;
		for (int d = 0; d < 3; ++d) {
			c->L[d] = fPeriod[d];
			c->C[d] = fCenter[d];
		}
		c->bPeriodic = bPeriodic;
		c->bDiag = bDiag;
		CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
		{ // keep freed blocks cached in the stream-ordered pool (DevBuf, common.cuh)
			cudaMemPool_t pool;
			unsigned long long keep = ~0ull;
			CK(cudaDeviceGetDefaultMemPool(&pool, device));
			CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
		}
		CK(cudaEventCreate(&c->ev0));
		CK(cudaEventCreate(&c->ev1));
		*pctx = c;
	} catch (const std::exception &e) {
		g_create_err = e.what();
		fprintf(stderr, "%s\n", e.what());
		delete c;
		return SKIDGPU_ERR;
	}
	return SKIDGPU_OK;
}

// Device memory comes from the stream-ordered pool, which grows on demand: on a first pass that growth (driver
// allocation + mapping, ~11 GB at 2^24 particles) is host time on the critical path.  A caller that knows the
// particle count early (the TIPSY header) can grow the pool ahead of time, e.g. while the records are still
// being read: the block is allocated and freed at once and stays cached in the pool.
extern "C" int skidgpu_reserve(skidgpu_ctx *ctx, unsigned long long bytes)
{
	API_BEGIN(ctx)
	size_t freeB = 0, totalB = 0;
	CK(cudaMemGetInfo(&freeB, &totalB));
	if (bytes > freeB / 2) bytes = freeB / 2; // a hint, never a reason to fail
	if (bytes > 0) {
		void *p = nullptr;
		CK(cudaMallocAsync(&p, (size_t)bytes, ctx->stream));
		CK(cudaFreeAsync(p, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
	}
	API_END(ctx)
}

extern "C" void skidgpu_destroy(skidgpu_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	dist_comm_destroy(*ctx);
	for (cudaEvent_t e : ctx->logEv) cudaEventDestroy(e);
	if (ctx->hLog) cudaFreeHost(ctx->hLog);
	cudaStream_t s = ctx->stream;
	g_skid_stream = s;
	delete ctx; // DevBuf destructors free on s
	if (s) {
		cudaStreamSynchronize(s);
		cudaStreamDestroy(s);
	}
	g_skid_stream = 0;
}

extern "C" const char *skidgpu_last_error(skidgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int skidgpu_set_shard(skidgpu_ctx *ctx, int rank, int nranks)
{
	API_BEGIN(ctx)
	if (nranks < 1 || rank < 0 || rank >= nranks) throw SkidError("skidgpu_set_shard: bad rank/nranks");
	ctx->rank = rank;
	ctx->nranks = nranks;
	API_END(ctx)
}

extern "C" int skidgpu_comm_unique_id(void *id128)
{
	try {
		if (!id128) throw SkidError("skidgpu_comm_unique_id: null buffer");
		dist_unique_id(id128);
	} catch (const std::exception &e) {
		g_create_err = e.what();
		return SKIDGPU_ERR;
	}
	return SKIDGPU_OK;
}

extern "C" int skidgpu_comm_init(skidgpu_ctx *ctx, const void *id128, int rank, int nranks)
{
	API_BEGIN(ctx)
	if (!id128 && nranks > 1) throw SkidError("skidgpu_comm_init: null unique id");
	dist_comm_init(*ctx, id128, rank, nranks);
	API_END(ctx)
}

extern "C" long long skidgpu_comm_bytes(skidgpu_ctx *ctx, long long *nCalls)
{
	if (!ctx) return -1;
	if (nCalls) *nCalls = ctx->commCalls;
	return ctx->commBytes;
}

extern "C" int skidgpu_set_profile(skidgpu_ctx *ctx, int bOn)
{
	API_BEGIN(ctx)
	ctx->spans.on = bOn != 0;
	API_END(ctx)
}

extern "C" int skidgpu_debug_move_kernel(skidgpu_ctx *ctx, int which)
{
	API_BEGIN(ctx)
	if (which < 0 || which > 1) throw SkidError("skidgpu_debug_move_kernel: 0 = tiles, 1 = a tree walk per mover and step");
	ctx->moveKernel = which;
	API_END(ctx)
}

__global__ void __launch_bounds__(256)
    k_aos_to_soa(int n, const skidgpu_pinit *p, float *x, float *y, float *z, float *vx, float *vy, float *vz,
                 float *mass, float *soft, float *temp)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	skidgpu_pinit q = p[i];
	x[i] = q.r[0];
	y[i] = q.r[1];
	z[i] = q.r[2];
	vx[i] = q.v[0];
	vy[i] = q.v[1];
	vz[i] = q.v[2];
	mass[i] = q.fMass;
	soft[i] = q.fSoft;
	temp[i] = q.fTemp;
}

static void set_counts(skidgpu_ctx *c, int n, int nGas, int nDark, int nStar)
{
	if (n <= 0 || nGas < 0 || nDark < 0 || nStar < 0 || nGas + nDark + nStar != n)
		throw SkidError("skidgpu_set_particles: need n = nGas + nDark + nStar > 0");
	c->n = n;
	c->nGas = nGas;
	c->nDark = nDark;
	c->nStar = nStar;
	c->inType = (nDark ? SKIDGPU_DARK : 0) | (nGas ? SKIDGPU_GAS : 0) | (nStar ? SKIDGPU_STAR : 0); // kd.c:143-149
	c->nMove = c->nActive = 0;
	c->nGroup = 0;
	c->nEnt = c->nExtra = c->nAct = 0;
	c->haveCenters = false;
	c->haveRhoStat = false;
	if (n > c->reservedFor) { // first pass at this size: grow the pool in one piece, not buffer by buffer
		size_t freeB = 0, totalB = 0;
		// measured footprint of the hot path: ~250 B/particle replicated (input, trees, scatterers)
		// + ~450 B/particle for movers and tile lists, which are sharded across ranks
		unsigned long long want = (250ull + 450ull / (unsigned long long)c->nranks) * (unsigned long long)n;
		CK(cudaMemGetInfo(&freeB, &totalB));
		if (want > freeB / 2) want = freeB / 2;
		void *p = nullptr;
		if (want > 0 && cudaMallocAsync(&p, (size_t)want, c->stream) == cudaSuccess) CK(cudaFreeAsync(p, c->stream));
		else (void)cudaGetLastError();
		c->reservedFor = n;
	}
	c->x.alloc(n);
	c->y.alloc(n);
	c->z.alloc(n);
	c->vx.alloc(n);
	c->vy.alloc(n);
	c->vz.alloc(n);
	c->mass.alloc(n);
	c->soft.alloc(n);
	c->temp.alloc(n);
}

extern "C" int skidgpu_set_particles(skidgpu_ctx *ctx, const skidgpu_pinit *p, int n, int nGas, int nDark, int nStar)
{
	API_BEGIN(ctx)
	if (!p) throw SkidError("skidgpu_set_particles: null particle array");
	set_counts(ctx, n, nGas, nDark, nStar);
	for (int i = 0; i < n; i += (n > 4096 ? n / 7 + 1 : 1))
		if (p[i].iOrder != i) throw SkidError("skidgpu_set_particles: particles must be in file order (iOrder == index)");
	skidgpu_pinit *d;
	if (ctx->comm && ctx->nranks > 1) {
		// sharded run: every rank holds the same host snapshot (the premise of every collective stage call), so
		// the records need not cross PCIe N times - each rank uploads its 1/N slice and the slices are all-gathered
		// over NVLink (6.4 GB -> 0.8 GB of host->device traffic per rank at 2^27 on 8 GPUs)
		const size_t chunk = ceil_div((size_t)n, (size_t)ctx->nranks);
		d = ctx->aos.alloc(chunk * ctx->nranks);
		const size_t lo = std::min((size_t)n, chunk * ctx->rank), hi = std::min((size_t)n, chunk * (ctx->rank + 1));
		if (hi > lo) CK(cudaMemcpyAsync(d + lo, p + lo, sizeof(skidgpu_pinit) * (hi - lo), cudaMemcpyHostToDevice, ctx->stream));
		sk_allgather(*ctx, d, (long long)(chunk * sizeof(skidgpu_pinit)), SK_U8);
	} else {
		d = ctx->aos.alloc(n);
		CK(cudaMemcpyAsync(d, p, sizeof(skidgpu_pinit) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
	}
	SK_LAUNCH(k_aos_to_soa, (unsigned)ceil_div(n, 256), 256, 0, ctx->stream, n, d, ctx->x.p, ctx->y.p, ctx->z.p, ctx->vx.p,
	          ctx->vy.p, ctx->vz.p, ctx->mass.p, ctx->soft.p, ctx->temp.p);
	CK(cudaStreamSynchronize(ctx->stream));
	API_END(ctx)
}

extern "C" int skidgpu_set_particles_dev(skidgpu_ctx *ctx, const float *dx, const float *dy, const float *dz,
                                         const float *dvx, const float *dvy, const float *dvz, const float *dmass,
                                         const float *dsoft, const float *dtemp, int n, int nGas, int nDark,
                                         int nStar)
{
	API_BEGIN(ctx)
	set_counts(ctx, n, nGas, nDark, nStar);
	const float *src[9] = {dx, dy, dz, dvx, dvy, dvz, dmass, dsoft, dtemp};
	float *dst[9] = {ctx->x.p, ctx->y.p, ctx->z.p, ctx->vx.p, ctx->vy.p, ctx->vz.p, ctx->mass.p, ctx->soft.p, ctx->temp.p};
	for (int k = 0; k < 9; ++k) {
		if (!src[k]) throw SkidError("skidgpu_set_particles_dev: null device array");
		CK(cudaMemcpyAsync(dst[k], src[k], sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
	}
	CK(cudaStreamSynchronize(ctx->stream));
	API_END(ctx)
}

__global__ void __launch_bounds__(256) k_fill(int n, float *a, float v)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) a[i] = v;
}

extern "C" int skidgpu_set_soft(skidgpu_ctx *ctx, float fEps)
{
	API_BEGIN(ctx)
	if (ctx->n <= 0) throw SkidError("skidgpu_set_soft: no particles set");
	SK_LAUNCH(k_fill, (unsigned)ceil_div(ctx->n, 256), 256, 0, ctx->stream, ctx->n, ctx->soft.p, fEps);
	API_END(ctx)
}

extern "C" int skidgpu_density(skidgpu_ctx *ctx, int nSmooth, int bGasAndDark, int bGasOnly, float *rho_by_iOrder,
                               float *ball2_by_iOrder, int *nExtraScat)
{
	API_BEGIN(ctx)
	stage_density(*ctx, nSmooth, bGasAndDark, bGasOnly, nExtraScat);
	if (rho_by_iOrder)
		CK(cudaMemcpyAsync(rho_by_iOrder, ctx->rho.p, sizeof(float) * (size_t)ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
	if (ball2_by_iOrder)
		CK(cudaMemcpyAsync(ball2_by_iOrder, ctx->ball2.p, sizeof(float) * (size_t)ctx->n, cudaMemcpyDeviceToHost,
		                   ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	API_END(ctx)
}

extern "C" int skidgpu_keep_neighbors(skidgpu_ctx *ctx, int bKeep)
{
	API_BEGIN(ctx)
	ctx->keepNbr = bKeep != 0;
	API_END(ctx)
}

extern "C" int skidgpu_get_neighbors(skidgpu_ctx *ctx, int *nbr, float *d2)
{
	API_BEGIN(ctx)
	if (!ctx->keepNbr || !ctx->nbr.p) throw SkidError("skidgpu_get_neighbors: enable skidgpu_keep_neighbors before skidgpu_density");
	size_t cnt = (size_t)ctx->n * ctx->nSmooth;
	if (nbr) CK(cudaMemcpyAsync(nbr, ctx->nbr.p, sizeof(int) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
	if (d2) CK(cudaMemcpyAsync(d2, ctx->nbrD2.p, sizeof(float) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	API_END(ctx)
}

extern "C" int skidgpu_move(skidgpu_ctx *ctx, float fDensMin, float fTempMax, float fMassMax, float fCvg, float fStep,
                            int bForceInitialCut, int bNoPrune, skidgpu_log_cb cb, void *user, int *nMove, int *nIttr)
{
	API_BEGIN(ctx)
	stage_move(*ctx, fDensMin, fTempMax, fMassMax, fCvg, fStep, bForceInitialCut, bNoPrune, cb, user, nMove, nIttr);
	API_END(ctx)
}

extern "C" int skidgpu_keep_step0(skidgpu_ctx *ctx, int bKeep)
{
	API_BEGIN(ctx)
	ctx->keepStep0 = bKeep != 0;
	API_END(ctx)
}

extern "C" int skidgpu_get_step0(skidgpu_ctx *ctx, int *iOrder, float *a3, unsigned char *scat_alive_by_iOrder)
{
	API_BEGIN(ctx)
	if (!ctx->keepStep0 || !ctx->a0x.p) throw SkidError("skidgpu_get_step0: enable skidgpu_keep_step0 before skidgpu_move");
	const int m = ctx->nMove;
	std::vector<float> hx(m), hy(m), hz(m);
	cudaStream_t s = ctx->stream;
	if (iOrder) CK(cudaMemcpyAsync(iOrder, ctx->mOrd.p, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(hx.data(), ctx->a0x.p, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(hy.data(), ctx->a0y.p, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(hz.data(), ctx->a0z.p, sizeof(float) * m, cudaMemcpyDeviceToHost, s));
	if (scat_alive_by_iOrder && ctx->aliveByOrd.p)
		CK(cudaMemcpyAsync(scat_alive_by_iOrder, ctx->aliveByOrd.p, ctx->n, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	if (a3)
		for (int i = 0; i < m; ++i) {
			a3[3 * i] = hx[i];
			a3[3 * i + 1] = hy[i];
			a3[3 * i + 2] = hz[i];
		}
	API_END(ctx)
}

extern "C" int skidgpu_fof(skidgpu_ctx *ctx, float fTau, int *nGroup)
{
	API_BEGIN(ctx)
	stage_fof(*ctx, fTau, nGroup);
	API_END(ctx)
}

extern "C" int skidgpu_microstep(skidgpu_ctx *ctx, int nSteps, float fStep, skidgpu_log_cb cb, void *user)
{
	API_BEGIN(ctx)
	stage_microstep(*ctx, nSteps, fStep, cb, user);
	API_END(ctx)
}

// Movers live in Morton order on the device; kdOutVector wants them in ascending iOrder (kd.c:1561).  A mover's
// output slot is the number of movers with a smaller iOrder: mark, scan, scatter - all on the device, then two
// contiguous copies straight into the caller's arrays (a host-side index sort cost 1.4 s at 8.3 M movers).
__global__ void __launch_bounds__(256) k_mark_movers(int m, const int *mOrd, uint32_t *flags)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < m) flags[mOrd[j]] = 1u;
}

__global__ void __launch_bounds__(256)
    k_pack_moved(int m, const int *mOrd, const uint32_t *pos, const float *mx, const float *my, const float *mz,
                 int *outOrd, float *outR3)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m) return;
	const int o = mOrd[j];
	const size_t i = pos[o];
	outOrd[i] = o;
	outR3[3 * i] = mx[j];
	outR3[3 * i + 1] = my[j];
	outR3[3 * i + 2] = mz[j];
}

extern "C" int skidgpu_get_moved(skidgpu_ctx *ctx, int *iOrder, float *r3)
{
	API_BEGIN(ctx)
	const int m = ctx->nMove, n = ctx->n;
	if (m > 0) {
		cudaStream_t s = ctx->stream;
		DevBuf<int> dOrd;
		DevBuf<float> dR3;
		uint32_t *flags = ctx->flags.alloc((size_t)n + 1);
		uint32_t *scan = ctx->scan.alloc((size_t)n + 64);
		dOrd.alloc(m);
		dR3.alloc((size_t)3 * m);
		CK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * ((size_t)n + 1), s));
		SK_LAUNCH(k_mark_movers, (unsigned)ceil_div(m, 256), 256, 0, s, m, ctx->mOrd.p, flags);
		exclusive_scan_u32(flags, scan, n, ctx->ws, s);
		SK_LAUNCH(k_pack_moved, (unsigned)ceil_div(m, 256), 256, 0, s, m, ctx->mOrd.p, scan, ctx->mx.p, ctx->my.p,
		          ctx->mz.p, dOrd.p, dR3.p);
		if (iOrder) CK(cudaMemcpyAsync(iOrder, dOrd.p, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost, s));
		if (r3) CK(cudaMemcpyAsync(r3, dR3.p, sizeof(float) * 3 * (size_t)m, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s)); // before the local buffers are released
	}
	API_END(ctx)
}

static void fetch_catalogue(skidgpu_ctx *ctx, int *piGroup, skidgpu_pgroup *g)
{
	cudaStream_t s = ctx->stream;
	if (piGroup) CK(cudaMemcpyAsync(piGroup, ctx->gid.p, sizeof(int) * (size_t)ctx->n, cudaMemcpyDeviceToHost, s));
	if (g) CK(cudaMemcpyAsync(g, ctx->gCat.p, sizeof(skidgpu_pgroup) * (size_t)ctx->nGroup, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
}

extern "C" int skidgpu_centers(skidgpu_ctx *ctx, int *piGroup_by_iOrder, skidgpu_pgroup *g)
{
	API_BEGIN(ctx)
	stage_centers(*ctx);
	fetch_catalogue(ctx, piGroup_by_iOrder, g);
	API_END(ctx)
}

extern "C" int skidgpu_set_groups(skidgpu_ctx *ctx, const int *piGroup_by_iOrder, int nGroup,
                                  const skidgpu_pgroup *centres)
{
	API_BEGIN(ctx)
	if (!piGroup_by_iOrder) throw SkidError("skidgpu_set_groups: null group array");
	stage_set_groups(*ctx, piGroup_by_iOrder, nGroup, centres);
	API_END(ctx)
}

extern "C" int skidgpu_unbind(skidgpu_ctx *ctx, float fG, float z, double fCosmo, int iSoftType, float fScoop,
                              int bNoUnbind, int nMaxMembers, int nMinMembers, int *piGroup_by_iOrder,
                              skidgpu_pgroup *g, int *nGroup, int *nUnbound, int *nGroupBefore)
{
	API_BEGIN(ctx)
	stage_unbind(*ctx, fG, z, fCosmo, iSoftType, fScoop, bNoUnbind, nMaxMembers, nMinMembers, nUnbound, nGroupBefore);
	if (nGroup) *nGroup = ctx->nGroup;
	fetch_catalogue(ctx, piGroup_by_iOrder, g);
	API_END(ctx)
}

extern "C" int skidgpu_stats(skidgpu_ctx *ctx, float fG, float z, double fExpHub, float fDensMin, float fTempMax,
                             skidgpu_stat_row *rows)
{
	API_BEGIN(ctx)
	if (!rows) throw SkidError("skidgpu_stats: null output array");
	stage_stats(*ctx, fG, z, fExpHub, fDensMin, fTempMax, rows);
	API_END(ctx)
}

extern "C" double skidgpu_stage_ms(skidgpu_ctx *ctx, int stage)
{
	if (!ctx || stage < 0 || stage > 5) return -1.0;
	return ctx->stage_ms[stage];
}

extern "C" int skidgpu_set_reduce_cb(skidgpu_ctx *ctx, skidgpu_reduce_cb cb, void *user)
{
	API_BEGIN(ctx)
	ctx->reduceCb = cb;
	ctx->reduceUser = user;
	API_END(ctx)
}

extern "C" double skidgpu_kernel_ms(skidgpu_ctx *ctx, int which, int *nLaunches)
{
	if (!ctx || which < 0 || which >= KF_COUNT) return -1.0;
	if (nLaunches) *nLaunches = ctx->kernel_launches[which];
	return ctx->kernel_ms[which];
}

extern "C" void *skidgpu_stream(skidgpu_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" long long skidgpu_counter(skidgpu_ctx *ctx, int which)
{
	if (!ctx) return -1;
	switch (which) {
	case 0: return g_skid_launches;
	case 1: return ctx->moverSteps;
	case 2: return ctx->nQueries;
	case 3: return ctx->nPairs;
	}
	return -1;
}

// ---- test hooks for the hand-written primitives (host arrays in, host arrays out)
extern "C" int skidgpu_debug_sort(skidgpu_ctx *ctx, unsigned long long *keys, unsigned int *vals, long long n, int bits)
{
	API_BEGIN(ctx)
	DevBuf<uint64_t> k;
	DevBuf<uint32_t> v;
	k.alloc(n > 0 ? n : 1);
	v.alloc(n > 0 ? n : 1);
	cudaStream_t s = ctx->stream;
	CK(cudaMemcpyAsync(k.p, keys, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
	CK(cudaMemcpyAsync(v.p, vals, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
	radix_sort_pairs(k.p, v.p, (size_t)n, bits, ctx->ws, s);
	CK(cudaMemcpyAsync(keys, k.p, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(vals, v.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	API_END(ctx)
}

extern "C" int skidgpu_debug_scan(skidgpu_ctx *ctx, const unsigned int *in, unsigned int *out, long long n)
{
	API_BEGIN(ctx)
	DevBuf<uint32_t> a, b;
	a.alloc(n > 0 ? n : 1);
	b.alloc(n + 1);
	cudaStream_t s = ctx->stream;
	CK(cudaMemcpyAsync(a.p, in, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
	exclusive_scan_u32(a.p, b.p, (size_t)n, ctx->ws, s);
	CK(cudaMemcpyAsync(out, b.p, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	API_END(ctx)
}
