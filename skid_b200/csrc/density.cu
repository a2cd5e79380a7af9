// Stage 2: exact periodic k-nearest-neighbour search + symmetric cubic-spline density, then the
// periodic replica scatterers and the static ball-inflated scatterer tree used by the move loop.
//
// Replaces kdScatterActive (kd.c:669-693, ScatterCriterion 600-627), kdBuildTree (kd.c:371),
// smBallSearch (smooth1.c:41-129), the smDensityInit main loop (smooth1.c:150-277) and its
// replica construction (smooth1.c:278-332).
//
// Design: one warp per query.  Queries run in Morton order, the k best candidates live in
// registers (2, 4 or 8 packed (d2,index) words per lane for k <= 64 / 128 / 256, sorted across the warp), new
// candidates are staged in a 64-entry shared-memory buffer and merged 32 at a time with bitonic networks.
// The result is the k smallest by (d2, tree index): deterministic, independent of visit order.
#include "ctx.cuh"
#include <algorithm>

// ------------------------------------------------------------------ species rules
// ScatterCriterion (kd.c:600-627).  type by iOrder range (kdParticleType, kd.c:113-119).
__device__ __forceinline__ int ptype(int i, int nGas, int nDark)
{
	return i < nGas ? SKIDGPU_GAS : (i < nGas + nDark ? SKIDGPU_DARK : SKIDGPU_STAR);
}

__global__ void __launch_bounds__(256) k_scatter_flags(int n, int nGas, int nDark, int inType, int bGasAndDark,
                                                       int bGasOnly, uint32_t *flags)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int t = ptype(i, nGas, nDark);
	int f = 0;
	switch (inType) {
	case SKIDGPU_DARK: f = 1; break;
	case SKIDGPU_GAS:
	case SKIDGPU_DARK | SKIDGPU_GAS: f = bGasAndDark ? 1 : (t == SKIDGPU_GAS); break;
	case SKIDGPU_STAR:
	case SKIDGPU_DARK | SKIDGPU_STAR: f = (t == SKIDGPU_STAR); break;
	case SKIDGPU_GAS | SKIDGPU_STAR:
	case SKIDGPU_DARK | SKIDGPU_GAS | SKIDGPU_STAR:
		if (bGasAndDark) f = 1;
		else if (t == SKIDGPU_GAS) f = 1;
		else if (t == SKIDGPU_STAR && !bGasOnly) f = 1;
		break;
	}
	flags[i] = (uint32_t)f;
}

__global__ void __launch_bounds__(256) k_compact_idx(int n, const uint32_t *flags, const uint32_t *scan,
                                                     uint32_t *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flags[i]) out[scan[i]] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_gather3(int m, const uint32_t *idx, const float *x, const float *y,
                                                 const float *z, float *ox, float *oy, float *oz)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t j = idx[i];
	ox[i] = x[j];
	oy[i] = y[j];
	oz[i] = z[j];
}

// sorted arrays of the active set: posA = (x,y,z,mass), iordA = file index
__global__ void __launch_bounds__(256) k_gather_sortedA(int m, const uint32_t *perm, const uint32_t *actIdx,
                                                        const float *x, const float *y, const float *z,
                                                        const float *mass, float4 *posA, int *iordA)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t j = actIdx[perm[i]];
	posA[i] = make_float4(x[j], y[j], z[j], mass[j]);
	iordA[i] = (int)j;
}

// ------------------------------------------------------------------ warp-level k-best (k <= 32 R)
#define KNN_INF 0x7f800000ffffffffull

__device__ __forceinline__ uint64_t u64min(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t u64max(uint64_t a, uint64_t b) { return a < b ? b : a; }

// One compare-exchange of the networks below: this lane keeps the smaller (keepMin) or the larger of its value
// and its partner's.  ONE 64-bit compare decides (ncu: the kernel is bound by the ALU pipe - ISETP/SEL - not by
// issue slots, so the min-and-max-then-select form cost twice the compares).
// (Comparing the keys as positive float64 - DSETP on the idle FP64 pipe - was measured: 99.3 -> 98.5 ms, the
// register-pair moves eat the gain; the integer compare stays.)
__device__ __forceinline__ bool key_lt(uint64_t a, uint64_t b) { return a < b; }
__device__ __forceinline__ uint64_t cmpx(uint64_t v, uint64_t o, bool keepMin) { return (key_lt(v, o) == keepMin) ? v : o; }

// ascending bitonic sort of one value per lane
__device__ __forceinline__ uint64_t warp_sort32(uint64_t v, int lane)
{
#pragma unroll
	for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
		for (int j = k >> 1; j > 0; j >>= 1) {
			const uint64_t o = __shfl_xor_sync(SK_FULL, v, j);
			v = cmpx(v, o, ((lane & k) == 0) == ((lane & j) == 0));
		}
	}
	return v;
}
// ascending merge of a bitonic sequence (one value per lane)
__device__ __forceinline__ uint64_t warp_bitonic_merge32(uint64_t v, int lane)
{
#pragma unroll
	for (int j = 16; j > 0; j >>= 1) {
		const uint64_t o = __shfl_xor_sync(SK_FULL, v, j);
		v = cmpx(v, o, (lane & j) == 0);
	}
	return v;
}
// A = R registers per lane, a[r] holds ranks 32r .. 32r+31 ascending; b = 32 unsorted candidates.  Keeps the
// 32 R smallest of A U b.  The new block cascades down from the top: against a[R-1] only its 32 smallest
// survive, against every lower register the pair is split into a low and a high half; the cascade stops as
// soon as everything still travelling is >= the next register's largest entry (the common case late in a
// walk: new candidates sit just under the bound).
template <int R> __device__ __forceinline__ void kbest_merge(uint64_t (&a)[R], uint64_t b, int lane)
{
	b = warp_sort32(b, lane);
	uint64_t l = warp_bitonic_merge32(u64min(a[R - 1], __shfl_sync(SK_FULL, b, 31 - lane)), lane);
#pragma unroll
	for (int r = R - 2; r >= 0; --r) {
		if (__shfl_sync(SK_FULL, l, 0) >= __shfl_sync(SK_FULL, a[r], 31)) {
			a[r + 1] = l;
			return;
		}
		const uint64_t lr = __shfl_sync(SK_FULL, l, 31 - lane);
		const bool lt = key_lt(a[r], lr);
		const uint64_t lo = lt ? a[r] : lr, hi = lt ? lr : a[r];
		a[r + 1] = warp_bitonic_merge32(hi, lane);
		l = warp_bitonic_merge32(lo, lane);
	}
	a[0] = l;
}
// d2 of the k-th best (+inf while fewer than k candidates are known)
template <int R> __device__ __forceinline__ float kbest_kth(const uint64_t (&a)[R], int k)
{
	uint64_t v = a[0];
#pragma unroll
	for (int r = 1; r < R; ++r)
		if (((k - 1) >> 5) == r) v = a[r];
	return __uint_as_float((uint32_t)(__shfl_sync(SK_FULL, v, (k - 1) & 31) >> 32));
}

struct KnnArgs {
	TreeView tv;
	const float4 *pos4;
	const int *iord;
	int n, k;
	int qLo, qHi; // this rank's range of queries (sorted order); all of [0,n) on one GPU
	float L[3], hL[3];
	float boxLo[3], boxHi[3]; // the periodic box (centre -+ L/2); +-FLT_MAX when not periodic
	float *ball2;
	double *rho64;
	int *nbr;     // nullable, [nFile*k] by file index
	float *nbrD2; // nullable
};

constexpr int KNN_WARPS = 8;
constexpr int KNN_QPW = 8; // consecutive (Morton-adjacent) queries per warp, one after the other

struct KnnQuery {
	float x0, y0, z0, xp, xm, yp, ym, zp, zm, hx, hy, hz;
};

// squared distance from the query to point p / to box [lo,hi]; PER = query ball may cross the box faces
template <bool PER> __device__ __forceinline__ float knn_d2(const KnnQuery &q, const float4 &p)
{
	if (PER)
		return dist2_rn(minimg_dx(q.x0, q.xp, q.xm, q.hx, p.x), minimg_dx(q.y0, q.yp, q.ym, q.hy, p.y),
		                minimg_dx(q.z0, q.zp, q.zm, q.hz, p.z));
	return dist2_rn(__fsub_rn(q.x0, p.x), __fsub_rn(q.y0, p.y), __fsub_rn(q.z0, p.z));
}
template <bool PER> __device__ __forceinline__ float knn_box_d2(const KnnQuery &q, const float4 &lo, const float4 &hi)
{
	if (PER)
		return dist2_rn(axis_gap_periodic(q.x0, q.xp, q.xm, lo.x, hi.x), axis_gap_periodic(q.y0, q.yp, q.ym, lo.y, hi.y),
		                axis_gap_periodic(q.z0, q.zp, q.zm, lo.z, hi.z));
	return dist2_rn(axis_gap(q.x0, lo.x, hi.x), axis_gap(q.y0, lo.y, hi.y), axis_gap(q.z0, lo.z, hi.z));
}

// Stage the lanes with `hit` (candidate idx at squared distance d2) in the warp's 64-entry buffer; 32 staged
// candidates are merged into the k-best at once and the bound follows the k-th best.
template <int R>
__device__ __forceinline__ void knn_stage(bool hit, float d2, int idx, uint64_t *s_buf, int lane, int k, uint64_t (&a)[R],
                                          int &cnt, float &bound)
{
	const uint32_t hm = __ballot_sync(SK_FULL, hit);
	if (!hm) return;
	if (hit) s_buf[cnt + __popc(hm & ((1u << lane) - 1u))] = ((uint64_t)__float_as_uint(d2) << 32) | (uint32_t)idx;
	cnt += __popc(hm);
	__syncwarp();
	if (cnt >= 32) {
		const uint64_t b = s_buf[lane];
		kbest_merge<R>(a, b, lane);
		const uint64_t t = (lane + 32 < cnt) ? s_buf[lane + 32] : KNN_INF;
		__syncwarp();
		s_buf[lane] = t;
		__syncwarp();
		cnt -= 32;
		bound = fminf(bound, kbest_kth<R>(a, k));
	}
}

// Tree walk of one query (one warp).  Buckets [skipLo, skipHi] were merged before the walk.
template <bool PER, int R>
__device__ __forceinline__ void knn_walk(const KnnArgs &a, const KnnQuery &q, uint64_t *s_buf, float (*s_dist)[32],
                                         int lane, int skipLo, int skipHi, uint64_t (&best)[R], int &cnt, float &bound)
{
	const int n = a.n, k = a.k;
	int lev = a.tv.top - 1;
	uint32_t node = 0;
	uint32_t mymask = 0;
	// (Measured and dropped: testing the boxes of a bucket's four 8-point runs instead of its one box, to visit fewer
	// buckets - the 4 x strided box loads and tests per level-0 node cost far more than the visits they save,
	// 99 -> 226 ms.)
	// Leaf buckets (children of a level-0 node) are visited nearest box first: the bound tightens on the
	// buckets that hold the true neighbours, and once the nearest unvisited box is outside the bound the
	// whole node is done.  key0 = (box distance bits, child) of this lane's child, ~0 when visited/outside.
	uint32_t key0 = 0xffffffffu;
#define KNN_TEST_CHILDREN()                                                                            \
	{                                                                                              \
		const float4 *bx = a.tv.box[lev] + 2 * ((size_t)node * 32 + lane);                     \
		float d = knn_box_d2<PER>(q, bx[0], bx[1]);                                            \
		if (lev == 0) key0 = d <= bound ? ((__float_as_uint(d) & ~31u) | (uint32_t)lane) : 0xffffffffu; \
		else {                                                                                 \
			s_dist[lev][lane] = d;                                                         \
			uint32_t m_ = __ballot_sync(SK_FULL, d <= bound);                              \
			if (lane == lev) mymask = m_;                                                  \
			__syncwarp();                                                                  \
		}                                                                                      \
	}
	KNN_TEST_CHILDREN();
	while (true) {
		if (lev == 0) {
			const uint32_t kmin = __reduce_min_sync(SK_FULL, key0);
			// (distance rounded down to a multiple of 32 ulp: never prunes a box that is inside the bound)
			if (kmin == 0xffffffffu || __uint_as_float(kmin & ~31u) > bound) {
				++lev;
				if (lev >= a.tv.top) break;
				node >>= 5;
				continue;
			}
			const int c = (int)(kmin & 31u);
			if (lane == c) key0 = 0xffffffffu;
			const uint32_t child = node * 32 + c;
			if ((int)child >= skipLo && (int)child <= skipHi) continue;
			// leaf bucket: 32 points, one per lane
			int idx = (int)child * 32 + lane;
			bool valid = idx < n;
			float4 p = a.pos4[valid ? idx : 0];
			float d2 = knn_d2<PER>(q, p);
			knn_stage<R>(valid && d2 <= bound, d2, idx, s_buf, lane, k, best, cnt, bound);
			continue;
		}
		// upper levels: nearest child box first as well, so that the walk starts in the subtree that holds
		// the query; when the nearest unvisited child is outside the bound the node is done
		const uint32_t m = __shfl_sync(SK_FULL, mymask, lev);
		const uint32_t kq = ((m >> lane) & 1u) ? ((__float_as_uint(s_dist[lev][lane]) & ~31u) | (uint32_t)lane) : 0xffffffffu;
		const uint32_t kmin = __reduce_min_sync(SK_FULL, kq);
		if (kmin == 0xffffffffu || __uint_as_float(kmin & ~31u) > bound) {
			++lev;
			if (lev >= a.tv.top) break;
			node >>= 5;
			continue;
		}
		const int c = (int)(kmin & 31u);
		if (lane == lev) mymask = m & ~(1u << c);
		--lev;
		node = node * 32 + c;
		KNN_TEST_CHILDREN();
	}
#undef KNN_TEST_CHILDREN
}

// One warp takes KNN_QPW consecutive queries.  From the second one on, the k neighbours of the previous query
// (a Morton neighbour, typically a fraction of the ball radius away) are k distinct candidates, so the largest of
// their distances to the new query is an upper bound of its k-th distance before anything else is known: the
// phase-A buckets and the walk then only stage what can still enter, and a query needs ~k/32 + 1 merges instead
// of ~10 (the merges were 39 % of the kernel's instructions, profiles/r01_v6_knn_2e22_lines.txt).
// (61 registers, 4 blocks per SM.  Measured: 51 registers / 5 blocks 68.0 ms, 42 registers / 6 blocks with spills
// 69.9 ms, against 67.7 ms - occupancy is not what limits it; 4, 16 or 32 queries per warp: 67.9 / 68.6 / 70.0 ms.)
template <int R> __global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_density(const KnnArgs a)
{
	__shared__ uint64_t s_buf[KNN_WARPS][64];
	__shared__ float s_dist[KNN_WARPS][SK_MAXLEV][32];
	const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int n = a.n, k = a.k;
	const int nB = (n + 31) >> 5;
	uint64_t best[R];
	bool havePrev = false;
	const long long q0 = (long long)a.qLo + ((long long)blockIdx.x * KNN_WARPS + w) * KNN_QPW;
	for (int t = 0; t < KNN_QPW; ++t) {
		const long long qll = q0 + t;
		if (qll >= a.qHi) break;
		const int qi = (int)qll;
		const float4 qp = a.pos4[qi];
		KnnQuery q;
		q.x0 = qp.x;
		q.y0 = qp.y;
		q.z0 = qp.z;
		q.xp = __fadd_rn(qp.x, a.L[0]);
		q.xm = __fsub_rn(qp.x, a.L[0]);
		q.yp = __fadd_rn(qp.y, a.L[1]);
		q.ym = __fsub_rn(qp.y, a.L[1]);
		q.zp = __fadd_rn(qp.z, a.L[2]);
		q.zm = __fsub_rn(qp.z, a.L[2]);
		q.hx = a.hL[0];
		q.hy = a.hL[1];
		q.hz = a.hL[2];

		float bound = __uint_as_float(0x7f800000u);
		if (havePrev) {
			uint32_t mb = 0u; // non-negative floats order like their bit patterns
#pragma unroll
			for (int r = 0; r < R; ++r)
				if (r * 32 + lane < k) mb = max(mb, __float_as_uint(knn_d2<true>(q, a.pos4[(uint32_t)best[r]])));
			bound = __uint_as_float(__reduce_max_sync(SK_FULL, mb));
		}
#pragma unroll
		for (int r = 0; r < R; ++r) best[r] = KNN_INF;
		int cnt = 0;
		// Phase A: the query's own bucket and its Morton neighbours (R + 1 buckets >= k candidates) hold most of
		// the k nearest.
		int b0 = (qi >> 5) - R / 2;
		if (b0 > nB - (R + 1)) b0 = nB - (R + 1);
		if (b0 < 0) b0 = 0;
		const int b1 = b0 + R < nB - 1 ? b0 + R : nB - 1;
		for (int b = b0; b <= b1; ++b) {
			const int idx = b * 32 + lane;
			const bool valid = idx < n;
			const float4 p = a.pos4[valid ? idx : 0];
			const float d2 = knn_d2<true>(q, p);
			knn_stage<R>(valid && d2 <= bound, d2, idx, s_buf[w], lane, k, best, cnt, bound);
		}
		// Phase B: walk the tree for everything else inside the bound.  If the ball cannot reach a face
		// of the periodic box no image can be closer than the point itself: plain differences are then
		// bit-identical to the min-image ones and much cheaper.
		const float r0 = sqrtf(bound) * 1.000001f;
		const bool per = !(q.x0 - r0 >= a.boxLo[0] && q.x0 + r0 <= a.boxHi[0] && q.y0 - r0 >= a.boxLo[1] &&
		                   q.y0 + r0 <= a.boxHi[1] && q.z0 - r0 >= a.boxLo[2] && q.z0 + r0 <= a.boxHi[2]);
		if (per) knn_walk<true, R>(a, q, s_buf[w], s_dist[w], lane, b0, b1, best, cnt, bound);
		else knn_walk<false, R>(a, q, s_buf[w], s_dist[w], lane, b0, b1, best, cnt, bound);
		if (cnt > 0) {
			const uint64_t b = (lane < cnt) ? s_buf[w][lane] : KNN_INF;
			kbest_merge<R>(best, b, lane);
		}
		__syncwarp();
		havePrev = true;
		const float fBall2 = kbest_kth<R>(best, k);
		if (lane == 0) a.ball2[qi] = fBall2;

		// ---- density (smooth1.c:249-263): every PQ entry except the farthest (pqHead)
		const float ih2 = __fdiv_rn(4.0f, fBall2);                                          // (float)(4.0/h2)
		const float fNorm = (float)(0.5 * 0.318309886183790671538 * sqrt((double)ih2) * (double)ih2);
		const float mi = qp.w;
		double gsum = 0.0;
#pragma unroll
		for (int r = 0; r < R; ++r) {
			const int e = r * 32 + lane;
			const uint64_t ent = best[r];
			const int j = (int)(uint32_t)ent;
			const float key = __uint_as_float((uint32_t)(ent >> 32));
			if (e < k - 1) {
				float r2 = __fmul_rn(key, ih2);
				float rs = (float)(2.0 - sqrt((double)r2));
				if (r2 < 1.0f) rs = (float)(1.0 - 0.75 * (double)rs * (double)r2);
				else rs = (float)(0.25 * (double)rs * (double)rs * (double)rs);
				rs = __fmul_rn(rs, fNorm);
				float mj = a.pos4[j].w;
				gsum += (double)__fmul_rn(rs, mj);
				atomicAdd(&a.rho64[j], (double)__fmul_rn(rs, mi));
			}
			if (a.nbr && e < k) {
				size_t row = (size_t)a.iord[qi] * k + e;
				a.nbr[row] = a.iord[j];
				a.nbrD2[row] = key;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(SK_FULL, gsum, o);
		if (lane == 0) atomicAdd(&a.rho64[qi], gsum);
	}
}

__global__ void __launch_bounds__(256) k_density_finish(int m, const int *iordA, const double *rho64,
                                                        const float *ball2A, float *rhoA, float *rho,
                                                        float *ball2)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	float r = (float)rho64[i];
	rhoA[i] = r;
	int j = iordA[i];
	rho[j] = r;
	ball2[j] = ball2A[i];
}

// ------------------------------------------------------------------ replicas (smooth1.c:278-332)
// INTERSECTNP (kd.h:102-119) against the periodic box, float32, left-to-right accumulation.
__device__ __forceinline__ float box_dist2_np(float x, float y, float z, const float *lo, const float *hi)
{
	float dx = __fsub_rn(lo[0], x), dx1 = __fsub_rn(x, hi[0]);
	float dy = __fsub_rn(lo[1], y), dy1 = __fsub_rn(y, hi[1]);
	float dz = __fsub_rn(lo[2], z), dz1 = __fsub_rn(z, hi[2]);
	float d2;
	if (dx > 0.0f) d2 = __fmul_rn(dx, dx);
	else if (dx1 > 0.0f) d2 = __fmul_rn(dx1, dx1);
	else d2 = 0.0f;
	if (dy > 0.0f) d2 = __fadd_rn(d2, __fmul_rn(dy, dy));
	else if (dy1 > 0.0f) d2 = __fadd_rn(d2, __fmul_rn(dy1, dy1));
	if (dz > 0.0f) d2 = __fadd_rn(d2, __fmul_rn(dz, dz));
	else if (dz1 > 0.0f) d2 = __fadd_rn(d2, __fmul_rn(dz1, dz1));
	return d2;
}

struct RepArgs {
	int m;
	const float4 *posA;
	const float *ball2A;
	const float *rhoA;
	float L[3], lo[3], hi[3];
};

// fNorm of the gradient kernel (smooth1.c:447-448): (float)(M_1_PI*ih2*ih2*sqrt(ih2)*fMass)
__device__ __forceinline__ float grad_norm(float ball2, float mass)
{
	float ih2 = __fdiv_rn(4.0f, ball2);
	return (float)(0.318309886183790671538 * (double)ih2 * (double)ih2 * sqrt((double)ih2) * (double)mass);
}

// mode 0: count replicas per particle into cnt[i]; mode 1: write originals + replicas.
template <int MODE>
__global__ void __launch_bounds__(256)
    k_replicas(const RepArgs a, uint32_t *cnt, const uint32_t *scan, float4 *entPos, float4 *entNR,
               uint32_t *entSrc, float *ex, float *ey, float *ez, float *einfl)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.m) return;
	float4 p = a.posA[i];
	float b2 = a.ball2A[i];
	uint32_t c = 0;
	uint32_t o = 0;
	float fn = 0.0f, rho = 0.0f, infl = 0.0f;
	if (MODE == 1) {
		fn = grad_norm(b2, p.w);
		rho = a.rhoA[i];
		infl = __fmul_ru(__fsqrt_ru(b2), 1.000001f);
		// original
		entPos[i] = make_float4(p.x, p.y, p.z, b2);
		entNR[i] = make_float4(__fdiv_rn(4.0f, b2), fn, rho, 0.0f);
		entSrc[i] = (uint32_t)i;
		ex[i] = p.x;
		ey[i] = p.y;
		ez[i] = p.z;
		einfl[i] = infl;
		o = (uint32_t)a.m + scan[i];
	}
	for (int ix = -1; ix <= 1; ++ix) {
		float x = __fadd_rn(p.x, __fmul_rn((float)ix, a.L[0]));
		for (int iy = -1; iy <= 1; ++iy) {
			float y = __fadd_rn(p.y, __fmul_rn((float)iy, a.L[1]));
			for (int iz = -1; iz <= 1; ++iz) {
				float z = __fadd_rn(p.z, __fmul_rn((float)iz, a.L[2]));
				if (ix || iy || iz) {
					float d2 = box_dist2_np(x, y, z, a.lo, a.hi);
					if (d2 < b2) {
						if (MODE == 1) {
							entPos[o] = make_float4(x, y, z, b2);
							entNR[o] = make_float4(__fdiv_rn(4.0f, b2), fn, rho, 0.0f);
							entSrc[o] = (uint32_t)i | 0x80000000u;
							ex[o] = x;
							ey[o] = y;
							ez[o] = z;
							einfl[o] = infl;
							++o;
						}
						++c;
					}
				}
			}
		}
	}
	if (MODE == 0) cnt[i] = c;
}

__global__ void __launch_bounds__(256)
    k_gather_entities(int m, const uint32_t *perm, const float4 *posU, const float4 *nrU, const uint32_t *srcU,
                      const float *inflU, float4 *pos, float4 *nr, uint32_t *src, float *infl, float *rho,
                      float4 *rec)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t j = perm[i];
	float4 pp = posU[j];
	pos[i] = pp;
	float4 v = nrU[j];
	nr[i] = v;
	rec[2 * (size_t)i] = pp;
	rec[2 * (size_t)i + 1] = v;
	src[i] = srcU[j];
	infl[i] = inflU[j];
	rho[i] = v.z;
}

__global__ void k_pad_entities(int ne, float4 *pos, float4 *aux)
{
	int i = ne + threadIdx.x;
	if (threadIdx.x < 64) {
		pos[i] = make_float4(3.0e38f, 3.0e38f, 3.0e38f, -1.0f);
		aux[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
}

// ------------------------------------------------------------------ host side of the stage
void stage_density(skidgpu_ctx &c, int nSmooth, int bGasAndDark, int bGasOnly, int *nExtraScat)
{
	cudaStream_t s = c.stream;
	const int n = c.n;
	if (n <= 0) throw SkidError("skidgpu_density: no particles set");
	if (nSmooth < 1 || nSmooth > 256) throw SkidError("skidgpu_density: nSmooth must be in [1,256] (the k best live in 2, 4 or 8 registers per lane)");
	StageTimer tm(c, 0);
	c.nSmooth = nSmooth;
	c.bGasAndDark = bGasAndDark;
	c.bGasOnly = bGasOnly;

	// scatter-active set (kdScatterActive)
	uint32_t *flags = c.flags.alloc(n);
	uint32_t *scan = c.scan.alloc((size_t)n + 64);
	SK_LAUNCH(k_scatter_flags, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.nGas, c.nDark, c.inType, bGasAndDark,
	          bGasOnly, flags);
	exclusive_scan_u32(flags, scan, n, c.ws, s);
	uint32_t nAct = 0;
	CK(cudaMemcpyAsync(&nAct, scan + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	c.nAct = (int)nAct;
	CK(cudaMemsetAsync(c.rho.alloc(n), 0, sizeof(float) * n, s));
	CK(cudaMemsetAsync(c.ball2.alloc(n), 0, sizeof(float) * n, s));
	c.nEnt = 0;
	c.nExtra = 0;
	if (nExtraScat) *nExtraScat = 0;
	if (c.keepNbr) {
		CK(cudaMemsetAsync(c.nbr.alloc((size_t)n * nSmooth), 0xff, sizeof(int) * (size_t)n * nSmooth, s));
		CK(cudaMemsetAsync(c.nbrD2.alloc((size_t)n * nSmooth), 0xff, sizeof(float) * (size_t)n * nSmooth, s));
	}
	if (c.nAct == 0) { // kd.c:379-383: no tree, densities stay 0
		tm.stop();
		return;
	}
	if (nSmooth > c.nAct) throw SkidError("skidgpu_density: nSmooth > number of scatter-active particles (smooth1.c:12)");
	const int m = c.nAct;
	uint32_t *actIdx = c.actIdx.alloc(m);
	SK_LAUNCH(k_compact_idx, (unsigned)ceil_div(n, 256), 256, 0, s, n, flags, scan, actIdx);
	float *gx = c.ax_.alloc(m), *gy = c.ay_.alloc(m), *gz = c.az_.alloc(m);
	SK_LAUNCH(k_gather3, (unsigned)ceil_div(m, 256), 256, 0, s, m, actIdx, c.x.p, c.y.p, c.z.p, gx, gy, gz);

	// tree over the active set (kdBuildTree)
	tree_sort_points(c.treeA, gx, gy, gz, m, c.ws, s, &c);
	float4 *posA = c.posA.alloc(m);
	int *iordA = c.iordA.alloc(m);
	SK_LAUNCH(k_gather_sortedA, (unsigned)ceil_div(m, 256), 256, 0, s, m, c.treeA.perm.p, actIdx, c.x.p, c.y.p,
	          c.z.p, c.mass.p, posA, iordA);
	tree_build_boxes(c.treeA, posA, nullptr, nullptr, m, s);

	// kNN + density
	KnnArgs ka;
	ka.tv = tree_view(c.treeA);
	ka.pos4 = posA;
	ka.iord = iordA;
	ka.n = m;
	ka.k = nSmooth;
	for (int d = 0; d < 3; ++d) {
		ka.L[d] = c.L[d];
		ka.hL[d] = 0.5f * c.L[d];
		ka.boxLo[d] = c.bPeriodic ? (float)((double)c.C[d] - 0.5 * (double)c.L[d]) : -3.4e38f;
		ka.boxHi[d] = c.bPeriodic ? (float)((double)c.C[d] + 0.5 * (double)c.L[d]) : 3.4e38f;
	}
	// multi-GPU: queries are sharded by contiguous Morton ranges of equal length; fBall2 has one writer per entry
	// (all-gather), the f64 density partials scatter onto neighbours of other ranges (all-reduce sum), so every
	// rank ends with identical arrays
	const int chunk = (int)ceil_div(m, c.nranks);
	ka.ball2 = c.ball2A.alloc((size_t)chunk * c.nranks);
	ka.rho64 = c.rho64A.alloc(m);
	ka.nbr = c.keepNbr ? c.nbr.p : nullptr;
	ka.nbrD2 = c.keepNbr ? c.nbrD2.p : nullptr;
	CK(cudaMemsetAsync(ka.rho64, 0, sizeof(double) * m, s));
	ka.qLo = std::min(m, chunk * c.rank);
	ka.qHi = std::min(m, chunk * (c.rank + 1));
	c.kernel_ms[KF_KNN] = 0;
	c.kernel_launches[KF_KNN] = 0;
	c.spans.begin(KF_KNN, s);
	if (ka.qHi > ka.qLo) {
		const unsigned grid = (unsigned)ceil_div(ka.qHi - ka.qLo, KNN_WARPS * KNN_QPW);
		if (nSmooth <= 64) SK_LAUNCH(k_knn_density<2>, grid, KNN_WARPS * 32, 0, s, ka);
		else if (nSmooth <= 128) SK_LAUNCH(k_knn_density<4>, grid, KNN_WARPS * 32, 0, s, ka);
		else SK_LAUNCH(k_knn_density<8>, grid, KNN_WARPS * 32, 0, s, ka);
	}
	c.spans.end(s);
	sk_allgather(c, ka.ball2, chunk, SK_F32);
	sk_reduce(c, ka.rho64, m, SK_F64, SK_SUM);
	c.nQueries += ka.qHi - ka.qLo;
	float *rhoA = c.rhoA.alloc(m);
	SK_LAUNCH(k_density_finish, (unsigned)ceil_div(m, 256), 256, 0, s, m, iordA, ka.rho64, ka.ball2, rhoA, c.rho.p,
	          c.ball2.p);

	// replicas + scatterer entities
	RepArgs ra;
	ra.m = m;
	ra.posA = posA;
	ra.ball2A = ka.ball2;
	ra.rhoA = rhoA;
	for (int d = 0; d < 3; ++d) {
		ra.L[d] = c.L[d];
		ra.lo[d] = (float)((double)c.C[d] - 0.5 * (double)c.L[d]); // smooth1.c:284-285
		ra.hi[d] = (float)((double)c.C[d] + 0.5 * (double)c.L[d]);
	}
	uint32_t nExtra = 0;
	if (c.bPeriodic) {
		SK_LAUNCH(k_replicas<0>, (unsigned)ceil_div(m, 256), 256, 0, s, ra, flags, nullptr, nullptr, nullptr,
		          nullptr, nullptr, nullptr, nullptr, nullptr);
		exclusive_scan_u32(flags, scan, m, c.ws, s);
		CK(cudaMemcpyAsync(&nExtra, scan + m, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
	} else {
		// not periodic: no replicas (smooth1.c:288); reuse the write kernel with a zero scan and an
		// unreachable period so that nothing qualifies
		CK(cudaMemsetAsync(scan, 0, sizeof(uint32_t) * (m + 1), s));
	}
	c.nExtra = (int)nExtra;
	c.nEnt = m + c.nExtra;
	if (nExtraScat) *nExtraScat = c.nExtra;
	const int ne = c.nEnt;
	float4 *posU = c.entPosU.alloc(ne);
	float4 *nrU = c.entNRU.alloc(ne);
	uint32_t *srcU = c.entSrcU.alloc(ne);
	float *ex = c.ex.alloc(ne), *ey = c.ey.alloc(ne), *ez = c.ez.alloc(ne), *einfl = c.eInfl.alloc(ne);
	if (!c.bPeriodic)
		for (int d = 0; d < 3; ++d) { // make every shifted copy miss: box = a point far away
			ra.lo[d] = 3.0e38f;
			ra.hi[d] = 3.0e38f;
			ra.L[d] = 0.0f;
		}
	SK_LAUNCH(k_replicas<1>, (unsigned)ceil_div(m, 256), 256, 0, s, ra, nullptr, scan, posU, nrU, srcU, ex, ey, ez,
	          einfl);
	// (Sorting the scatterers by (size class of the ball, curve index) so that a bucket holds balls of similar size
	// and its inflated box hugs them was measured at 2^24: list builds 80 -> 109 ms, tile step 149 -> 156 ms - a
	// tile's neighbourhood then lies in as many subtrees as there are size classes.  Position only.)
	tree_sort_points(c.treeE, ex, ey, ez, ne, c.ws, s, &c);
	float4 *ep = c.entPos.alloc(ne + 64);
	float4 *enr = c.entNR.alloc(ne + 64);
	uint32_t *esrc = c.entSrc.alloc(ne);
	float *inflS = c.tmpx.alloc(ne);
	float *rhoS = c.eRhoSorted.alloc((size_t)ne + 64); // compact copy of rhoEff: read beside the positions by the list walks
	float4 *erec = c.entRec.alloc(2 * ((size_t)ne + 64));
	SK_LAUNCH(k_gather_entities, (unsigned)ceil_div(ne, 256), 256, 0, s, ne, c.treeE.perm.p, posU, nrU, srcU, einfl, ep,
	          enr, esrc, inflS, rhoS, erec);
	// pad the sorted scatterer arrays to whole leaves with dummies that can never be hit (fBall2 = -1)
	SK_LAUNCH(k_pad_entities, 1, 64, 0, s, ne, ep, enr);
	CK(cudaMemsetAsync(rhoS + ne, 0, sizeof(float) * 64, s));
	tree_build_boxes(c.treeE, ep, inflS, rhoS, ne, s, 32, 32);
	CK(cudaMemsetAsync(c.entTouched.alloc(ne + 64), 0, ne + 64, s));
	tm.stop();
	c.spans.resolve(c.kernel_ms, c.kernel_launches);
}
