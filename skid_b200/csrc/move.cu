// Stage 3: density-gradient "flow" of the moving particles until they converge.
//
// Replaces kdInitMove/CutCriterion (kd.c:555-666), kdBuildMoveTree (kd.c:463-552),
// smBallGather (smooth1.c:338-384), smAccDensity + ScatterCut (smooth1.c:387-518),
// kdMoveParticles (kd.c:702-732), kdPruneInactive (kd.c:735-793), kdReactivateMove (kd.c:796)
// and the loops of main.c:394-419,431-438.
//
// The reference rebuilds a kd-tree over the MOVERS every step and lets every fixed scatterer
// scatter grad(W) onto the movers inside its ball.  Here the form is inverted: the scatterers
// (originals + explicit periodic replicas) never move, so ONE static tree with ball-inflated boxes
// is built once (density.cu) and every mover gathers from the scatterers whose ball contains it.
// Same set of (scatterer, mover) interactions, same float32 hit test, no per-step tree build.
// Scatterer pruning state of the reference is reproduced with two numbers per step:
// T (entities with rho < T are gone) and the per-entity "cut at step 0" flag (rhoEff = 0).
#include "ctx.cuh"
#include <utility>

#define T_NONE 0x7f800000u // +inf bits: "no entity was hit this step"


// CutCriterion (kd.c:555-597)
__global__ void __launch_bounds__(256)
    k_mover_flags(int n, int nGas, int nDark, int inType, int bGasAndDark, int bGasOnly, const float *rho,
                  const float *temp, const float *mass, float fDensMin, float fTempMax, float fMassMax,
                  uint32_t *flags)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int t = i < nGas ? SKIDGPU_GAS : (i < nGas + nDark ? SKIDGPU_DARK : SKIDGPU_STAR);
	int f = 0;
	float d = rho[i];
	if (!(mass[i] > fMassMax)) {
		switch (inType) {
		case SKIDGPU_DARK: f = d >= fDensMin; break;
		case SKIDGPU_GAS:
		case SKIDGPU_DARK | SKIDGPU_GAS:
			if (bGasAndDark && t == SKIDGPU_DARK && d >= fDensMin) f = 1;
			if (t == SKIDGPU_GAS && d >= fDensMin && temp[i] <= fTempMax) f = 1;
			break;
		case SKIDGPU_STAR:
		case SKIDGPU_DARK | SKIDGPU_STAR: f = (t == SKIDGPU_STAR); break;
		case SKIDGPU_GAS | SKIDGPU_STAR:
		case SKIDGPU_DARK | SKIDGPU_GAS | SKIDGPU_STAR:
			if (bGasAndDark && t == SKIDGPU_DARK && d >= fDensMin) f = 1;
			if (t == SKIDGPU_GAS) {
				if (d >= fDensMin && temp[i] <= fTempMax) f = 1;
			} else if (t == SKIDGPU_STAR && !bGasOnly) f = 1;
			break;
		}
	}
	flags[i] = (uint32_t)f;
}

__global__ void __launch_bounds__(256) k_compact_idx2(int n, const uint32_t *flags, const uint32_t *scan,
                                                      uint32_t *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flags[i]) out[scan[i]] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_gather3b(int m, const uint32_t *idx, const float *x, const float *y,
                                                  const float *z, float *ox, float *oy, float *oz)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t j = idx[i];
	ox[i] = x[j];
	oy[i] = y[j];
	oz[i] = z[j];
}

// movers in Morton order: r = rOld = initial position, mOrd = iOrder (kd.c:653-662)
__global__ void __launch_bounds__(256)
    k_init_movers(int m, const uint32_t *perm, const uint32_t *fileIdx, const float *x, const float *y,
                  const float *z, float *mx, float *my, float *mz, float *rox, float *roy, float *roz, int *mOrd,
                  const float *ball2, float *lhmin, int *lcnt, float initFactor)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t j = fileIdx[perm[i]];
	float px = x[j], py = y[j], pz = z[j];
	mx[i] = px;
	my[i] = py;
	mz[i] = pz;
	rox[i] = px;
	roy[i] = py;
	roz[i] = pz;
	mOrd[i] = (int)j;
	lhmin[i] = initFactor * sqrtf(fmaxf(ball2[j], 0.0f)); // first margin: a fraction of the mover's own ball radius
	lcnt[i] = -1;
}

__global__ void __launch_bounds__(256) k_iota(int lo, int cnt, uint32_t *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < cnt) out[i] = (uint32_t)(lo + i);
}

struct StepArgs {
	TreeView tv;
	const float4 *entPos; // (x,y,z,fBall2)
	const float4 *entNR;  // (4/fBall2, fNorm, rhoEff, 0)
	const float4 *entRec; // the two interleaved: rec[2e] = entPos[e], rec[2e+1] = entNR[e] (list path gathers)
	uint8_t *touched;     // nullable: set for entities with >= 1 hit (step 0, initial cut)
	float *mx, *my, *mz;
	const uint32_t *act;
	int nActive;
	int nEnt;
	uint32_t *dT; // [0] threshold T (float bits), [1] running min of rho over hit entities (float bits)
	float fStep;
	float L[3];
	double wrapLo[3], wrapHi[3];
	float *a0x, *a0y, *a0z; // nullable: keep accelerations
	// candidate lists (k_move_list): per mover LIST_CAP scatterer indices, list origin, margin,
	// count (-1 = no valid list), and the smallest containing ball radius seen at the last walk
	uint32_t *list;
	uint32_t listBase; // first mover id of this shard
	float *lx0, *ly0, *lz0, *ldelta, *lhmin;
	int *lcnt;
	int walkAlways; // debug (SKIDGPU_LIST_WALK_ALWAYS=1): never use the lists
	float polShrink, polGrow; // margin feedback (see k_move_list)
	int polGrowMax;
	uint32_t *queue;      // movers whose list must be refreshed this step
	uint32_t *queueCount; // = dT + 2, reset by k_update_T
};

constexpr int STEP_WARPS = 8;

__global__ void __launch_bounds__(STEP_WARPS * 32) k_move_step(const StepArgs a)
{
	const int lane = threadIdx.x & 31;
	const int wi = blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
	if (wi >= a.nActive) return;
	const uint32_t id = a.act[wi];
	const float x = a.mx[id], y = a.my[id], z = a.mz[id];
	const float T = __uint_as_float(a.dT[0]);
	float ax = 0.0f, ay = 0.0f, az = 0.0f;
	float rmin = 3.0e38f;

	int lev = a.tv.top - 1;
	uint32_t node = 0;
	uint32_t mymask = 0;
#define STEP_TEST_CHILDREN()                                                                           \
	{                                                                                              \
		const float4 *bx = a.tv.box[lev] + 2 * ((size_t)node * 32 + lane);                     \
		float4 lo = bx[0], hi = bx[1];                                                         \
		bool in_ = x >= lo.x && x <= hi.x && y >= lo.y && y <= hi.y && z >= lo.z && z <= hi.z && \
		           lo.w >= T;                                                                  \
		uint32_t m_ = __ballot_sync(SK_FULL, in_);                                             \
		if (lane == lev) mymask = m_;                                                          \
	}
	STEP_TEST_CHILDREN();
	while (true) {
		uint32_t m = __shfl_sync(SK_FULL, mymask, lev);
		if (m == 0) {
			++lev;
			if (lev >= a.tv.top) break;
			node >>= 5;
			continue;
		}
		int c = __ffs(m) - 1;
		m &= m - 1;
		if (lane == lev) mymask = m;
		uint32_t child = node * 32 + c;
		if (lev > 0) {
			--lev;
			node = child;
			STEP_TEST_CHILDREN();
			continue;
		}
		int e = (int)child * 32 + lane;
		if (e < a.nEnt) {
			float4 p = a.entPos[e];
			// smBallGather (smooth1.c:365-369): dx = x_scatterer - x_mover, float32, no FMA
			float dx = __fsub_rn(p.x, x), dy = __fsub_rn(p.y, y), dz = __fsub_rn(p.z, z);
			float d2 = dist2_rn(dx, dy, dz);
			if (d2 < p.w) {
				float4 nr = a.entNR[e];
				if (nr.z >= T) {
					// smAccDensity (smooth1.c:447-459)
					float r2 = __fmul_rn(d2, nr.x);
					float rs = __fsqrt_rn(r2);
					if (r2 < 1.0f) rs = (float)(-3.0 + 2.25 * (double)rs);
					else rs = (float)(-3.0 / (double)rs + 3.0 - 0.75 * (double)rs);
					rs = __fmul_rn(rs, nr.y);
					ax = __fadd_rn(ax, __fmul_rn(dx, rs));
					ay = __fadd_rn(ay, __fmul_rn(dy, rs));
					az = __fadd_rn(az, __fmul_rn(dz, rs));
					rmin = fminf(rmin, nr.z);
					if (a.touched) a.touched[e] = 1;
				}
			}
		}
	}
#undef STEP_TEST_CHILDREN
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		ax += __shfl_xor_sync(SK_FULL, ax, o);
		ay += __shfl_xor_sync(SK_FULL, ay, o);
		az += __shfl_xor_sync(SK_FULL, az, o);
		rmin = fminf(rmin, __shfl_xor_sync(SK_FULL, rmin, o));
	}
	if (lane == 0) {
		if (rmin < 3.0e38f) atomicMin(&a.dT[1], __float_as_uint(rmin)); // smooth1.c:460-461 (rho > 0)
		if (a.a0x) {
			a.a0x[id] = ax;
			a.a0y[id] = ay;
			a.a0z[id] = az;
		}
		// kdMoveParticles (kd.c:711-729)
		float s2 = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
		float ai = (float)sqrt((double)s2);
		if (ai > 0.0f) ai = (float)((double)a.fStep / sqrt((double)s2));
		else ai = 0.0f;
		float r[3] = {__fsub_rn(x, __fmul_rn(ai, ax)), __fsub_rn(y, __fmul_rn(ai, ay)),
		              __fsub_rn(z, __fmul_rn(ai, az))};
#pragma unroll
		for (int j = 0; j < 3; ++j) {
			if ((double)r[j] > a.wrapHi[j]) r[j] = __fsub_rn(r[j], a.L[j]);
			if ((double)r[j] <= a.wrapLo[j]) r[j] = __fadd_rn(r[j], a.L[j]);
		}
		a.mx[id] = r[0];
		a.my[id] = r[1];
		a.mz[id] = r[2];
	}
}

// kdMoveParticles (kd.c:711-729) for one mover
__device__ __forceinline__ void move_one(const StepArgs &a, uint32_t id, float x, float y, float z, float ax, float ay,
                                         float az)
{
	float s2 = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
	float ai = (float)sqrt((double)s2);
	if (ai > 0.0f) ai = (float)((double)a.fStep / sqrt((double)s2));
	else ai = 0.0f;
	float r[3] = {__fsub_rn(x, __fmul_rn(ai, ax)), __fsub_rn(y, __fmul_rn(ai, ay)), __fsub_rn(z, __fmul_rn(ai, az))};
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		if ((double)r[j] > a.wrapHi[j]) r[j] = __fsub_rn(r[j], a.L[j]);
		if ((double)r[j] <= a.wrapLo[j]) r[j] = __fadd_rn(r[j], a.L[j]);
	}
	a.mx[id] = r[0];
	a.my[id] = r[1];
	a.mz[id] = r[2];
}

// smAccDensity (smooth1.c:447-459) for one hit.  Same float32 operations as the reference
// (r2 = d2*ih2, rs = sqrt(r2), rs *= fNorm, a += dx*rs, all round-to-nearest, no FMA on the sums);
// the spline factor, which the reference evaluates in double and rounds to float, is formed with
// one FMA (inner branch, a single rounding of the exact value) or with a float division plus its
// exact-remainder correction (outer branch), i.e. to float accuracy without double arithmetic.
#define ACC_HIT(dx, dy, dz, d2, q)                                                                     \
{                                                                                              \
	const float r2_ = __fmul_rn((d2), (q).x);                                              \
	const float rs_ = __fsqrt_rn(r2_);                                                     \
	float g_;                                                                              \
	if (r2_ < 1.0f) g_ = fmaf(2.25f, rs_, -3.0f);                                          \
	else {                                                                                 \
		const float t_ = __fdiv_rn(-3.0f, rs_);                                        \
		const float c_ = __fdividef(fmaf(-t_, rs_, -3.0f), rs_);                       \
		g_ = fmaf(-0.75f, rs_, 3.0f + t_) + c_;                                        \
	}                                                                                      \
	g_ = __fmul_rn(g_, (q).y);                                                             \
	ax = __fadd_rn(ax, __fmul_rn((dx), g_));                                               \
	ay = __fadd_rn(ay, __fmul_rn((dy), g_));                                               \
	az = __fadd_rn(az, __fmul_rn((dz), g_));                                               \
	rmin = fminf(rmin, (q).z);                                                             \
}


// Default kernels: per-mover candidate lists ("Verlet lists").
//
// ncu on the v1 kernel above (profiles/r01_v1_move_*): issue-bound, ~3600 warp instructions per
// mover-step, of which ~85 % are tree bookkeeping and misses (1500 scatterers tested for 85 hits).
// A mover travels exactly fStep (= tau/4) per step while the balls that contain it have radii of
// many tau, so the set of scatterers that CAN contain it changes slowly.  Each mover keeps the list
// of scatterers e with |x_e - x0| < h_e + delta collected by one tree walk at x0; while it stays
// within delta of x0 every scatterer that contains it is in the list (triangle inequality), and a
// step only reads the list, gathers those scatterers and runs the SAME float32 hit test.
//
//  * k_list_eval   (every step, one warp per active mover, small register footprint -> full
//                   occupancy; it is latency bound): movers with a valid list take their step;
//                   the others are appended to a refresh queue.
//  * k_list_refresh(every step, one warp per queued mover): walks the tree, evaluating this step
//                   and emitting the new list in the same pass.  Leaf buckets are fetched two at a
//                   time to keep more loads in flight.
//
// Splitting the two keeps warps that run ~10x longer out of the blocks of the short ones.  The margin
// delta is feedback-controlled per mover (tools/sweep_policy.sh).  The hit set, the hit test and the
// pruning rule are exactly those of v1 (verified against it and against the reference: same groups,
// same Ittr trace).  A periodic wrap moves the mover by L, which fails the drift check by construction.
constexpr int LIST_CAP = 384;
constexpr int EVAL_WARPS = 4;
constexpr int REFRESH_WARPS = 4;

// the end of every step for one mover: min density of the scatterers that hit it, optional copy of
// the acceleration, kdMoveParticles
__device__ __forceinline__ void finish_step(const StepArgs &a, uint32_t id, float x, float y, float z, float ax,
                                            float ay, float az, float rmin, int lane)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		ax += __shfl_xor_sync(SK_FULL, ax, o);
		ay += __shfl_xor_sync(SK_FULL, ay, o);
		az += __shfl_xor_sync(SK_FULL, az, o);
		rmin = fminf(rmin, __shfl_xor_sync(SK_FULL, rmin, o));
	}
	if (lane == 0) {
		if (rmin < 3.0e38f) atomicMin(&a.dT[1], __float_as_uint(rmin)); // smooth1.c:460-461 (rho > 0)
		if (a.a0x) {
			a.a0x[id] = ax;
			a.a0y[id] = ay;
			a.a0z[id] = az;
		}
		move_one(a, id, x, y, z, ax, ay, az);
	}
}

__global__ void __launch_bounds__(EVAL_WARPS * 32, 16) k_list_eval(const StepArgs a)
{
	const int lane = threadIdx.x & 31;
	const int wi = blockIdx.x * EVAL_WARPS + (threadIdx.x >> 5);
	if (wi >= a.nActive) return;
	const uint32_t id = a.act[wi];
	const float x = a.mx[id], y = a.my[id], z = a.mz[id];
	const int cnt0 = a.lcnt[id];
	bool useList = false;
	if (cnt0 >= 0 && !a.walkAlways) {
		const float ox = x - a.lx0[id], oy = y - a.ly0[id], oz = z - a.lz0[id];
		const float dl = a.ldelta[id];
		useList = (ox * ox + oy * oy + oz * oz) * 1.0001f <= dl * dl;
	}
	if (!useList) { // drifted past the margin (or no list yet): the refresh kernel takes this step
		if (lane == 0) a.queue[atomicAdd(a.queueCount, 1u)] = id;
		return;
	}
	const float T = __uint_as_float(a.dT[0]);
	float ax = 0.0f, ay = 0.0f, az = 0.0f;
	float rmin = 3.0e38f;
	const uint32_t *list = a.list + (size_t)(id - a.listBase) * LIST_CAP;
	// both halves of a scatterer's 32-byte record (one L2 sector) are fetched together
	for (int s0 = 0; s0 < cnt0; s0 += 32) {
		const int s = s0 + lane;
		if (s < cnt0) {
			const uint32_t e = list[s];
			const float4 p = a.entRec[2 * (size_t)e];
			const float4 q = a.entRec[2 * (size_t)e + 1];
			// smBallGather (smooth1.c:365-369): dx = x_scatterer - x_mover, float32, no FMA
			const float dx = __fsub_rn(p.x, x), dy = __fsub_rn(p.y, y), dz = __fsub_rn(p.z, z);
			const float d2 = dist2_rn(dx, dy, dz);
			if (d2 < p.w && q.z >= T) ACC_HIT(dx, dy, dz, d2, q);
		}
	}
	finish_step(a, id, x, y, z, ax, ay, az, rmin, lane);
}

__global__ void __launch_bounds__(REFRESH_WARPS * 32) k_list_refresh(const StepArgs a)
{
	const int lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t wi = blockIdx.x * REFRESH_WARPS + (threadIdx.x >> 5);
	if (wi >= *a.queueCount) return;
	const uint32_t id = a.queue[wi];
	const float x = a.mx[id], y = a.my[id], z = a.mz[id];
	const float T = __uint_as_float(a.dT[0]);
	float ax = 0.0f, ay = 0.0f, az = 0.0f;
	float rmin = 3.0e38f;
	uint32_t *list = a.list + (size_t)(id - a.listBase) * LIST_CAP; // lists exist for this shard's movers only
	const float delta = fminf(fmaxf(a.lhmin[id], 2.0f * a.fStep), 64.0f * a.fStep); // lhmin = margin to use
	const float delta2 = delta * delta;
	int nHit = 0;
	int cnt = 0;
	bool overflow = false;
	// one scatterer per lane: hit test for this step, candidate test for the list.
	// candidate <=> d <= h + delta <=> u = d2 - h^2 - delta^2 <= 2 h delta (no square root; 1e-4 slack)
#define LIST_PROCESS(e_, p)                                                                            \
	{                                                                                              \
		const float dx = __fsub_rn(p.x, x), dy = __fsub_rn(p.y, y), dz = __fsub_rn(p.z, z);    \
		const float d2 = dist2_rn(dx, dy, dz);                                                 \
		const float u = d2 - p.w - delta2;                                                     \
		bool cand = p.w > 0.0f && (u <= 0.0f || u * u <= 4.0004f * p.w * delta2);              \
		if (cand) {                                                                            \
			const float4 q = a.entNR[e_];                                                  \
			cand = q.z >= T; /* dead scatterers never come back */                         \
			if (cand && d2 < p.w) {                                                        \
				ACC_HIT(dx, dy, dz, d2, q);                                            \
				++nHit;                                                                \
				if (a.touched) a.touched[e_] = 1;                                      \
			}                                                                              \
		}                                                                                      \
		const uint32_t cm = __ballot_sync(SK_FULL, cand);                                      \
		const int nc = __popc(cm);                                                             \
		if (cnt + nc <= LIST_CAP) {                                                            \
			if (cand) list[cnt + __popc(cm & lt)] = e_;                                    \
		} else overflow = true;                                                                \
		cnt += nc;                                                                             \
	}
	int lev = a.tv.top - 1;
	uint32_t node = 0, mymask = 0;
	const float bx0 = x - delta, bx1 = x + delta, by0 = y - delta, by1 = y + delta, bz0 = z - delta, bz1 = z + delta;
#define LIST_TEST_CHILDREN()                                                                           \
	{                                                                                              \
		const float4 *bx = a.tv.box[lev] + 2 * ((size_t)node * 32 + lane);                     \
		float4 lo = bx[0], hi = bx[1];                                                         \
		bool in_ = bx1 >= lo.x && bx0 <= hi.x && by1 >= lo.y && by0 <= hi.y && bz1 >= lo.z && bz0 <= hi.z && \
		           lo.w >= T;                                                                  \
		uint32_t m_ = __ballot_sync(SK_FULL, in_);                                             \
		if (lane == lev) mymask = m_;                                                          \
	}
	LIST_TEST_CHILDREN();
	while (true) {
		uint32_t m = __shfl_sync(SK_FULL, mymask, lev);
		if (m == 0) {
			++lev;
			if (lev >= a.tv.top) break;
			node >>= 5;
			continue;
		}
		int c = __ffs(m) - 1;
		m &= m - 1;
		if (lev > 0) {
			if (lane == lev) mymask = m;
			uint32_t child = node * 32 + c;
			--lev;
			node = child;
			LIST_TEST_CHILDREN();
			continue;
		}
		// leaf level: take two buckets per round so that two record loads are in flight
		const uint32_t e0 = (node * 32 + c) * 32 + lane; // arrays are padded with fBall2 = -1 dummies
		const float4 p0 = a.entPos[e0];
		if (m) {
			const int c1 = __ffs(m) - 1;
			m &= m - 1;
			const uint32_t e1 = (node * 32 + c1) * 32 + lane;
			const float4 p1 = a.entPos[e1];
			if (lane == 0) mymask = m;
			LIST_PROCESS(e0, p0);
			LIST_PROCESS(e1, p1);
		} else {
			if (lane == 0) mymask = m;
			LIST_PROCESS(e0, p0);
		}
	}
#undef LIST_TEST_CHILDREN
#undef LIST_PROCESS
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) nHit += __shfl_xor_sync(SK_FULL, nHit, o);
	if (lane == 0) {
		a.lcnt[id] = overflow ? -1 : cnt;
		a.lx0[id] = x;
		a.ly0[id] = y;
		a.lz0[id] = z;
		a.ldelta[id] = delta;
		// feedback on the margin: grow it while the list stays short relative to the hits, shrink it
		// when the list is long (a long list makes every step slower, a short one refreshes often)
		float next = delta;
		if (overflow || (float)cnt > a.polShrink * nHit + 32.0f) next = 0.6f * delta;
		else if ((float)cnt < a.polGrow * nHit + 32.0f && cnt < a.polGrowMax) next = 1.5f * delta;
		a.lhmin[id] = next;
	}
	finish_step(a, id, x, y, z, ax, ay, az, rmin, lane);
}

// After a step: adopt the new threshold (ScatterCut, smooth1.c:509-513).  If nothing was hit the
// reference's fScatDens stays 0.0 and nothing is cut.
__global__ void k_update_T(uint32_t *dT, int bNoPrune)
{
	uint32_t nx = dT[1];
	if (!bNoPrune && nx != T_NONE) dT[0] = nx;
	dT[1] = T_NONE;
	dT[2] = 0u; // refresh queue of the next step
}

// Initial cut (smooth1.c:463-470,500-507): entities that scattered onto nobody get fDensity = 0.
__global__ void __launch_bounds__(256) k_initial_cut(int nEnt, const uint8_t *touched, float4 *entNR, float4 *entRec)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < nEnt && !touched[e]) {
		entNR[e].z = 0.0f;
		entRec[2 * (size_t)e + 1].z = 0.0f;
	}
}

// nScatter of the log line = surviving originals + surviving replicas (smooth1.c:517)
__global__ void __launch_bounds__(256) k_count_scatter(int nEnt, const float4 *entNR, const uint32_t *dT,
                                                       uint32_t *out)
{
	float T = __uint_as_float(dT[0]);
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	bool alive = e < nEnt && entNR[e].z >= T;
	uint32_t b = __ballot_sync(SK_FULL, alive);
	if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (uint32_t)__popc(b));
}

__global__ void __launch_bounds__(256)
    k_alive_by_order(int nEnt, const float4 *entNR, const uint32_t *entSrc, const int *iordA, const uint32_t *dT,
                     uint8_t *alive)
{
	float T = __uint_as_float(dT[0]);
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nEnt) return;
	uint32_t s = entSrc[e];
	if (s & 0x80000000u) return;
	alive[iordA[s]] = entNR[e].z >= T ? 1 : 0;
}

// kdPruneInactive (kd.c:735-793): a mover stays active iff it moved >= fCvg (min image) since the
// last check.  flags -> scan -> stable compaction of the active list.
__global__ void __launch_bounds__(256)
    k_prune_flags(int nActive, const uint32_t *act, const float *mx, const float *my, const float *mz,
                  const float *rox, const float *roy, const float *roz, float hx, float hy, float hz, float fCvg2,
                  uint32_t *flags)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nActive) return;
	uint32_t id = act[i];
	float dx = __fsub_rn(mx[id], rox[id]);
	float dy = __fsub_rn(my[id], roy[id]);
	float dz = __fsub_rn(mz[id], roz[id]);
	float tx = __fmul_rn(2.0f, hx), ty = __fmul_rn(2.0f, hy), tz = __fmul_rn(2.0f, hz);
	if (dx > hx) dx = __fsub_rn(dx, tx);
	if (dx <= -hx) dx = __fadd_rn(dx, tx);
	if (dy > hy) dy = __fsub_rn(dy, ty);
	if (dy <= -hy) dy = __fadd_rn(dy, ty);
	if (dz > hz) dz = __fsub_rn(dz, tz);
	if (dz <= -hz) dz = __fadd_rn(dz, tz);
	float dr2 = dist2_rn(dx, dy, dz);
	flags[i] = dr2 >= fCvg2 ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
    k_prune_compact(int nActive, const uint32_t *act, const uint32_t *flags, const uint32_t *scan,
                    const float *mx, const float *my, const float *mz, float *rox, float *roy, float *roz,
                    uint32_t *actOut)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nActive || !flags[i]) return;
	uint32_t id = act[i];
	actOut[scan[i]] = id;
	rox[id] = mx[id];
	roy[id] = my[id];
	roz[id] = mz[id];
}

static void fill_step_args(skidgpu_ctx &c, StepArgs &sa, float fStep)
{
	sa.tv = tree_view(c.treeE);
	sa.entPos = c.entPos.p;
	sa.entNR = c.entNR.p;
	sa.entRec = c.entRec.p;
	sa.touched = nullptr;
	sa.mx = c.mx.p;
	sa.my = c.my.p;
	sa.mz = c.mz.p;
	sa.act = c.actList.p;
	sa.nActive = c.nActive;
	sa.nEnt = c.nEnt;
	sa.dT = c.dT.p;
	sa.fStep = fStep;
	for (int d = 0; d < 3; ++d) {
		sa.L[d] = c.L[d];
		sa.wrapHi[d] = (double)c.C[d] + 0.5 * (double)c.L[d]; // kd.c:724
		sa.wrapLo[d] = (double)c.C[d] - 0.5 * (double)c.L[d]; // kd.c:726
	}
	sa.a0x = sa.a0y = sa.a0z = nullptr;
	sa.list = c.mList.p;
	sa.listBase = (uint32_t)c.shardLo;
	sa.lx0 = c.lx0.p;
	sa.ly0 = c.ly0.p;
	sa.lz0 = c.lz0.p;
	sa.ldelta = c.ldelta.p;
	sa.lhmin = c.lhmin.p;
	sa.lcnt = c.lcnt.p;
	static int wa = -1;
	if (wa < 0) wa = getenv("SKIDGPU_LIST_WALK_ALWAYS") ? 1 : 0;
	sa.walkAlways = wa;
	static float pol[4] = {-1, 0, 0, 0};
	if (pol[0] < 0) {
		pol[0] = 3.0f;  // shrink when candidates > 3 x hits (+32)
		pol[1] = 1.5f;  // grow while candidates < 1.5 x hits (+32)
		pol[2] = 96.f;  // ... and fewer than this (measured sweep: tools/sweep_policy.sh)
		pol[3] = 0.3f;  // first margin = 0.3 x own ball radius
		if (const char *e = getenv("SKIDGPU_LIST_POLICY")) sscanf(e, "%f,%f,%f,%f", &pol[0], &pol[1], &pol[2], &pol[3]);
	}
	sa.polShrink = pol[0];
	sa.polGrow = pol[1];
	sa.polGrowMax = (int)pol[2];
	c.listInitFactor = pol[3];
	sa.queue = c.mQueue.p;
	sa.queueCount = c.dT.p ? c.dT.p + 2 : nullptr;
}

static int count_scatterers(skidgpu_ctx &c)
{
	cudaStream_t s = c.stream;
	uint32_t *cnt = c.dCount.alloc(4);
	CK(cudaMemsetAsync(cnt, 0, sizeof(uint32_t), s));
	if (c.nEnt > 0)
		SK_LAUNCH(k_count_scatter, (unsigned)ceil_div(c.nEnt, 256), 256, 0, s, c.nEnt, c.entNR.p, c.dT.p, cnt);
	uint32_t h = 0;
	CK(cudaMemcpyAsync(&h, cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	return (int)h;
}

// SKIDGPU_MOVE_KERNEL=warp selects the v1 kernel (a tree walk every step; kept for A/B
// measurements); default = candidate lists.
bool use_list_kernel();
bool use_list_kernel()
{
	static int v = -1;
	if (v < 0) {
		const char *e = getenv("SKIDGPU_MOVE_KERNEL");
		v = (e && !strcmp(e, "warp")) ? 0 : 1;
	}
	return v == 1;
}

static int one_step(skidgpu_ctx &c, StepArgs &sa, int bNoPrune)
{
	int launched = 0;
	if (c.nActive > 0 && c.nEnt > 0) {
		launched = 1;
		sa.act = c.actList.p;
		sa.nActive = c.nActive;
		if (use_list_kernel()) {
			SK_LAUNCH(k_list_eval, (unsigned)ceil_div(c.nActive, EVAL_WARPS), EVAL_WARPS * 32, 0, c.stream, sa);
			// grid sized for the worst case (everybody refreshes); warps beyond the queue length exit at once
			SK_LAUNCH(k_list_refresh, (unsigned)ceil_div(c.nActive, REFRESH_WARPS), REFRESH_WARPS * 32, 0, c.stream, sa);
		} else
			SK_LAUNCH(k_move_step, (unsigned)ceil_div(c.nActive, STEP_WARPS), STEP_WARPS * 32, 0, c.stream, sa);
		c.moverSteps += c.nActive;
	}
	if (!bNoPrune) sk_reduce(c, c.dT.p + 1, 1, SK_I32, SK_MIN); // fScatDens over all ranks' movers (+inf bits = none)
	SK_LAUNCH(k_update_T, 1, 1, 0, c.stream, c.dT.p, bNoPrune);
	return launched;
}

void stage_move(skidgpu_ctx &c, float fDensMin, float fTempMax, float fMassMax, float fCvg, float fStep,
                int bForceInitialCut, int bNoPrune, skidgpu_log_cb cb, void *user, int *nMoveOut, int *nIttrOut)
{
	cudaStream_t s = c.stream;
	const int n = c.n;
	if (n <= 0) throw SkidError("skidgpu_move: no particles set");
	if (!c.rho.p) throw SkidError("skidgpu_move: skidgpu_density has not run");
	StageTimer tm(c, 1);
	c.bNoPrune = bNoPrune;
	{
		StepArgs tmp;
		fill_step_args(c, tmp, fStep); // also reads the list policy (listInitFactor) from the environment
	}

	// ---- kdInitMove
	uint32_t *flags = c.flags.alloc(n);
	uint32_t *scan = c.scan.alloc((size_t)n + 64);
	SK_LAUNCH(k_mover_flags, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.nGas, c.nDark, c.inType, c.bGasAndDark,
	          c.bGasOnly, c.rho.p, c.temp.p, c.mass.p, fDensMin, fTempMax, fMassMax, flags);
	exclusive_scan_u32(flags, scan, n, c.ws, s);
	uint32_t nm = 0;
	CK(cudaMemcpyAsync(&nm, scan + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	c.nMove = (int)nm;
	c.haveCenters = false;
	const int m = c.nMove;
	uint32_t *dT = c.dT.alloc(4);
	uint32_t initT[3] = {0u, T_NONE, 0u}; // threshold, running min of this step, refresh-queue length
	CK(cudaMemcpyAsync(dT, initT, sizeof initT, cudaMemcpyHostToDevice, s));
	c.shardLo = (int)((long long)m * c.rank / c.nranks);
	c.shardHi = (int)((long long)m * (c.rank + 1) / c.nranks);
	c.nActive = c.shardHi - c.shardLo;
	if (m > 0) {
		uint32_t *fileIdx = c.actList2.alloc(m);
		SK_LAUNCH(k_compact_idx2, (unsigned)ceil_div(n, 256), 256, 0, s, n, flags, scan, fileIdx);
		float *gx = c.tmpx.alloc(m), *gy = c.tmpy.alloc(m), *gz = c.tmpz.alloc(m);
		SK_LAUNCH(k_gather3b, (unsigned)ceil_div(m, 256), 256, 0, s, m, fileIdx, c.x.p, c.y.p, c.z.p, gx, gy, gz);
		tree_sort_points(c.treeM, gx, gy, gz, m, c.ws, s);
		c.mx.alloc(m);
		c.my.alloc(m);
		c.mz.alloc(m);
		c.rox.alloc(m);
		c.roy.alloc(m);
		c.roz.alloc(m);
		c.mOrd.alloc(m);
		c.lx0.alloc(m);
		c.ly0.alloc(m);
		c.lz0.alloc(m);
		c.ldelta.alloc(m);
		c.lhmin.alloc(m);
		c.lcnt.alloc(m);
		if (use_list_kernel()) c.mList.alloc((size_t)(c.shardHi - c.shardLo > 0 ? c.shardHi - c.shardLo : 1) * LIST_CAP);
		c.mQueue.alloc(c.shardHi - c.shardLo > 0 ? c.shardHi - c.shardLo : 1);
		SK_LAUNCH(k_init_movers, (unsigned)ceil_div(m, 256), 256, 0, s, m, c.treeM.perm.p, fileIdx, c.x.p, c.y.p,
		          c.z.p, c.mx.p, c.my.p, c.mz.p, c.rox.p, c.roy.p, c.roz.p, c.mOrd.p, c.ball2.p, c.lhmin.p, c.lcnt.p, c.listInitFactor);
		c.actList.alloc(m);
		c.actList2.alloc(m); // fileIdx no longer needed after k_init_movers (same stream)
		if (c.nActive > 0)
			SK_LAUNCH(k_iota, (unsigned)ceil_div(c.nActive, 256), 256, 0, s, c.shardLo, c.nActive, c.actList.p);
	}
	if (nMoveOut) *nMoveOut = m;

	// ---- step 0 (main.c:396-404)
	const int bInitial = ((c.inType == SKIDGPU_DARK) || bForceInitialCut) && !bNoPrune;
	StepArgs sa;
	fill_step_args(c, sa, fStep);
	if (c.keepStep0 && m > 0) {
		sa.a0x = c.a0x.alloc(m);
		sa.a0y = c.a0y.alloc(m);
		sa.a0z = c.a0z.alloc(m);
		CK(cudaMemsetAsync(sa.a0x, 0, sizeof(float) * m, s));
		CK(cudaMemsetAsync(sa.a0y, 0, sizeof(float) * m, s));
		CK(cudaMemsetAsync(sa.a0z, 0, sizeof(float) * m, s));
	}
	if (bInitial && c.nEnt > 0) {
		CK(cudaMemsetAsync(c.entTouched.p, 0, c.nEnt, s));
		sa.touched = c.entTouched.p;
	}
	int nActiveLog = c.nActive;
	KernelTimer kt(c, 0);
	c.kernel_ms[0] = 0;
	c.kernel_launches[0] = 0;
	kt.start();
	kt.stop(one_step(c, sa, bNoPrune));
	if (bInitial && c.nEnt > 0) sk_reduce(c, c.entTouched.p, c.nEnt, SK_U8, SK_MAX);
	if (bInitial && c.nEnt > 0)
		SK_LAUNCH(k_initial_cut, (unsigned)ceil_div(c.nEnt, 256), 256, 0, s, c.nEnt, c.entTouched.p, c.entNR.p, c.entRec.p);
	sa.touched = nullptr;
	sa.a0x = sa.a0y = sa.a0z = nullptr;
	if (c.keepStep0 && c.nEnt > 0) {
		CK(cudaMemsetAsync(c.aliveByOrd.alloc(n), 0, n, s));
		SK_LAUNCH(k_alive_by_order, (unsigned)ceil_div(c.nEnt, 256), 256, 0, s, c.nEnt, c.entNR.p, c.entSrc.p,
		          c.iordA.p, c.dT.p, c.aliveByOrd.p);
	}
	int nScat = count_scatterers(c);
	if (cb) cb(user, 0, 0, c.nranks > 1 ? m : nActiveLog, nScat);

	// ---- main flow loop (main.c:408-419)
	const float hx = (float)(0.5 * (double)c.L[0]), hy = (float)(0.5 * (double)c.L[1]),
	            hz = (float)(0.5 * (double)c.L[2]);
	const float fCvg2 = fCvg * fCvg;
	int nIttr = 1;
	// all ranks iterate until NO rank has an active mover (the per-step fScatDens is a global minimum)
	auto global_active = [&](int local) -> long long {
		if (c.nranks <= 1) return local;
		uint32_t *d = c.dCount.alloc(4);
		uint32_t h = (uint32_t)local;
		CK(cudaMemcpyAsync(d + 1, &h, sizeof h, cudaMemcpyHostToDevice, s));
		sk_reduce(c, d + 1, 1, SK_I32, SK_SUM);
		CK(cudaMemcpyAsync(&h, d + 1, sizeof h, cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		return (long long)h;
	};
	long long nGlobal = global_active(c.nActive);
	while (nGlobal) {
		int nl = 0;
		kt.start();
		for (int i = 0; i < 5; ++i) nl += one_step(c, sa, bNoPrune);
		kt.stop(nl);
		// kdPruneInactive
		uint32_t na = 0;
		if (c.nActive > 0) {
			uint32_t *pf = c.flags.alloc(c.nActive);
			uint32_t *ps = c.scan.alloc((size_t)c.nActive + 64);
			SK_LAUNCH(k_prune_flags, (unsigned)ceil_div(c.nActive, 256), 256, 0, s, c.nActive, c.actList.p, c.mx.p,
			          c.my.p, c.mz.p, c.rox.p, c.roy.p, c.roz.p, hx, hy, hz, fCvg2, pf);
			exclusive_scan_u32(pf, ps, c.nActive, c.ws, s);
			SK_LAUNCH(k_prune_compact, (unsigned)ceil_div(c.nActive, 256), 256, 0, s, c.nActive, c.actList.p, pf, ps,
			          c.mx.p, c.my.p, c.mz.p, c.rox.p, c.roy.p, c.roz.p, c.actList2.p);
			CK(cudaMemcpyAsync(&na, ps + c.nActive, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		}
		nScat = count_scatterers(c); // synchronises
		c.nActive = (int)na;
		std::swap(c.actList.p, c.actList2.p);
		std::swap(c.actList.cap, c.actList2.cap);
		nGlobal = global_active(c.nActive);
		if (cb) cb(user, 0, nIttr, (int)nGlobal, nScat);
		++nIttr;
	}
	if (nIttrOut) *nIttrOut = nIttr;
	tm.stop();
}

void stage_microstep(skidgpu_ctx &c, int nSteps, float fStep, skidgpu_log_cb cb, void *user)
{
	cudaStream_t s = c.stream;
	StageTimer tm(c, 3);
	// kdReactivateMove (kd.c:796-799)
	c.nActive = c.shardHi - c.shardLo;
	if (c.nActive > 0)
		SK_LAUNCH(k_iota, (unsigned)ceil_div(c.nActive, 256), 256, 0, s, c.shardLo, c.nActive, c.actList.p);
	StepArgs sa;
	fill_step_args(c, sa, fStep);
	for (int i = 0; i < nSteps; ++i) {
		one_step(c, sa, c.bNoPrune); // smAccDensity keeps cutting scatterers during the micro steps
		if (cb) cb(user, 1, i + 1, c.nActive, count_scatterers(c));
	}
	tm.stop();
}
