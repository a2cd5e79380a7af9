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
#include <cuda_pipeline.h>
#include <utility>

#define T_NONE 0x7f800000u // +inf bits: "no entity was hit this step"
#define DT_NACT_HOST 8     // = DT_NACT (the enum follows StepArgs)


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
                  const float *z, float *mx, float *my, float *mz, float *rox, float *roy, float *roz, int *mOrd)
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
}

__global__ void __launch_bounds__(256) k_iota(int lo, int cnt, uint32_t *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < cnt) out[i] = (uint32_t)(lo + i);
}

// Multi-GPU ownership of movers: block-cyclic over the Morton-ordered mover list (block j of OWN_BLOCK
// movers belongs to rank j % nranks).  Contiguous ranges put whole halos on one rank, and the ranks that
// converged early then wait at every step's agreement point for the one that holds the largest halo
// (measured on 2 GPUs: move 447 ms against 290 ms ideal); interleaved blocks give every rank a
// statistically identical sample.  The blocks are large (4096 Morton-consecutive movers = 128 supertiles):
// with 256 every rank touched every scatterer of the box (the lists of neighbouring blocks overlap almost
// completely), which cost the step kernel a third of its speed at 8 ranks.
constexpr int OWN_BLOCK = MOVE_OWN_BLOCK;
__global__ void __launch_bounds__(256) k_owned_ids(int rank, int nranks, int own, uint32_t *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= own) return;
	const int lb = i / OWN_BLOCK, off = i % OWN_BLOCK;
	out[i] = (uint32_t)((lb * nranks + rank) * OWN_BLOCK + off);
}
static int owned_count(int m, int rank, int nranks)
{
	const int nb = (int)ceil_div(m, OWN_BLOCK);
	int own = 0;
	for (int j = rank; j < nb; j += nranks) own += (j == nb - 1) ? m - j * OWN_BLOCK : OWN_BLOCK;
	return own;
}
__global__ void k_set_u32(uint32_t *p, uint32_t v) { *p = v; }

// Exchange of mover positions (before FoF and before the centres every rank needs every mover):
// the owned blocks are packed into this rank's slot of one buffer (x | y | z planes), all-gathered in place
// and unpacked - 12 bytes per mover on the wire, one collective.
__global__ void __launch_bounds__(256)
    k_pack_owned(int m, int rank, int nranks, int perBlocks, const float *x, const float *y, const float *z, float *buf)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // index within this rank's slot plane
	if (i >= (size_t)perBlocks * OWN_BLOCK) return;
	const size_t lb = i / OWN_BLOCK, off = i % OWN_BLOCK;
	const size_t id = (lb * nranks + rank) * OWN_BLOCK + off;
	const size_t plane = (size_t)perBlocks * OWN_BLOCK;
	float *o = buf + (size_t)rank * 3 * plane;
	const bool ok = id < (size_t)m;
	o[i] = ok ? x[id] : 0.0f;
	o[plane + i] = ok ? y[id] : 0.0f;
	o[2 * plane + i] = ok ? z[id] : 0.0f;
}
__global__ void __launch_bounds__(256)
    k_unpack_all(int m, int rank, int nranks, int perBlocks, const float *buf, float *x, float *y, float *z)
{
	const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= (size_t)m) return;
	const size_t blk = id / OWN_BLOCK, off = id % OWN_BLOCK;
	const int r = (int)(blk % nranks);
	if (r == rank) return;
	const size_t plane = (size_t)perBlocks * OWN_BLOCK;
	const float *o = buf + (size_t)r * 3 * plane + (blk / nranks) * OWN_BLOCK + off;
	x[id] = o[0];
	y[id] = o[plane];
	z[id] = o[2 * plane];
}
static void exchange_positions(skidgpu_ctx &c)
{
	if (c.nranks <= 1 || c.nMove <= 0) return;
	const int m = c.nMove;
	const int nb = (int)ceil_div(m, OWN_BLOCK), perBlocks = (int)ceil_div(nb, c.nranks);
	const size_t plane = (size_t)perBlocks * OWN_BLOCK;
	float *buf = c.mxyz.alloc(3 * plane * c.nranks);
	SK_LAUNCH(k_pack_owned, (unsigned)ceil_div(plane, 256), 256, 0, c.stream, m, c.rank, c.nranks, perBlocks, c.mx.p, c.my.p,
	          c.mz.p, buf);
	sk_allgather(c, buf, (long long)(3 * plane), SK_F32);
	SK_LAUNCH(k_unpack_all, (unsigned)ceil_div(m, 256), 256, 0, c.stream, m, c.rank, c.nranks, perBlocks, buf, c.mx.p, c.my.p,
	          c.mz.p);
}
// all owned movers active, in mover order; the device-side count follows
static void init_active_list(skidgpu_ctx &c)
{
	c.nActiveBound = c.nOwned;
	SK_LAUNCH(k_set_u32, 1, 1, 0, c.stream, c.dT.p + DT_NACT_HOST + c.actPar, (uint32_t)c.nOwned);
	if (c.nOwned <= 0) return;
	if (c.nranks > 1)
		SK_LAUNCH(k_owned_ids, (unsigned)ceil_div(c.nOwned, 256), 256, 0, c.stream, c.rank, c.nranks, c.nOwned, c.actList.p);
	else SK_LAUNCH(k_iota, (unsigned)ceil_div(c.nOwned, 256), 256, 0, c.stream, 0, c.nOwned, c.actList.p);
}

struct StepArgs {
	TreeView tv;
	const float4 *entPos; // (x,y,z,fBall2)
	const float4 *entNR;  // (4/fBall2, fNorm, rhoEff, 0)
	const float *entRho;  // rhoEff alone (the list walks read it beside entPos)
	const float4 *entRec; // the two interleaved: rec[2e] = entPos[e], rec[2e+1] = entNR[e] (list path gathers)
	uint8_t *touched;     // nullable: set for entities with >= 1 hit (step 0, initial cut)
	float *mx, *my, *mz;
	const uint32_t *act;
	int par;  // the current active count lives on the device: dT[8 + par] (the host only knows an upper bound)
	int nEnt;
	uint32_t *dT; // [0] threshold T (float bits), [1] running min of rho over hit entities (float bits); see DT_*
	float fStep;
	float L[3];
	double wrapLo[3], wrapHi[3];
	float *a0x, *a0y, *a0z; // nullable: keep accelerations
	float wrapHiF[3], wrapLoF[3]; // largest floats <= wrapHi / wrapLo: same decisions as the double compares
	// tiles (k_tile_build / k_tile_step): TILE consecutive entries of the position-sorted active list
	uint32_t *tList; // nTiles small slots (TILE_CAP), then nBig big slots (BIG_CAP) from bigBase
	uint32_t *tOff;  // per tile: where its list starts
	uint32_t bigBase, nBig;
	uint32_t *bigCount; // = dT + 6
	float4 *tPos;    // per active slot: position at the last rebuild, reach
	int *tCnt;       // per tile: list length, -1 = overflow (members walk the tree themselves)
	uint32_t *tileQueue;      // tiles that build their list with their own walk
	uint32_t *tileQueueCount; // = dT + 3
	uint32_t *supList; // supCap per supertile (scratch between k_super_walk and k_tile_filter)
	int *supCnt;
	int supCap;
};
// device-side counters of the move loop (uint32 words of dT)
enum { DT_T = 0, DT_MIN = 1, DT_QUEUE = 2, DT_TILEQ = 3, DT_SHORTQ = 5, DT_BIG = 6, DT_TICKET = 7, DT_NACT = 8 /* and 9 */,
       DT_STEPS = 12 /* 64-bit: mover-steps of this stage */, DT_WORDS = 16 };
#define N_ACTIVE(a) ((int)(a).dT[DT_NACT + (a).par])

// MUFU.RSQ of a float that is known to be normal (callers clamp to >= 1e-30): the plain rsqrtf() wraps the same
// instruction in a denormal rescue (compare, two predicated multiplies) that never fires here - 3 of the 27
// instructions of the spline term.  Same bits for normal inputs.
__device__ __forceinline__ float rsqrt_normal(float v)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v));
	return y;
}

// kdMoveParticles (kd.c:711-729) for one mover.  The reference forms ai = fStep/sqrt(|a|^2) in double and
// rounds it to float; here a refined float reciprocal square root gives the same value to <= 2 ulp
// (a position change of ~1e-7 fStep, far below the float spacing of the coordinates), and the wrap
// compares use the largest floats <= centre +- L/2, which decide exactly like the double compares.
__device__ __forceinline__ void move_one(const StepArgs &a, uint32_t id, float x, float y, float z, float ax, float ay,
                                         float az)
{
	const float s2 = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
	float ai;
	if (s2 > 1.0e-30f && s2 < 1.0e30f) {
		float yv = rsqrt_normal(s2);
		yv = yv * fmaf(-0.5f * s2, yv * yv, 1.5f); // one Newton step
		ai = a.fStep * yv;
	} else { // zero, denormal, huge or non-finite: the reference's own arithmetic
		ai = (float)sqrt((double)s2);
		if (ai > 0.0f) ai = (float)((double)a.fStep / sqrt((double)s2));
		else ai = 0.0f;
	}
	float r[3] = {__fsub_rn(x, __fmul_rn(ai, ax)), __fsub_rn(y, __fmul_rn(ai, ay)), __fsub_rn(z, __fmul_rn(ai, az))};
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		if (r[j] > a.wrapHiF[j]) r[j] = __fsub_rn(r[j], a.L[j]);
		if (r[j] <= a.wrapLoF[j]) r[j] = __fadd_rn(r[j], a.L[j]);
	}
	a.mx[id] = r[0];
	a.my[id] = r[1];
	a.mz[id] = r[2];
}

// smAccDensity (smooth1.c:447-459) for one hit: r2 = d2*ih2, rs = sqrt(r2), spline factor, rs *= fNorm,
// a += dx*rs.  The reference evaluates the factor in double from the float rs and rounds it to float;
// here rs and 1/rs come from one MUFU.RSQ refined by a Newton step each (~1 ulp) and the factor is
// formed with FMAs, i.e. to float accuracy without the IEEE sqrt/div sequences (ncu: they were 40 % of
// the instructions of a step).  d2 = 0 (a mover sitting on a scatterer) gives rs = 0, factor -3, as in
// the reference.
#define ACC_HIT(dx, dy, dz, d2, q)                                                                     \
{                                                                                              \
	const float r2_ = __fmul_rn((d2), (q).x);                                              \
	float y_ = rsqrt_normal(fmaxf(r2_, 1.0e-30f));                                               \
	float rs_ = r2_ * y_;                                                                  \
	rs_ = fmaf(0.5f * y_, fmaf(-rs_, rs_, r2_), rs_);                                      \
	y_ = fmaf(y_, fmaf(-rs_, y_, 1.0f), y_);                                               \
	const float gi_ = fmaf(2.25f, rs_, -3.0f);                                             \
	const float go_ = fmaf(-0.75f, rs_, fmaf(-3.0f, y_, 3.0f));                            \
	const float g_ = __fmul_rn(r2_ < 1.0f ? gi_ : go_, (q).y);                             \
	ax = __fadd_rn(ax, __fmul_rn((dx), g_));                                               \
	ay = __fadd_rn(ay, __fmul_rn((dy), g_));                                               \
	az = __fadd_rn(az, __fmul_rn((dz), g_));                                               \
	rmin = fminf(rmin, (q).z);                                                             \
}


// the end of every step for one mover (one thread): min density of the scatterers that hit it, optional
// copy of the acceleration, kdMoveParticles
__device__ __forceinline__ void finish_mover(const StepArgs &a, uint32_t id, float x, float y, float z, float ax,
                                             float ay, float az, float rmin)
{
	// smooth1.c:460-461 (rho > 0).  Millions of atomics on one address serialise in its L2 slice
	// (~2 clocks each): look first, the running minimum settles after a few thousand movers.
	if (rmin < 3.0e38f && __float_as_uint(rmin) < *(volatile uint32_t *)&a.dT[1])
		atomicMin(&a.dT[1], __float_as_uint(rmin));
	if (a.a0x) {
		a.a0x[id] = ax;
		a.a0y[id] = ay;
		a.a0z[id] = az;
	}
	move_one(a, id, x, y, z, ax, ay, az);
}

// warp-wide version: one warp per mover, partial sums in the 32 lanes
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
	if (lane == 0) finish_mover(a, id, x, y, z, ax, ay, az, rmin);
}

// ------------------------------------------------------------------------------------------------
// Default path: tiles.  ncu on the list kernels (profiles/r01_v3_eval_*, r01_v4_*): every unit idle,
// yet 1.1-1.4 ns per mover-step - the per-mover gathers of ~200 scatterer records are 32 different
// 128-byte lines per load instruction, and the L1 tag stage retires about one such line per clock
// and SM (l1tex__t_sectors / cycle = 0.8).  Scattered record fetches, not DRAM bytes and not issue
// slots, are the currency of this stage; sharing them between movers is the only way down.
//
//  * Every 5 steps (kdPruneInactive's rhythm) the ACTIVE movers are sorted by the Morton key of
//    their current position and cut into tiles of TILE consecutive movers: compact by construction,
//    no idle slots as movers freeze.
//  * k_super_walk + k_tile_filter (tile build): per tile, every scatterer whose ball comes within
//    `reach` of a member, reach = the distance a mover can travel before the next rebuild (a step
//    moves a mover by exactly fStep, kd.c:716-721).  Until then the list is complete for every
//    member, so no validity bookkeeping is needed at all.
//  * k_tile_step (one block per tile, one warp per mover, every step): the tile's records are fetched
//    ONCE (TILE x fewer scattered fetches per mover-step) into shared memory, then every warp runs
//    the reference's float32 hit test for its mover against the staged records.
//  * Movers that leave their ball (a periodic wrap moves them by L) and tiles whose list overflows even a big slot
//    (members on both sides of a Morton discontinuity) take the step with their own tree walk
//    (k_move_step on a queue).
// Hit set, hit test and pruning rule are those of the reference (and of the v1 kernel).
constexpr int TILE = 8;  // movers per tile = warps per block
constexpr int TILE_CAP = 512;  // scatterers per tile list (every tile owns a slot of this size)
constexpr int BIG_CAP = 4096;  // ... and the ~4 % of tiles that need more take a big slot from a pool

// Append the lanes with `cand` to the tile's list (warp-wide, order preserving).  A list that outgrows its
// small slot moves to a big slot of the pool; one that outgrows that too (or finds the pool empty) overflows.
#define TILE_APPEND(cand, e_)                                                                          \
	{                                                                                              \
		const uint32_t cm_ = __ballot_sync(SK_FULL, cand);                                     \
		const int nc_ = __popc(cm_);                                                           \
		if (cnt + nc_ > cap) {                                                                 \
			uint32_t slot_ = 0xffffffffu;                                                  \
			if (cap == TILE_CAP) {                                                         \
				if (lane == 0) slot_ = atomicAdd(a.bigCount, 1u);                      \
				slot_ = __shfl_sync(SK_FULL, slot_, 0);                                \
			}                                                                              \
			if (slot_ < a.nBig) {                                                          \
				const uint32_t noff_ = a.bigBase + slot_ * (uint32_t)BIG_CAP;          \
				uint32_t *nl_ = a.tList + noff_;                                       \
				__syncwarp();                                                          \
				for (int i_ = lane; i_ < cnt; i_ += 32) nl_[i_] = list[i_];            \
				list = nl_;                                                            \
				off = noff_;                                                           \
				cap = BIG_CAP;                                                         \
			} else overflow = true;                                                        \
		}                                                                                      \
		if (!overflow) {                                                                       \
			if (cand) list[cnt + __popc(cm_ & lt)] = e_;                                   \
			cnt += nc_;                                                                    \
		}                                                                                      \
	}

// 48-bit Hilbert key (common.cuh) of the current position of every active mover (bbox = box of the initial positions)
__global__ void __launch_bounds__(256) k_mover_keys(int bound, const uint32_t *dN, const uint32_t *act, const float *mx, const float *my,
                                                    const float *mz, const float *bbox, uint64_t *keys, uint32_t *vals)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= bound) return;
	if (i >= (int)*dN) { // beyond the device-side count (the host's bound is stale): sorts to the end
		keys[i] = (1ull << TREE_KEY_BITS) - 1ull;
		vals[i] = 0u;
		return;
	}
	const uint32_t id = act[i];
	const float p[3] = {mx[id], my[id], mz[id]};
	uint32_t q[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		const float ext = bbox[3 + d] - bbox[d];
		float t = ext > 0.0f ? (p[d] - bbox[d]) / ext : 0.0f;
		t = fminf(fmaxf(t, 0.0f), 1.0f);
		q[d] = min((uint32_t)(t * 65536.0f), 65535u);
	}
	keys[i] = hilbert3(q[0], q[1], q[2], 16);
	vals[i] = id;
}

// The list of one tile = the union of its members' candidate sets: every scatterer with
// |x_e - x_m| <= h_e + r for some member m (r = reach + rounding slack).
// candidate <=> d <= h + r for the nearest member <=> u = d2 - h^2 - r^2 <= 2 h r (no square root, 1e-4 slack)
#define TILE_MEMBER_TEST(p, cand)                                                                      \
	{                                                                                              \
		float d2_ = 3.0e38f;                                                                   \
		_Pragma("unroll") for (int m = 0; m < TILE; ++m)                                       \
		{                                                                                      \
			const float dx = (p).x - mxv[m], dy = (p).y - myv[m], dz = (p).z - mzv[m];     \
			d2_ = fminf(d2_, dx * dx + dy * dy + dz * dz);                                 \
		}                                                                                      \
		const float u_ = d2_ - (p).w - r2;                                                     \
		cand = (p).w > 0.0f && (u_ <= 0.0f || u_ * u_ <= 4.0004f * (p).w * r2);                \
	}

// (Measured and dropped: a cheap pre-test of the bucket's scatterers against the members' bounding box before
// the 8-member test - in the dense cores this kernel serves nearly every visited bucket holds a candidate, the
// pre-test only adds to it: list builds 80 -> 84 ms per pass.)
// Slow path of the tile build: the tiles in `queue` walk the tree themselves; nodes and leaf buckets are
// pruned against the members (not only their bounding box), so a tile that straddles a Morton
// discontinuity still gets a short list.
// Short tiles: in the cores of dense clumps the ball radii are smaller than the window's reach and the
// list explodes ((h + 4 fStep)^3 / h^3).  A tile whose list overflows with `reach` is rebuilt with
// `reachShort` (0: valid where the movers are now) and appended to `shortQueue`: it is then walked again
// before EVERY step of the window (one_step launches this kernel on the short queue with reach < 0 =
// "skip the first attempt").
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_tile_walk(const StepArgs a, const uint32_t *queue, const uint32_t *queueCount,
                                                   float reach, float reachShort, uint32_t *shortQueue,
                                                   uint32_t *shortCount)
{
	const int lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1u;
	const uint32_t nq = *queueCount;
	const float T = __uint_as_float(a.dT[0]);
	const int nActive = N_ACTIVE(a);
	// the queued tiles differ a lot in cost (dense cores): a warp draws its next tile from a ticket counter
	// (reset by k_update_T at the end of every step; one launch of this kernel per step)
	uint32_t wi = 0;
	if (lane == 0) wi = atomicAdd(a.dT + DT_TICKET, 1u);
	wi = __shfl_sync(SK_FULL, wi, 0);
	for (; wi < nq;) {
		const int t = (int)queue[wi];
		if (lane == 0) wi = atomicAdd(a.dT + DT_TICKET, 1u);
		wi = __shfl_sync(SK_FULL, wi, 0);
		uint32_t off = (uint32_t)t * TILE_CAP;
		uint32_t *list = a.tList + off;
		int cap = TILE_CAP;
		float mxv[TILE], myv[TILE], mzv[TILE];
		float amax = 0.0f;
#pragma unroll
		for (int m = 0; m < TILE; ++m) {
			const uint32_t id = a.act[min(t * TILE + m, nActive - 1)]; // a short last tile repeats its last member
			mxv[m] = a.mx[id], myv[m] = a.my[id], mzv[m] = a.mz[id];
			amax = fmaxf(amax, fmaxf(fabsf(mxv[m]), fmaxf(fabsf(myv[m]), fabsf(mzv[m]))));
		}
		int cnt = 0;
		bool overflow = true;
		float r = 0.0f;
		for (int attempt = 0; attempt < 2 && overflow; ++attempt) {
			const float rr = attempt == 0 ? reach : reachShort;
			if (rr < 0.0f) continue; // attempt not wanted
			r = rr * 1.001f + 4.0e-7f * amax; // slack for the rounding of the moves and of the tests
			const float r2 = r * r;
			float x0 = 3.0e38f, x1 = -3.0e38f, y0 = 3.0e38f, y1 = -3.0e38f, z0 = 3.0e38f, z1 = -3.0e38f;
#pragma unroll
			for (int m = 0; m < TILE; ++m) {
				x0 = fminf(x0, mxv[m] - r), x1 = fmaxf(x1, mxv[m] + r);
				y0 = fminf(y0, myv[m] - r), y1 = fmaxf(y1, myv[m] + r);
				z0 = fminf(z0, mzv[m] - r), z1 = fmaxf(z1, mzv[m] + r);
			}
			cnt = 0;
			overflow = false;
			int lev = a.tv.top - 1;
			uint32_t node = 0, mymask = 0;
#define TILE_TEST_CHILDREN()                                                                           \
	{                                                                                              \
		const float4 *bx = a.tv.box[lev] + 2 * ((size_t)node * 32 + lane);                     \
		const float4 lo = bx[0], hi = bx[1];                                                   \
		bool in_ = x1 >= lo.x && x0 <= hi.x && y1 >= lo.y && y0 <= hi.y && z1 >= lo.z && z0 <= hi.z && \
		           lo.w >= T;                                                                  \
		if (in_) {                                                                             \
			bool any_ = false;                                                             \
			_Pragma("unroll") for (int m = 0; m < TILE; ++m)                               \
			    any_ |= mxv[m] + r >= lo.x && mxv[m] - r <= hi.x && myv[m] + r >= lo.y && myv[m] - r <= hi.y && \
			            mzv[m] + r >= lo.z && mzv[m] - r <= hi.z;                          \
			in_ = any_;                                                                    \
		}                                                                                      \
		const uint32_t m_ = __ballot_sync(SK_FULL, in_);                                       \
		if (lane == lev) mymask = m_;                                                          \
	}
			TILE_TEST_CHILDREN();
			while (!overflow) {
				uint32_t mk = __shfl_sync(SK_FULL, mymask, lev);
				if (mk == 0) {
					++lev;
					if (lev >= a.tv.top) break;
					node >>= 5;
					continue;
				}
				const int c = __ffs(mk) - 1;
				mk &= mk - 1;
				if (lev > 0) {
					if (lane == lev) mymask = mk;
					--lev;
					node = node * 32 + c;
					TILE_TEST_CHILDREN();
					continue;
				}
				// all the buckets of this node, in order, the records of the next one in flight while this one is
				// tested (a walk is a chain of dependent loads: one warp alone runs at a few per cent of an issue slot;
				// the launch list shows ~400 us per rebuild whether 4 or 200 million instructions are executed - the
				// latency of one tile's walk.  Measured at 2^24: list builds 79.6 -> 73.7 ms, on the massive-halo box
				// builds 53.9 -> 49.2 and the per-step short-tile walks 34.3 -> 29.1 ms; 6 or 5 blocks per SM with
				// 80 / 96 registers instead of 8 with 64: 74.6 / 75.8 ms)
				if (lane == 0) mymask = 0u;
				uint32_t e = (node * 32 + c) * 32 + lane; // arrays are padded with fBall2 = -1 dummies
				float4 p = a.entPos[e];
				float rho = a.entRho[e];
				while (true) {
					const bool more = mk != 0u;
					uint32_t en = 0u;
					float4 pn = p;
					float rhon = 0.0f;
					if (more) {
						const int c1 = __ffs(mk) - 1;
						mk &= mk - 1;
						en = (node * 32 + c1) * 32 + lane;
						pn = a.entPos[en];
						rhon = a.entRho[en];
					}
					bool cand;
					TILE_MEMBER_TEST(p, cand);
					cand = cand && rho >= T; // dead scatterers never come back
					TILE_APPEND(cand, e);
					if (overflow || !more) break;
					e = en, p = pn, rho = rhon;
				}
			}
#undef TILE_TEST_CHILDREN
			if (!overflow && attempt == 1 && shortQueue && lane == 0) shortQueue[atomicAdd(shortCount, 1u)] = (uint32_t)t;
		}
		if (!overflow && lane < TILE && t * TILE + lane < nActive) { // where the list was built and how far it reaches
			float px = mxv[0], py = myv[0], pz = mzv[0];
#pragma unroll
			for (int m = 1; m < TILE; ++m)
				if (lane == m) px = mxv[m], py = myv[m], pz = mzv[m];
			a.tPos[t * TILE + lane] = make_float4(px, py, pz, r);
		}
		if (lane == 0) {
			a.tCnt[t] = overflow ? -1 : cnt;
			a.tOff[t] = off;
		}
	}
}

// Tile lists are built in two kernels.  ncu on the first version (one walk per tile with member tests
// everywhere, profiles/r01_v4_tilebuild_*): 12 k warp instructions per tile, issue bound.
//  * k_super_walk (one warp per supertile = SUPER consecutive tiles = 32 consecutive movers of the
//    sorted active list): ONE walk with cheap bounding-box tests collects the scatterers whose ball
//    comes within r of the supertile's box - a superset of every member tile's list.
//  * k_tile_filter (one warp per tile): filters that superset with the exact member test - 4x fewer
//    tree walks, and the expensive test runs on ~1000 entries instead of ~6000.
// Supertiles whose superset overflows (members far apart) fall back to tile_walk in k_tile_filter.
constexpr int SUPER = 32 / TILE;
constexpr int SUPER_CAP = 2048; // superset capacity per supertile (3072/4096/6144 measured: +-1 %, 6144 slower)
__global__ void __launch_bounds__(128) k_super_walk(const StepArgs a, float reach)
{
	const int lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1u;
	const int st = blockIdx.x * 4 + (threadIdx.x >> 5);
	const int nActive = N_ACTIVE(a);
	if (st * 32 >= nActive) return;
	const float T = __uint_as_float(a.dT[0]);
	const int mi = st * 32 + lane;
	const uint32_t id = a.act[min(mi, nActive - 1)]; // a short last supertile repeats the last mover
	const float x = a.mx[id], y = a.my[id], z = a.mz[id];
	float x0 = x, x1 = x, y0 = y, y1 = y, z0 = z, z1 = z;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		x0 = fminf(x0, __shfl_xor_sync(SK_FULL, x0, o));
		x1 = fmaxf(x1, __shfl_xor_sync(SK_FULL, x1, o));
		y0 = fminf(y0, __shfl_xor_sync(SK_FULL, y0, o));
		y1 = fmaxf(y1, __shfl_xor_sync(SK_FULL, y1, o));
		z0 = fminf(z0, __shfl_xor_sync(SK_FULL, z0, o));
		z1 = fmaxf(z1, __shfl_xor_sync(SK_FULL, z1, o));
	}
	// reach = moves until the last evaluation before the next rebuild x fStep; slack for the rounding of the
	// moves and of the tests
	const float r = reach * 1.001f + 4.0e-7f * fmaxf(fmaxf(fabsf(x0), fabsf(x1)),
	                                                 fmaxf(fmaxf(fabsf(y0), fabsf(y1)), fmaxf(fabsf(z0), fabsf(z1))));
	const float r2 = r * r;
	if (mi < nActive) a.tPos[mi] = make_float4(x, y, z, r);
	uint32_t *sup = a.supList + (size_t)st * a.supCap;
	int ns = 0;
	bool overflow = false;
	int lev = a.tv.top - 1;
	uint32_t node = 0, mymask = 0;
	const float qx0 = x0 - r, qx1 = x1 + r, qy0 = y0 - r, qy1 = y1 + r, qz0 = z0 - r, qz1 = z1 + r;
#define SUPER_TEST_CHILDREN()                                                                          \
	{                                                                                              \
		const float4 *bx = a.tv.box[lev] + 2 * ((size_t)node * 32 + lane);                     \
		const float4 lo = bx[0], hi = bx[1];                                                   \
		const bool in_ = qx1 >= lo.x && qx0 <= hi.x && qy1 >= lo.y && qy0 <= hi.y && qz1 >= lo.z && qz0 <= hi.z && \
		                 lo.w >= T;                                                            \
		const uint32_t m_ = __ballot_sync(SK_FULL, in_);                                       \
		if (lane == lev) mymask = m_;                                                          \
	}
	// superset candidate <=> dist(x_e, box) <= h + r (same algebra as the member test)
#define SUPER_LEAF(e_, p, rho_)                                                                            \
	{                                                                                              \
		const float gx = fmaxf(fmaxf(x0 - (p).x, (p).x - x1), 0.0f), gy = fmaxf(fmaxf(y0 - (p).y, (p).y - y1), 0.0f), \
		            gz = fmaxf(fmaxf(z0 - (p).z, (p).z - z1), 0.0f);                           \
		const float u_ = gx * gx + gy * gy + gz * gz - (p).w - r2;                             \
		bool cand = (p).w > 0.0f && (u_ <= 0.0f || u_ * u_ <= 4.0004f * (p).w * r2);           \
		cand = cand && (rho_) >= T; /* dead scatterers never come back */                      \
		const uint32_t cm = __ballot_sync(SK_FULL, cand);                                      \
		const int nc = __popc(cm);                                                             \
		if (ns + nc > a.supCap) overflow = true;                                               \
		else if (cand) sup[ns + __popc(cm & lt)] = e_;                                         \
		ns += nc;                                                                              \
	}
	SUPER_TEST_CHILDREN();
	while (!overflow) {
		uint32_t mk = __shfl_sync(SK_FULL, mymask, lev);
		if (mk == 0) {
			++lev;
			if (lev >= a.tv.top) break;
			node >>= 5;
			continue;
		}
		const int c = __ffs(mk) - 1;
		mk &= mk - 1;
		if (lev > 0) {
			if (lane == lev) mymask = mk;
			--lev;
			node = node * 32 + c;
			SUPER_TEST_CHILDREN();
			continue;
		}
		// leaf level: two buckets per round so that two record loads are in flight
		const uint32_t e0 = (node * 32 + c) * 32 + lane; // arrays are padded with fBall2 = -1 dummies
		const float4 p0 = a.entPos[e0];
		const float rho0 = a.entRho[e0]; // (a 4-byte coalesced row beside the positions, not a gather behind the test)
		if (mk) {
			const int c1 = __ffs(mk) - 1;
			mk &= mk - 1;
			const uint32_t e1 = (node * 32 + c1) * 32 + lane;
			const float4 p1 = a.entPos[e1];
			const float rho1 = a.entRho[e1];
			if (lane == 0) mymask = mk;
			SUPER_LEAF(e0, p0, rho0);
			SUPER_LEAF(e1, p1, rho1);
		} else {
			if (lane == 0) mymask = mk;
			SUPER_LEAF(e0, p0, rho0);
		}
	}
#undef SUPER_TEST_CHILDREN
#undef SUPER_LEAF
	if (lane == 0) a.supCnt[st] = overflow ? -1 : ns;
}

__global__ void __launch_bounds__(128) k_tile_filter(const StepArgs a)
{
	const int lane = threadIdx.x & 31;
	const uint32_t lt = (1u << lane) - 1u;
	const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
	const int nActive = N_ACTIVE(a);
	if (t * TILE >= nActive) return;
	const float T = __uint_as_float(a.dT[0]);
	const int st = t / SUPER;
	const int ns = a.supCnt[st];
	const uint32_t *sup = a.supList + (size_t)st * a.supCap;
	const float r = a.tPos[t * TILE].w;
	if (ns < 0) { // the supertile's members are far apart: own walk (k_tile_walk)
		if (lane == 0) {
			a.tileQueue[atomicAdd(a.tileQueueCount, 1u)] = (uint32_t)t;
		}
		return;
	}
	uint32_t eN = ns > 0 ? sup[min(lane, ns - 1)] : 0u;
	float mxv[TILE], myv[TILE], mzv[TILE];
#pragma unroll
	for (int m = 0; m < TILE; ++m) {
		const uint32_t id = a.act[min(t * TILE + m, nActive - 1)]; // a short last tile repeats its last member
		mxv[m] = a.mx[id], myv[m] = a.my[id], mzv[m] = a.mz[id];
	}
	const float r2 = r * r;
	uint32_t off = (uint32_t)t * TILE_CAP;
	uint32_t *list = a.tList + off;
	int cap = TILE_CAP;
	int cnt = 0;
	{
		// (A two-pass variant - bounding-box pretest of the tile, survivors compacted in shared memory, member test
		// on the dense survivors - was measured and dropped: the box of 8 members rejects too little of the
		// supertile's superset to pay for the compaction; list builds 117 -> 135 ms per pass.  Nor does fetching the
		// next 32 records while the current ones are tested: 56 -> 64 registers, builds 73.5 -> 79.8 ms.)
		bool overflow = false;
		for (int s0 = 0; s0 < ns && !overflow; s0 += 32) {
			const uint32_t e = eN;
			const float4 p = a.entPos[e];
			if (s0 + 32 < ns) eN = sup[min(s0 + 32 + lane, ns - 1)];
			bool cand;
			TILE_MEMBER_TEST(p, cand);
			cand = cand && s0 + lane < ns;
			TILE_APPEND(cand, e);
		}
		if (overflow) { // too long for the window's reach: k_tile_walk retries and falls back to a zero-reach list
			if (lane == 0) {
				a.tileQueue[atomicAdd(a.tileQueueCount, 1u)] = (uint32_t)t;
			}
			return;
		}
	}
	if (lane == 0) {
		a.tCnt[t] = cnt;
		a.tOff[t] = off;
	}
}

// One mover against the whole scatterer tree (one warp): smBallGather + smAccDensity + the move.  The v1 design
// (profiles/r01_v1_move_*: ~3600 warp instructions per mover-step, 1500 scatterers tested for 85 hits); it now
// serves the few movers a tile cannot (left the reach of their tile's list: a periodic wrap moves them by L, or
// their tile's list overflowed) and the skidgpu_debug_move_kernel test hook.
__device__ __forceinline__ void walk_one_mover(const StepArgs &a, const uint32_t id, const float T, const int lane)
{
	const float x = a.mx[id], y = a.my[id], z = a.mz[id];
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
		const uint32_t e = child * 32 + lane; // arrays are padded with fBall2 = -1 dummies
		const float4 p = a.entPos[e];
		// smBallGather (smooth1.c:365-369): dx = x_scatterer - x_mover, float32, no FMA
		const float dx = __fsub_rn(p.x, x), dy = __fsub_rn(p.y, y), dz = __fsub_rn(p.z, z);
		const float d2 = dist2_rn(dx, dy, dz);
		if (d2 < p.w) {
			const float4 q = a.entNR[e];
			if (q.z >= T) {
				ACC_HIT(dx, dy, dz, d2, q);
				if (a.touched) a.touched[e] = 1;
			}
		}
	}
#undef STEP_TEST_CHILDREN
	finish_step(a, id, x, y, z, ax, ay, az, rmin, lane);
}

// ncu on the first k_tile_step (one warp per mover, profiles/r01_v5_tilestep_*): issue bound at 606
// warp instructions per mover-step, of which only ~310 are hit tests and spline terms - the rest is
// per-warp overhead (prologue, staging loop, 5-stage reductions of 4 values, the move itself).  The
// pair tests per tile are fixed (members x list length), so LPM lanes per mover instead of 32 keep the
// lane work unchanged and divide the per-warp overhead by 32/LPM: one block of TILE*LPM threads per
// tile, 32/LPM movers per warp.
constexpr int TILE_CHUNK = 128; // records staged in shared memory at a time (two buffers)
constexpr int LPM = 8;          // lanes per mover
constexpr int TILE_THREADS = TILE * LPM;
struct TileShared {
	float4 rec[2][TILE_CHUNK][2]; // the scatterer records as they lie in entRec: (x,y,z,fBall2), (4/fBall2, fNorm, rho, 0)
	uint32_t e[2][TILE_CHUNK];    // their indices (only read at step 0, for the touched flags)
};

// ncu on the synchronous version (profiles/r02_tilestep_2e24_*): issue active 68 %, 4.4 warps per issue slot waiting
// on the long scoreboard - the chain list index -> record gather -> shared store of the staging loop.  The records
// now travel with asynchronous copies (cp.async, 16 bytes each, LDGSTS in SASS) into one of two buffers while the
// previous chunk is being evaluated; the list indices of the chunk after that are already in registers.
// (Measured per pass at 2^24: synchronous 168 ms, 2 x 128 records 155 ms, 4 x 64 records 165 ms, 3 x 128 188 ms.)
// Also measured and dropped: two phases per chunk - the 16 hit tests of a lane branch-free into a bit mask, then
// the spline terms of the set bits only (recomputing dx, dy, dz): 149 -> 185 ms.  The hits are not spread evenly
// (a list is ordered along the scatterers' curve: some chunks hit with nearly every record, others with none, and
// the test loop already skips the term warp-wide there), so the mask loop runs max-over-lanes popc times on top
// of the recomputation.
// ncu on this version (profiles/r02b_tilestep_2e24_*): issue active 82 %, the shared-memory pipe at 70 % (LDS.128 of
// records 32 bytes apart: 2-way bank conflicts), 51 % of the warp instructions in the spline term at 12.8 of 32
// lanes.  Measured against it, each within +-3 %: a split [pos | norm] buffer without the conflicts (+2 ms), the
// term without the two Newton refinements and with FMA accumulation (8 of its 27 instructions; -5 ms, not worth
// three more ulp per term), the test loop unrolled by 2 or 4 (+2, +4 ms).  With both units near their limit and
// ~10 warps per scheduler, relieving one of them leaves the other.  Fetching the first chunk's list entries from the
// tile's small slot before count and offset are known (one dependent round trip less per block): +2 ms.
__device__ __forceinline__ void tile_step_body(const StepArgs &a, const int t, TileShared &sh)
{
	const int j = threadIdx.x & (LPM - 1), m = threadIdx.x / LPM;
	const int cnt = a.tCnt[t];
	const int mi = t * TILE + m;
	const bool have = mi < N_ACTIVE(a);
	const uint32_t id = have ? a.act[mi] : 0u;
	float x = 0.0f, y = 0.0f, z = 0.0f;
	if (have) x = a.mx[id], y = a.my[id], z = a.mz[id];
	const float T = __uint_as_float(a.dT[0]);
	bool inside = false;
	if (have && cnt >= 0) { // still within reach of where the list was built? (a periodic wrap moves it by L)
		const float4 b = a.tPos[mi];
		const float ox = x - b.x, oy = y - b.y, oz = z - b.z;
		inside = ox * ox + oy * oy + oz * oz <= b.w * b.w;
	}
	const bool run = have && inside;
	const uint32_t *list = a.tList + a.tOff[t];
	float ax = 0.0f, ay = 0.0f, az = 0.0f;
	float rmin = 3.0e38f;
	constexpr int PER = TILE_CHUNK / TILE_THREADS; // records per thread and chunk
	const int nch = cnt > 0 ? (cnt + TILE_CHUNK - 1) / TILE_CHUNK : 0;
	uint32_t en[PER]; // list entries of the next chunk to be issued
#pragma unroll
	for (int k = 0; k < PER; ++k) en[k] = (int)threadIdx.x + k * TILE_THREADS < cnt ? list[threadIdx.x + k * TILE_THREADS] : 0u;
	auto issue = [&](int ci) { // start the copies of chunk ci (its indices are in en[]), fetch the indices of chunk ci + 1
		const int c0 = ci * TILE_CHUNK, bsel = ci & 1;
#pragma unroll
		for (int k = 0; k < PER; ++k) {
			const int s = (int)threadIdx.x + k * TILE_THREADS;
			if (c0 + s < cnt) {
				const float4 *src = a.entRec + 2 * (size_t)en[k];
				__pipeline_memcpy_async(&sh.rec[bsel][s][0], src, 16);
				__pipeline_memcpy_async(&sh.rec[bsel][s][1], src + 1, 16);
				if (a.touched) sh.e[bsel][s] = en[k];
			}
			const int nx = c0 + TILE_CHUNK + s;
			en[k] = nx < cnt ? list[nx] : 0u;
		}
		__pipeline_commit();
	};
	if (nch > 0) issue(0);
	for (int ci = 0; ci < nch; ++ci) {
		const int bsel = ci & 1;
		const int nc = min(cnt - ci * TILE_CHUNK, TILE_CHUNK);
		if (ci + 1 < nch) {
			issue(ci + 1);
			__pipeline_wait_prior(1);
		} else __pipeline_wait_prior(0);
		__syncthreads(); // every thread's copies of chunk ci have landed
		if (run) {
			for (int s = j; s < nc; s += LPM) {
				const float4 p = sh.rec[bsel][s][0];
				// smBallGather (smooth1.c:365-369): dx = x_scatterer - x_mover, float32, no FMA
				const float dx = __fsub_rn(p.x, x), dy = __fsub_rn(p.y, y), dz = __fsub_rn(p.z, z);
				const float d2 = dist2_rn(dx, dy, dz);
				if (d2 < p.w) {
					const float4 q = sh.rec[bsel][s][1];
					if (q.z >= T) { // not pruned since the list was built
						ACC_HIT(dx, dy, dz, d2, q);
						if (a.touched) a.touched[sh.e[bsel][s]] = 1;
					}
				}
			}
		}
		__syncthreads(); // the buffer is free for chunk ci + 2
	}
#pragma unroll
	for (int o = LPM / 2; o > 0; o >>= 1) {
		ax += __shfl_xor_sync(SK_FULL, ax, o);
		ay += __shfl_xor_sync(SK_FULL, ay, o);
		az += __shfl_xor_sync(SK_FULL, az, o);
		rmin = fminf(rmin, __shfl_xor_sync(SK_FULL, rmin, o));
	}
	if (run && j == 0) finish_mover(a, id, x, y, z, ax, ay, az, rmin);
	// members the list cannot serve (overflowed tile, or left its reach) take the step with their own tree walk,
	// one after the other on this warp: rare, and hidden behind the other tiles of the launch
	uint32_t own = __ballot_sync(SK_FULL, have && !inside && j == 0);
	while (own) {
		const int src = __ffs(own) - 1;
		own &= own - 1;
		walk_one_mover(a, __shfl_sync(SK_FULL, id, src), T, threadIdx.x & 31);
	}
}

// After a step: adopt the new threshold (ScatterCut, smooth1.c:509-513).  If nothing was hit the
// reference's fScatDens stays 0.0 and nothing is cut.  Also resets the per-step counters and adds the step's
// active movers to the stage's mover-step counter.
__device__ __forceinline__ void update_T(uint32_t *dT, int bNoPrune, int par, int launched)
{
	uint32_t nx = dT[DT_MIN];
	if (!bNoPrune && nx != T_NONE) dT[DT_T] = nx;
	dT[DT_MIN] = T_NONE;
	dT[DT_QUEUE] = 0u;
	dT[DT_TILEQ] = 0u;
	dT[4] = 0u;
	dT[DT_TICKET] = 0u; // ticket counter of k_tile_walk
	if (launched) *(unsigned long long *)(dT + DT_STEPS) += dT[DT_NACT + par];
}
__global__ void k_update_T(uint32_t *dT, int bNoPrune, int par, int launched) { update_T(dT, bNoPrune, par, launched); }

// One block per tile; the grid is sized from the host's upper bound of the active count.  (Letting the last
// block to finish adopt the step's threshold instead of launching k_update_T was measured and dropped: the
// completion count - even two-level, 256 blocks per word - and its fences cost every one of the ~10^6 small
// blocks more than the launch it saves: k_tile_step 164 -> 175 ms per pass.)
__global__ void __launch_bounds__(TILE_THREADS) k_tile_step(const StepArgs a)
{
	__shared__ __align__(16) TileShared sh;
	const int t = blockIdx.x;
	if (t * TILE < N_ACTIVE(a)) tile_step_body(a, t, sh);
}

// Test hook (skidgpu_debug_move_kernel): every active mover takes the step with its own tree walk.
constexpr int STEP_WARPS = 8;
__global__ void __launch_bounds__(STEP_WARPS * 32) k_move_step(const StepArgs a, const uint32_t *ids, const uint32_t *countPtr)
{
	const int lane = threadIdx.x & 31;
	const int n = (int)*countPtr;
	const float T = __uint_as_float(a.dT[0]);
	for (int wi = blockIdx.x * STEP_WARPS + (threadIdx.x >> 5); wi < n; wi += gridDim.x * STEP_WARPS)
		walk_one_mover(a, ids[wi], T, lane);
}

// Initial cut (smooth1.c:463-470,500-507): entities that scattered onto nobody get fDensity = 0.
__global__ void __launch_bounds__(256) k_initial_cut(int nEnt, const uint8_t *touched, float4 *entNR, float4 *entRec, float *entRho)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < nEnt && !touched[e]) {
		entNR[e].z = 0.0f;
		entRec[2 * (size_t)e + 1].z = 0.0f;
		entRho[e] = 0.0f;
	}
}

// ... and their fDensity reads as 0 afterwards, which kdOutStats' "gas mass" test sees (kd.c:1792-1794): a copy of
// the densities by iOrder with the cut originals zeroed (only made when the input has gas, i.e. with -fic)
__global__ void __launch_bounds__(256)
    k_cut_density(int nEnt, const uint8_t *touched, const uint32_t *entSrc, const int *iordA, float *rhoStat)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nEnt || touched[e]) return;
	const uint32_t src = entSrc[e];
	if (!(src & 0x80000000u)) rhoStat[iordA[src]] = 0.0f;
}

// nScatter of the log line = surviving originals + surviving replicas (smooth1.c:517).  Grid-stride over
// [lo, hi) with one atomic per block (one per warp serialised 570 k same-address atomics per call).
__global__ void __launch_bounds__(256) k_count_scatter(int lo, int hi, const float *entRho, const uint32_t *dT,
                                                       uint32_t *out)
{
	__shared__ uint32_t s_cnt;
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads();
	const float T = __uint_as_float(dT[0]);
	uint32_t mine = 0;
	for (int e = lo + blockIdx.x * blockDim.x + threadIdx.x; e < hi; e += gridDim.x * blockDim.x)
		mine += entRho[e] >= T ? 1u : 0u; // (the compact copy of rhoEff: 4 bytes per entity instead of a 16-byte stride)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(SK_FULL, mine, o);
	if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
	__syncthreads();
	if (threadIdx.x == 0 && s_cnt) atomicAdd(out, s_cnt);
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
// last check.  flags -> scan -> stable compaction of the active list.  `bound` is the host's upper bound of
// the device-side count dN[0]; entries beyond the count get flag 0.
__global__ void __launch_bounds__(256)
    k_prune_flags(int bound, const uint32_t *dN, const uint32_t *act, const float *mx, const float *my, const float *mz,
                  const float *rox, const float *roy, const float *roz, float hx, float hy, float hz, float fCvg2,
                  uint32_t *flags)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= bound) return;
	if (i >= (int)*dN) {
		flags[i] = 0u;
		return;
	}
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

// ... and the new count goes to the other device-side slot and to this block's log slot:
// log[0] = active movers (summed over the ranks afterwards), log[2] = this rank's own count
__global__ void __launch_bounds__(256)
    k_prune_compact(int bound, const uint32_t *act, const uint32_t *flags, const uint32_t *scan,
                    const float *mx, const float *my, const float *mz, float *rox, float *roy, float *roz,
                    uint32_t *actOut, uint32_t *dNnext, uint32_t *log)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) {
		const uint32_t tot = bound > 0 ? scan[bound] : 0u;
		*dNnext = tot;
		log[0] = tot;
		log[2] = tot;
	}
	if (i >= bound || !flags[i]) return;
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
	sa.entRho = c.eRhoSorted.p;
	sa.entRec = c.entRec.p;
	sa.touched = nullptr;
	sa.mx = c.mx.p;
	sa.my = c.my.p;
	sa.mz = c.mz.p;
	sa.act = c.actList.p;
	sa.par = c.actPar;
	sa.nEnt = c.nEnt;
	sa.dT = c.dT.p;
	sa.fStep = fStep;
	for (int d = 0; d < 3; ++d) {
		sa.L[d] = c.L[d];
		sa.wrapHi[d] = (double)c.C[d] + 0.5 * (double)c.L[d]; // kd.c:724
		sa.wrapLo[d] = (double)c.C[d] - 0.5 * (double)c.L[d]; // kd.c:726
	}
	sa.a0x = sa.a0y = sa.a0z = nullptr;
	for (int d = 0; d < 3; ++d) { // r > t  <=>  r > (largest float <= t) for float r (same for <=)
		float h = (float)sa.wrapHi[d], l = (float)sa.wrapLo[d];
		if ((double)h > sa.wrapHi[d]) h = nextafterf(h, -INFINITY);
		if ((double)l > sa.wrapLo[d]) l = nextafterf(l, -INFINITY);
		sa.wrapHiF[d] = h;
		sa.wrapLoF[d] = l;
	}
	sa.tList = c.tList.p;
	sa.tOff = c.tOff.p;
	sa.bigBase = c.bigBase;
	sa.nBig = c.nBig;
	sa.bigCount = c.dT.p ? c.dT.p + DT_BIG : nullptr;
	sa.tPos = c.tPos.p;
	sa.tCnt = c.tCnt.p;
	sa.supList = c.supList.p;
	sa.tileQueue = c.tileQueue.p;
	sa.tileQueueCount = c.dT.p ? c.dT.p + DT_TILEQ : nullptr;
	sa.supCnt = c.supCnt.p;
	sa.supCap = c.superCap;
}

// Log slots: one per Ittr block / micro step, 4 words: [0] active movers (all ranks), [1] surviving scatterers
// (all ranks), [2] this rank's active movers.  Written on the device, copied to pinned host memory behind the
// block that produced them and read by the host a few blocks later: the loop never waits for the GPU.
constexpr int LOG_SLOTS = 1024;
constexpr int LOG_LAG = 2; // blocks enqueued ahead of the last count the host has seen

static uint32_t *log_slot(skidgpu_ctx &c, int b) { return c.dLog.p + 4 * (size_t)(b % LOG_SLOTS); }

static void log_prepare(skidgpu_ctx &c)
{
	c.dLog.alloc(4 * LOG_SLOTS);
	if (!c.hLog) CK(cudaHostAlloc((void **)&c.hLog, sizeof(uint32_t) * 4 * LOG_SLOTS, cudaHostAllocDefault));
}
static cudaEvent_t log_event(skidgpu_ctx &c, int b)
{
	while ((int)c.logEv.size() <= b % LOG_SLOTS) {
		cudaEvent_t e;
		CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		c.logEv.push_back(e);
	}
	return c.logEv[b % LOG_SLOTS];
}

// nScatter of a log line into slot[1]; every rank counts its slice of the (replicated) scatterers
static void enqueue_count_scatterers(skidgpu_ctx &c, uint32_t *slot)
{
	const int lo = (int)((long long)c.nEnt * c.rank / c.nranks), hi = (int)((long long)c.nEnt * (c.rank + 1) / c.nranks);
	if (hi > lo) {
		unsigned g = (unsigned)ceil_div(hi - lo, 256 * 8);
		SK_LAUNCH(k_count_scatter, g > 148u * 8u ? 148u * 8u : g, 256, 0, c.stream, lo, hi, c.eRhoSorted.p, c.dT.p, slot + 1);
	}
}
// close a log slot: sum over the ranks, copy to the pinned mirror, mark with an event
static void enqueue_log_fetch(skidgpu_ctx &c, int b)
{
	uint32_t *slot = log_slot(c, b);
	sk_reduce(c, slot, 2, SK_I32, SK_SUM);
	CK(cudaMemcpyAsync(c.hLog + 4 * (size_t)(b % LOG_SLOTS), slot, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
	CK(cudaEventRecord(log_event(c, b), c.stream));
}
static const uint32_t *wait_log(skidgpu_ctx &c, int b)
{
	CK(cudaEventSynchronize(log_event(c, b)));
	return c.hLog + 4 * (size_t)(b % LOG_SLOTS);
}

// (Measured and dropped, twice: running this kernel on a second stream - beside k_tile_step in round 1, beside
// k_tile_filter for the tiles of overflowed supertiles in round 2 (queued by k_super_walk, so that the walks start
// before the filter): list builds 73.5 -> 75.1 ms at 2^24, 49.2 -> 52.0 ms on the massive-halo box.  The kernels it
// would hide behind are issue bound; what it gains in latency they lose in issue slots.)
// k_tile_walk is latency bound (dependent tree loads): 8 resident blocks per SM (64 registers via the launch
// bounds), persistent grid of that many blocks per SM, tiles drawn from a ticket counter.
static void launch_tile_walk(skidgpu_ctx &c, const StepArgs &sa, const uint32_t *queue, const uint32_t *queueCount,
                             float reach, float reachShort, uint32_t *shortQueue, uint32_t *shortCount)
{
	// (at most one warp per tile of the host's bound: small runs do not pay for 1184 idle blocks)
	const unsigned want = (unsigned)ceil_div(ceil_div(c.nActiveBound > 0 ? c.nActiveBound : 1, TILE), 4);
	SK_LAUNCH(k_tile_walk<8>, want < 148u * 8u ? want : 148u * 8u, 128, 0, c.stream, sa, queue, queueCount, reach, reachShort,
	          shortQueue, shortCount);
}

// Sort the active movers by position and build the tile lists; valid for `steps` steps of length fStep.
static void rebuild_tiles(skidgpu_ctx &c, StepArgs &sa, int steps)
{
	cudaStream_t s = c.stream;
	c.tileStepsLeft = steps;
	const int bound = c.nActiveBound;
	if (bound <= 0 || c.nEnt <= 0) return;
	c.spans.begin(KF_BUILD, s);
	// The active list starts in Morton order of the initial positions and compaction keeps its order, so
	// tiles stay compact for a while: re-sort by current position only every 4th rebuild (the lists are
	// unions of per-member neighbourhoods - compactness is efficiency, never correctness).
	if (c.tileBuilds % 4 == 3) {
		uint64_t *keys = c.tKeys.alloc(bound);
		SK_LAUNCH(k_mover_keys, (unsigned)ceil_div(bound, 256), 256, 0, s, bound, c.dT.p + DT_NACT + c.actPar, c.actList.p, c.mx.p,
		          c.my.p, c.mz.p, c.treeM.bbox.p, keys, c.actList2.p);
		radix_sort_pairs(keys, c.actList2.p, bound, TREE_KEY_BITS, c.ws, s); // entries beyond the device-side count sort to the end
		std::swap(c.actList.p, c.actList2.p);
		std::swap(c.actList.cap, c.actList2.cap);
	}
	++c.tileBuilds;
	sa.act = c.actList.p;
	sa.par = c.actPar;
	const int nTiles = (int)ceil_div(bound, TILE), nSuper = (int)ceil_div(nTiles, SUPER);
	CK(cudaMemsetAsync(c.dT.p + DT_SHORTQ, 0, 2 * sizeof(uint32_t), s)); // short-tile queue, big slots in use
	// a list built now is evaluated at the current positions and after 1 .. steps-1 moves of length fStep
	const float reach = (float)(steps - 1) * sa.fStep;
	SK_LAUNCH(k_super_walk, (unsigned)ceil_div(nSuper, 4), 128, 0, s, sa, reach);
	SK_LAUNCH(k_tile_filter, (unsigned)ceil_div(nTiles, 4), 128, 0, s, sa);
	// spread-out supertiles build per tile; tiles whose 5-step list overflows become short tiles (DT_SHORTQ)
	launch_tile_walk(c, sa, sa.tileQueue, sa.tileQueueCount, reach, steps > 1 ? 0.0f : -1.0f, c.shortQueue.p, c.dT.p + DT_SHORTQ);
	c.spans.end(s);
	c.tileFresh = true;
}

// One step of smAccDensity + kdMoveParticles for every active mover (everything is enqueued; nothing waits).
static void one_step(skidgpu_ctx &c, StepArgs &sa, int bNoPrune)
{
	cudaStream_t s = c.stream;
	const int bound = c.nActiveBound;
	int launched = 0;
	if (bound > 0 && c.nEnt > 0) {
		launched = 1;
		sa.act = c.actList.p;
		sa.par = c.actPar;
		if (c.moveKernel == 0) {
			if (c.tileStepsLeft <= 0) rebuild_tiles(c, sa, c.tileWindow);
			--c.tileStepsLeft;
			if (!c.tileFresh) { // short tiles are rebuilt before every step (their list has no reach)
				c.spans.begin(KF_FALLBACK, s);
				launch_tile_walk(c, sa, c.shortQueue.p, c.dT.p + DT_SHORTQ, -1.0f, 0.0f, nullptr, nullptr);
				c.spans.end(s);
			}
			c.tileFresh = false;
			c.spans.begin(KF_TILE_STEP, s);
			SK_LAUNCH(k_tile_step, (unsigned)ceil_div(bound, TILE), TILE_THREADS, 0, s, sa);
			c.spans.end(s);
		} else { // test hook: a tree walk per mover and step (the v1 kernel)
			unsigned g = (unsigned)ceil_div(bound, STEP_WARPS);
			SK_LAUNCH(k_move_step, g > 148u * 64u ? 148u * 64u : g, STEP_WARPS * 32, 0, s, sa, sa.act,
			          (const uint32_t *)(c.dT.p + DT_NACT + c.actPar));
		}
	}
	if (!bNoPrune) sk_reduce(c, c.dT.p + DT_MIN, 1, SK_I32, SK_MIN); // fScatDens over all ranks' movers (+inf bits = none)
	SK_LAUNCH(k_update_T, 1, 1, 0, s, c.dT.p, bNoPrune, c.actPar, launched);
}

// kdPruneInactive + the counts of one "Ittr" line into log slot b
static void enqueue_prune(skidgpu_ctx &c, int b, float hx, float hy, float hz, float fCvg2)
{
	cudaStream_t s = c.stream;
	const int bound = c.nActiveBound;
	uint32_t *slot = log_slot(c, b);
	c.spans.begin(KF_PRUNE, s);
	CK(cudaMemsetAsync(slot, 0, 4 * sizeof(uint32_t), s));
	uint32_t *pf = c.flags.alloc((size_t)bound + 1);
	uint32_t *ps = c.scan.alloc((size_t)bound + 64);
	uint32_t *dN = c.dT.p + DT_NACT;
	if (bound > 0) {
		SK_LAUNCH(k_prune_flags, (unsigned)ceil_div(bound, 256), 256, 0, s, bound, dN + c.actPar, c.actList.p, c.mx.p, c.my.p,
		          c.mz.p, c.rox.p, c.roy.p, c.roz.p, hx, hy, hz, fCvg2, pf);
		exclusive_scan_u32(pf, ps, bound, c.ws, s);
	}
	SK_LAUNCH(k_prune_compact, (unsigned)ceil_div(bound > 0 ? bound : 1, 256), 256, 0, s, bound, c.actList.p, pf, ps, c.mx.p,
	          c.my.p, c.mz.p, c.rox.p, c.roy.p, c.roz.p, c.actList2.p, dN + (c.actPar ^ 1), slot);
	enqueue_count_scatterers(c, slot);
	c.spans.end(s);
	enqueue_log_fetch(c, b);
	c.actPar ^= 1;
	std::swap(c.actList.p, c.actList2.p);
	std::swap(c.actList.cap, c.actList2.cap);
	c.tileStepsLeft = 0; // the active list changed: new tiles
}

static void resolve_spans(skidgpu_ctx &c) { c.spans.resolve(c.kernel_ms, c.kernel_launches); }

void stage_move(skidgpu_ctx &c, float fDensMin, float fTempMax, float fMassMax, float fCvg, float fStep,
                int bForceInitialCut, int bNoPrune, skidgpu_log_cb cb, void *user, int *nMoveOut, int *nIttrOut)
{
	cudaStream_t s = c.stream;
	const int n = c.n;
	if (n <= 0) throw SkidError("skidgpu_move: no particles set");
	if (!c.rho.p) throw SkidError("skidgpu_move: skidgpu_density has not run");
	StageTimer tm(c, 1);
	c.bNoPrune = bNoPrune;
	c.haveRhoStat = false;
	for (int f : {KF_TILE_STEP, KF_BUILD, KF_FALLBACK, KF_PRUNE}) {
		c.kernel_ms[f] = 0;
		c.kernel_launches[f] = 0;
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
	uint32_t *dT = c.dT.alloc(DT_WORDS);
	uint32_t initT[DT_WORDS] = {0u, T_NONE}; // threshold, running min of this step, queue lengths, ticket, counts
	CK(cudaMemcpyAsync(dT, initT, sizeof initT, cudaMemcpyHostToDevice, s));
	log_prepare(c);
	c.actPar = 0;
	c.shardLo = (int)((long long)m * c.rank / c.nranks);
	c.shardHi = (int)((long long)m * (c.rank + 1) / c.nranks);
	c.nOwned = c.nranks > 1 ? owned_count(m, c.rank, c.nranks) : m; // block-cyclic ownership on several GPUs
	if (m > 0) {
		uint32_t *fileIdx = c.actList2.alloc(m);
		SK_LAUNCH(k_compact_idx2, (unsigned)ceil_div(n, 256), 256, 0, s, n, flags, scan, fileIdx);
		float *gx = c.tmpx.alloc(m), *gy = c.tmpy.alloc(m), *gz = c.tmpz.alloc(m);
		SK_LAUNCH(k_gather3b, (unsigned)ceil_div(m, 256), 256, 0, s, m, fileIdx, c.x.p, c.y.p, c.z.p, gx, gy, gz);
		tree_sort_points(c.treeM, gx, gy, gz, m, c.ws, s, &c);
		c.mx.alloc(m);
		c.my.alloc(m);
		c.mz.alloc(m);
		c.rox.alloc(m);
		c.roy.alloc(m);
		c.roz.alloc(m);
		c.mOrd.alloc(m);
		const int own = c.nOwned;
		{
			const size_t nt = ceil_div(own > 0 ? own : 1, TILE);
			// list offsets are 32-bit: 8.3 M tiles (67 M movers on ONE GPU) would overflow them - fail loudly
			if (nt * TILE_CAP + (nt / 8 + 1024) * (size_t)BIG_CAP >= (1ull << 32))
				throw SkidError("skidgpu_move: too many movers on one GPU for the 32-bit tile-list offsets; shard the "
				                "movers over more GPUs (skidgpu_comm_init)");
			c.bigBase = (uint32_t)(nt * TILE_CAP);
			c.nBig = (uint32_t)(nt / 8 + 1024);
			c.tList.alloc(nt * TILE_CAP + (size_t)c.nBig * BIG_CAP);
			c.tOff.alloc(nt);
			c.tPos.alloc(nt * TILE);
			c.tCnt.alloc(nt);
			c.superCap = SUPER_CAP;
			c.supList.alloc(ceil_div(nt, SUPER) * (size_t)c.superCap);
			c.supCnt.alloc(ceil_div(nt, SUPER));
			c.tileQueue.alloc(nt);
			c.shortQueue.alloc(nt);
		}
		c.tileStepsLeft = 0;
		c.tileBuilds = 0;
		SK_LAUNCH(k_init_movers, (unsigned)ceil_div(m, 256), 256, 0, s, m, c.treeM.perm.p, fileIdx, c.x.p, c.y.p,
		          c.z.p, c.mx.p, c.my.p, c.mz.p, c.rox.p, c.roy.p, c.roz.p, c.mOrd.p);
		c.actList.alloc(m);
		c.actList2.alloc(m); // fileIdx no longer needed after k_init_movers (same stream)
	}
	init_active_list(c);
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
	c.tileWindow = 1; // step 0 is followed by the initial cut and the log line; the blocks of 5 start after it
	one_step(c, sa, bNoPrune);
	c.tileStepsLeft = 0;
	c.tileWindow = 5;
	if (bInitial && c.nEnt > 0) {
		sk_reduce(c, c.entTouched.p, c.nEnt, SK_U8, SK_MAX);
		SK_LAUNCH(k_initial_cut, (unsigned)ceil_div(c.nEnt, 256), 256, 0, s, c.nEnt, c.entTouched.p, c.entNR.p, c.entRec.p,
		          c.eRhoSorted.p);
		if (c.nGas > 0) {
			CK(cudaMemcpyAsync(c.rhoStat.alloc(n), c.rho.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, s));
			SK_LAUNCH(k_cut_density, (unsigned)ceil_div(c.nEnt, 256), 256, 0, s, c.nEnt, c.entTouched.p, c.entSrc.p, c.iordA.p,
			          c.rhoStat.p);
			c.haveRhoStat = true;
		}
	}
	sa.touched = nullptr;
	sa.a0x = sa.a0y = sa.a0z = nullptr;
	if (c.keepStep0 && c.nEnt > 0) {
		CK(cudaMemsetAsync(c.aliveByOrd.alloc(n), 0, n, s));
		SK_LAUNCH(k_alive_by_order, (unsigned)ceil_div(c.nEnt, 256), 256, 0, s, c.nEnt, c.entNR.p, c.entSrc.p,
		          c.iordA.p, c.dT.p, c.aliveByOrd.p);
	}
	{ // "Ittr:0" line: all movers active
		uint32_t *slot = log_slot(c, 0);
		CK(cudaMemsetAsync(slot, 0, 4 * sizeof(uint32_t), s));
		enqueue_count_scatterers(c, slot);
		enqueue_log_fetch(c, 0);
	}

	// ---- main flow loop (main.c:408-419).  The host runs LOG_LAG blocks ahead of the counts it has seen: the
	// active count lives on the device, grids are sized from the last count the host knows (counts only fall),
	// and the blocks enqueued after the last mover froze find nothing to do.  All ranks see the same summed
	// count, hence take the same decisions.
	const float hx = (float)(0.5 * (double)c.L[0]), hy = (float)(0.5 * (double)c.L[1]),
	            hz = (float)(0.5 * (double)c.L[2]);
	const float fCvg2 = fCvg * fCvg;
	int nIttr = 1;      // log lines delivered
	int enq = 1;        // log slots enqueued (slot 0 = step 0)
	bool done = m == 0; // main.c:408: while (nActive)
	{
		const uint32_t *l0 = wait_log(c, 0);
		if (cb) cb(user, 0, 0, m, (int)l0[1]);
	}
	while (!done) {
		for (int i = 0; i < 5; ++i) one_step(c, sa, bNoPrune);
		enqueue_prune(c, enq, hx, hy, hz, fCvg2);
		++enq;
		// read the counts that are LOG_LAG blocks old (all ranks must enqueue the same blocks, so only a single
		// GPU may also take counts that happen to be ready earlier)
		while (nIttr < enq &&
		       (enq - nIttr > LOG_LAG || (c.nranks == 1 && cudaEventQuery(log_event(c, nIttr)) == cudaSuccess))) {
			const uint32_t *l = wait_log(c, nIttr);
			c.nActiveBound = (int)l[2];
			if (cb) cb(user, 0, nIttr, (int)l[0], (int)l[1]);
			++nIttr;
			if (l[0] == 0u) {
				done = true;
				break;
			}
		}
	}
	// blocks enqueued past the end moved nothing; wait for them and for the counters
	unsigned long long steps = 0;
	CK(cudaMemcpyAsync(&steps, c.dT.p + DT_STEPS, sizeof steps, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	c.moverSteps += (long long)steps;
	c.nActive = 0;
	c.nActiveBound = 0;
	exchange_positions(c);
	if (nIttrOut) *nIttrOut = nIttr;
	tm.stop();
	resolve_spans(c);
}

void stage_microstep(skidgpu_ctx &c, int nSteps, float fStep, skidgpu_log_cb cb, void *user)
{
	cudaStream_t s = c.stream;
	StageTimer tm(c, 3);
	if (!c.dT.p) throw SkidError("skidgpu_microstep: skidgpu_move has not run");
	if (nSteps > LOG_SLOTS) throw SkidError("skidgpu_microstep: too many steps");
	// kdReactivateMove (kd.c:796-799)
	init_active_list(c);
	CK(cudaMemsetAsync(c.dT.p + DT_STEPS, 0, 2 * sizeof(uint32_t), s));
	c.tileStepsLeft = 0;
	c.tileWindow = nSteps > 0 ? nSteps : 1;
	StepArgs sa;
	fill_step_args(c, sa, fStep);
	for (int i = 0; i < nSteps; ++i) {
		one_step(c, sa, c.bNoPrune); // smAccDensity keeps cutting scatterers during the micro steps
		uint32_t *slot = log_slot(c, i);
		CK(cudaMemsetAsync(slot, 0, 4 * sizeof(uint32_t), s));
		enqueue_count_scatterers(c, slot);
		enqueue_log_fetch(c, i);
	}
	for (int i = 0; i < nSteps; ++i) {
		const uint32_t *l = wait_log(c, i);
		if (cb) cb(user, 1, i + 1, c.nOwned, (int)l[1]);
	}
	unsigned long long steps = 0;
	CK(cudaMemcpyAsync(&steps, c.dT.p + DT_STEPS, sizeof steps, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	c.moverSteps += (long long)steps;
	exchange_positions(c);
	tm.stop();
	resolve_spans(c);
}
