// Stage 4/5 glue and stage 5: group catalogue, density centres, unbinding, too-small removal.
//
// Replaces kdInitpGroup (kd.c:969-1005), kdCalcCenter (kd.c:1013-1075), the centre-of-mass
// fallback of kdReadCenter (kd.c:1160-1193), kdGroupOrder (kd.c:1205-1245), kdUnbind
// (kd.c:1299-1466), kdCellPot/kdSubPot/kdAddScoopPot (grav.c:8-135, SPLINE_POT grav.h:11-31),
// kdTooSmall (kd.c:1251-1293) and the radius loop of kdWriteGroup (kd.c:1627-1644).
//
// Pair terms keep the reference's arithmetic: float32 geometry (no FMA), float64 softened
// 1/r rounded to float32 ("dir"), float32 products G*m*dir, float64 accumulation.  No tensor
// cores: this is an O(n^2) scalar-potential sum, not a dense contraction.
#include "ctx.cuh"
#include <cooperative_groups.h>

// per-group f64 accumulators
#define GA_MASS 0
#define GA_MVX 1
#define GA_MVY 2
#define GA_MVZ 3
#define GA_CX 4
#define GA_CY 5
#define GA_CZ 6
#define GA_STRIDE 8

__global__ void __launch_bounds__(256) k_rep_min(int n, const int *gid, int *repOrd)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int g = gid[i];
	if (g > 0) atomicMin(&repOrd[g], i);
}

__global__ void __launch_bounds__(256)
    k_group_acc(int lo, int n, const int *gid, const float *mass, const float *vx, const float *vy, const float *vz,
                int *gN, double *acc)
{
	int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int g = gid[i];
	if (g <= 0) return;
	atomicAdd(&gN[g], 1);
	float m = mass[i];
	double *a = acc + (size_t)g * GA_STRIDE;
	atomicAdd(&a[GA_MASS], (double)m);
	atomicAdd(&a[GA_MVX], (double)__fmul_rn(m, vx[i]));
	atomicAdd(&a[GA_MVY], (double)__fmul_rn(m, vy[i]));
	atomicAdd(&a[GA_MVZ], (double)__fmul_rn(m, vz[i]));
}

__device__ __forceinline__ float wrap_del(float del, float L)
{ // kd.c:1038-1039: compares against 0.5*fPeriod in double, subtracts the float period
	if ((double)del > 0.5 * (double)L) del = __fsub_rn(del, L);
	if ((double)del <= -0.5 * (double)L) del = __fadd_rn(del, L);
	return del;
}

// kdCalcCenter (kd.c:1032-1042): sum of min-image offsets of the MOVED positions from rel
__global__ void __launch_bounds__(256)
    k_center_acc_movers(int m, const int *mOrd, const int *gid, const int *repOrd, const float *mx, const float *my,
                        const float *mz, const float *x, const float *y, const float *z, float Lx, float Ly,
                        float Lz, double *acc, int ownBlock, int rank, int nranks)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m || (i / ownBlock) % nranks != rank) return; // every rank adds the movers it owns (move.cu: block-cyclic)
	int g = gid[mOrd[i]];
	if (g <= 0) return;
	int r = repOrd[g];
	double *a = acc + (size_t)g * GA_STRIDE;
	atomicAdd(&a[GA_CX], (double)wrap_del(__fsub_rn(mx[i], x[r]), Lx));
	atomicAdd(&a[GA_CY], (double)wrap_del(__fsub_rn(my[i], y[r]), Ly));
	atomicAdd(&a[GA_CZ], (double)wrap_del(__fsub_rn(mz[i], z[r]), Lz));
}

// kdReadCenter fallback (kd.c:1170-1181): mass-weighted offsets of the ORIGINAL positions
__global__ void __launch_bounds__(256)
    k_center_acc_com(int lo, int n, const int *gid, const int *repOrd, const float *x, const float *y, const float *z,
                     const float *mass, float Lx, float Ly, float Lz, double *acc)
{
	int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int g = gid[i];
	if (g <= 0) return;
	int r = repOrd[g];
	float m = mass[i];
	double *a = acc + (size_t)g * GA_STRIDE;
	atomicAdd(&a[GA_CX], (double)__fmul_rn(m, wrap_del(__fsub_rn(x[i], x[r]), Lx)));
	atomicAdd(&a[GA_CY], (double)__fmul_rn(m, wrap_del(__fsub_rn(y[i], y[r]), Ly)));
	atomicAdd(&a[GA_CZ], (double)__fmul_rn(m, wrap_del(__fsub_rn(z[i], z[r]), Lz)));
}

struct CatArgs {
	int nGroup, n;
	const int *repOrd;
	const int *gN;
	const double *acc;
	const float *x, *y, *z;
	float L[3];
	double wrapLo[3], wrapHi[3];
	int mode; // 0: divide centre sums by nMembers (kdCalcCenter), 1: by mass (COM), 2: keep given rCenter/vcm
};

__global__ void __launch_bounds__(256) k_catalogue(const CatArgs a, skidgpu_pgroup *cat)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= a.nGroup) return;
	skidgpu_pgroup p = cat[g];
	if (g == 0) {
		memset(&p, 0, sizeof p);
		p.nMembers = a.gN[0];
		cat[0] = p;
		return;
	}
	const double *ac = a.acc + (size_t)g * GA_STRIDE;
	int r = a.repOrd[g];
	p.nMembers = a.gN[g];
	p.fMass = (float)ac[GA_MASS];
	p.fRadius = 0.0f;
	p.pStart = p.pCurr = 0;
	float rel[3] = {0, 0, 0};
	if (r >= 0 && r < a.n) {
		rel[0] = a.x[r];
		rel[1] = a.y[r];
		rel[2] = a.z[r];
	}
	for (int j = 0; j < 3; ++j) {
		p.rel[j] = rel[j];
		p.rBound[j] = 0.0f;
	}
	if (a.mode != 2) {
		for (int j = 0; j < 3; ++j) {
			float cj = (float)ac[GA_CX + j];
			if (a.mode == 0) cj = __fdiv_rn(cj, (float)p.nMembers); // kd.c:1053
			else cj = __fdiv_rn(cj, p.fMass);                       // kd.c:1184
			cj = __fadd_rn(cj, rel[j]);
			if ((double)cj > a.wrapHi[j]) cj = __fsub_rn(cj, a.L[j]);
			if ((double)cj <= a.wrapLo[j]) cj = __fadd_rn(cj, a.L[j]);
			p.rCenter[j] = cj;
			p.vcm[j] = __fdiv_rn((float)ac[GA_MVX + j], p.fMass); // kd.c:1059
		}
	}
	cat[g] = p;
}

static void group_counts_and_catalogue(skidgpu_ctx &c, int mode, bool computeRep)
{
	cudaStream_t s = c.stream;
	const int n = c.n, G = c.nGroup;
	int *gN = c.gN.alloc(G + 1);
	double *acc = c.gAcc.alloc((size_t)(G + 1) * GA_STRIDE);
	CK(cudaMemsetAsync(gN, 0, sizeof(int) * (G + 1), s));
	CK(cudaMemsetAsync(acc, 0, sizeof(double) * (size_t)(G + 1) * GA_STRIDE, s));
	if (computeRep) {
		int *rep = c.repOrd.alloc(G + 1);
		CK(cudaMemsetAsync(rep, 0x7f, sizeof(int) * (G + 1), s));
		SK_LAUNCH(k_rep_min, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, rep);
	}
	// several GPUs: every rank accumulates its slice of the particles (and the movers it owns), the per-group sums
	// are all-reduced - the replicated version cost 7 ms at 2^27 / 8 GPUs against 0.8 ms on one
	const int lo = (int)((long long)n * c.rank / c.nranks), hi = (int)((long long)n * (c.rank + 1) / c.nranks);
	if (hi > lo)
		SK_LAUNCH(k_group_acc, (unsigned)ceil_div(hi - lo, 256), 256, 0, s, lo, hi, c.gid.p, c.mass.p, c.vx.p, c.vy.p, c.vz.p, gN,
		          acc);
	if (mode == 0 && c.nMove > 0)
		SK_LAUNCH(k_center_acc_movers, (unsigned)ceil_div(c.nMove, 256), 256, 0, s, c.nMove, c.mOrd.p, c.gid.p,
		          c.repOrd.p, c.mx.p, c.my.p, c.mz.p, c.x.p, c.y.p, c.z.p, c.L[0], c.L[1], c.L[2], acc, MOVE_OWN_BLOCK, c.rank,
		          c.nranks);
	if (mode == 1 && hi > lo)
		SK_LAUNCH(k_center_acc_com, (unsigned)ceil_div(hi - lo, 256), 256, 0, s, lo, hi, c.gid.p, c.repOrd.p, c.x.p, c.y.p, c.z.p,
		          c.mass.p, c.L[0], c.L[1], c.L[2], acc);
	sk_reduce(c, gN, G + 1, SK_I32, SK_SUM);
	sk_reduce(c, acc, (long long)(G + 1) * GA_STRIDE, SK_F64, SK_SUM);
	// members of group 0 = everything else (kd.c:996)
	std::vector<int> hN(G);
	CK(cudaMemcpyAsync(hN.data(), gN, sizeof(int) * G, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	long long tot = 0;
	for (int g = 1; g < G; ++g) tot += hN[g];
	int n0 = (int)((long long)n - tot);
	CK(cudaMemcpyAsync(gN, &n0, sizeof(int), cudaMemcpyHostToDevice, s));
	CatArgs ca;
	ca.nGroup = G;
	ca.n = n;
	ca.repOrd = c.repOrd.p;
	ca.gN = gN;
	ca.acc = acc;
	ca.x = c.x.p;
	ca.y = c.y.p;
	ca.z = c.z.p;
	for (int d = 0; d < 3; ++d) {
		ca.L[d] = c.L[d];
		ca.wrapHi[d] = (double)c.C[d] + 0.5 * (double)c.L[d];
		ca.wrapLo[d] = (double)c.C[d] - 0.5 * (double)c.L[d];
	}
	ca.mode = mode;
	SK_LAUNCH(k_catalogue, (unsigned)ceil_div(G, 256), 256, 0, s, ca, c.gCat.p);
	c.haveCenters = true;
}

void stage_centers(skidgpu_ctx &c)
{
	if (c.nGroup < 1 || !c.gid.p) throw SkidError("skidgpu_centers: skidgpu_fof has not run");
	StageTimer tm(c, 4);
	c.gCat.alloc(c.nGroup + 1);
	CK(cudaMemsetAsync(c.gCat.p, 0, sizeof(skidgpu_pgroup) * (c.nGroup + 1), c.stream));
	group_counts_and_catalogue(c, 0, false);
	tm.stop();
}

void stage_set_groups(skidgpu_ctx &c, const int *piGroup, int nGroup, const skidgpu_pgroup *centres)
{
	if (c.n <= 0) throw SkidError("skidgpu_set_groups: no particles set");
	if (nGroup < 1) throw SkidError("skidgpu_set_groups: nGroup must be >= 1");
	cudaStream_t s = c.stream;
	c.nGroup = nGroup;
	c.nMove = 0;
	CK(cudaMemcpyAsync(c.gid.alloc(c.n), piGroup, sizeof(int) * (size_t)c.n, cudaMemcpyHostToDevice, s));
	c.gCat.alloc(nGroup + 1);
	if (centres) CK(cudaMemcpyAsync(c.gCat.p, centres, sizeof(skidgpu_pgroup) * nGroup, cudaMemcpyHostToDevice, s));
	else CK(cudaMemsetAsync(c.gCat.p, 0, sizeof(skidgpu_pgroup) * (nGroup + 1), s));
	group_counts_and_catalogue(c, centres ? 2 : 1, true);
	CK(cudaStreamSynchronize(s));
}

// =====================================================================================
// unbinding
// =====================================================================================
// grouped / ungrouped flags and their compaction (ascending iOrder): only the grouped particles are sorted by label
__global__ void __launch_bounds__(256) k_label_flags(int n, const int *gid, int wantGrouped, uint32_t *flags)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flags[i] = ((gid[i] > 0) == (wantGrouped != 0)) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
    k_label_compact(int n, const int *gid, const uint32_t *flags, const uint32_t *scan, uint64_t *keys, uint32_t *idx)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !flags[i]) return;
	const uint32_t k = scan[i];
	if (keys) keys[k] = (uint64_t)(uint32_t)gid[i];
	idx[k] = (uint32_t)i;
}

// group-ordered member arrays: relative coordinates (kd.c:1341-1355)
struct MemArgs {
	int n, n0;
	const uint32_t *order; // the grouped particles sorted by group: file indices
	const int *gid;
	const skidgpu_pgroup *cat;
	const float *x, *y, *z, *vx, *vy, *vz, *mass, *soft;
	float hx, hy, hz;
	float4 *qr; // (dx,dy,dz,soft)
	float4 *qv; // (vx,vy,vz,mass)
	int *qord;
	float fEps; // < 0: keep per-particle softening
	int rank, nranks;
};

__global__ void __launch_bounds__(256) k_members(const MemArgs a)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= a.n - a.n0) return;
	uint32_t i = a.order[k];
	int g = a.gid[i];
	if (g % a.nranks != a.rank) return; // only the owner of a group reads its members (potentials, removal loop)
	const float *rel = a.cat[g].rel;
	float dx = __fsub_rn(a.x[i], rel[0]), dy = __fsub_rn(a.y[i], rel[1]), dz = __fsub_rn(a.z[i], rel[2]);
	float tx = __fmul_rn(2.0f, a.hx), ty = __fmul_rn(2.0f, a.hy), tz = __fmul_rn(2.0f, a.hz);
	if (dx > a.hx) dx = __fsub_rn(dx, tx);
	if (dx <= -a.hx) dx = __fadd_rn(dx, tx);
	if (dy > a.hy) dy = __fsub_rn(dy, ty);
	if (dy <= -a.hy) dy = __fadd_rn(dy, ty);
	if (dz > a.hz) dz = __fsub_rn(dz, tz);
	if (dz <= -a.hz) dz = __fadd_rn(dz, tz);
	a.qr[k] = make_float4(dx, dy, dz, a.soft[i]);
	a.qv[k] = make_float4(a.vx[i], a.vy[i], a.vz[i], a.mass[i]);
	a.qord[k] = (int)i;
}

// SPLINE_POT (grav.h:11-31) / Plummer (grav.c:24-26): float64 evaluation rounded to float32 "dir"
__device__ __forceinline__ float soft_dir(float d2, float twoh, int iSoftType)
{
	double a;
	if (iSoftType == SKIDGPU_PLUMMER) {
		a = 1.0 / sqrt((double)d2 + 0.25 * (double)twoh * (double)twoh);
	} else {
		double r = sqrt((double)d2);
		if (r < (double)twoh) {
			double dih = 2.0 / (double)twoh;
			double u = r * dih;
			if (u < 1.0) {
				a = dih * (7.0 / 5.0 - 2.0 / 3.0 * u * u + 3.0 / 10.0 * u * u * u * u -
				           1.0 / 10.0 * u * u * u * u * u);
			} else {
				double dir = 1.0 / r;
				a = -1.0 / 15.0 * dir + dih * (8.0 / 5.0 - 4.0 / 3.0 * u * u + u * u * u -
				                               3.0 / 10.0 * u * u * u * u + 1.0 / 30.0 * u * u * u * u * u);
			}
		} else {
			a = 1.0 / r;
		}
	}
	return (float)a;
}

// Far field of SPLINE_POT (r >= twoh): dir = (float)(1.0/sqrt((double)d2)) in the reference.  ncu on the
// massive-halo box (config 5): k_group_pot is issue bound and the double sqrt + double divide of every
// pair were ~50 of its ~70 instructions.  MUFU.RSQ + two FMA-residual Newton steps give the correctly
// rounded float in all but ~1e-3 of the cases (<= 1 ulp otherwise) without touching the FP64 pipe; pairs
// inside the softening (r < twoh, where d2 < twoh^2 decides exactly like the reference's double
// compare up to ties on the boundary, at which both branches agree) keep the double evaluation.
__device__ __forceinline__ float far_dir(float d2)
{
	float y;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d2)); // d2 >= twoh^2: never denormal in practice (ftz -> inf -> near path values only if twoh = 0)
	float g = d2 * y, h = 0.5f * y;
	float r = fmaf(-g, h, 0.5f);
	y = fmaf(y, r, y);
	g = d2 * y, h = 0.5f * y;
	r = fmaf(-g, h, 0.5f);
	return fmaf(y, r, y);
}

// scoop sources (grav.c:63-135): ungrouped particles within fScoop of rCenter (periodic), found in
// the tree over the ungrouped particles.  One warp per group.  MODE 0 counts, MODE 1 fills.
struct ScoopArgs {
	TreeView tv;
	const float4 *posS; // sorted ungrouped (x,y,z,mass)
	int nS;
	int nGroup;
	const skidgpu_pgroup *cat;
	float L[3], hL[3];
	float fBall2;
	uint32_t *cnt;
	const uint32_t *start;
	uint32_t *list;
	int rank, nranks;
	const int *iord; // non-null: the tree holds ALL particles (the kNN tree); sources are those with label 0
	const int *gid;
};

template <int MODE> __global__ void __launch_bounds__(256) k_scoop(const ScoopArgs a)
{
	const int lane = threadIdx.x & 31;
	const int g = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + 1;
	if (g >= a.nGroup || g % a.nranks != a.rank) return; // cnt[] was zeroed: foreign groups keep an empty list
	const uint32_t lt = (1u << lane) - 1u;
	const float x0 = a.cat[g].rCenter[0], y0 = a.cat[g].rCenter[1], z0 = a.cat[g].rCenter[2];
	const float xp = __fadd_rn(x0, a.L[0]), xm = __fsub_rn(x0, a.L[0]);
	const float yp = __fadd_rn(y0, a.L[1]), ym = __fsub_rn(y0, a.L[1]);
	const float zp = __fadd_rn(z0, a.L[2]), zm = __fsub_rn(z0, a.L[2]);
	uint32_t count = 0;
	const uint32_t outBase = MODE ? a.start[g] : 0;
	int lev = a.tv.top - 1;
	uint32_t node = 0, mymask = 0;
#define SCOOP_TEST()                                                                                   \
	{                                                                                              \
		const float4 *bx = a.tv.box[lev] + 2 * ((size_t)node * 32 + lane);                     \
		float4 lo = bx[0], hi = bx[1];                                                         \
		float d = dist2_rn(axis_gap_periodic(x0, xp, xm, lo.x, hi.x),                          \
		                   axis_gap_periodic(y0, yp, ym, lo.y, hi.y),                          \
		                   axis_gap_periodic(z0, zp, zm, lo.z, hi.z));                         \
		uint32_t m_ = __ballot_sync(SK_FULL, d <= a.fBall2);                                   \
		if (lane == lev) mymask = m_;                                                          \
	}
	if (a.nS > 0) {
		SCOOP_TEST();
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
				SCOOP_TEST();
				continue;
			}
			int idx = (int)child * 32 + lane;
			bool hit = false;
			if (idx < a.nS) {
				float4 p = a.posS[idx];
				float d2 = dist2_rn(minimg_dx(x0, xp, xm, a.hL[0], p.x), minimg_dx(y0, yp, ym, a.hL[1], p.y),
				                    minimg_dx(z0, zp, zm, a.hL[2], p.z));
				hit = d2 < a.fBall2; // grav.c:102
				if (hit && a.iord) hit = a.gid[a.iord[idx]] == 0;
			}
			uint32_t hm = __ballot_sync(SK_FULL, hit);
			if (MODE && hit) a.list[outBase + count + __popc(hm & lt)] = (uint32_t)idx;
			count += __popc(hm);
		}
	}
#undef SCOOP_TEST
	if (!MODE && lane == 0) a.cnt[g] = count;
}

// Potential of every group member: pairs within the group (kdCellPot) + scoop sources (kdAddScoopPot).
// One block of POT_T threads per tile I of POT_T members.  Like the reference's i < j loop (grav.c:14-34) every
// pair is evaluated ONCE: the block walks the tiles J >= I of its group; the diagonal tile is a plain loop, an
// off-diagonal tile is swept by rotation (thread i meets column (i + t) mod POT_T at step t), the row terms go
// to a register and the column terms to a per-warp shared array, so no two lanes touch the same column at the
// same step.  Per-pair arithmetic is the reference's (float32 geometry, dir, float32 product G*m*dir, float64
// sums); a tile's partial sums reach the float64 potential with one atomic per member.  (Storing the tile twice in a
// row so that the rotation index never wraps - no index arithmetic in the loop - was measured: no gain, 447 -> 462 ms.)
// ncu on the massive-halo box (config 5, profiles/r02_*): the two-sided loop that evaluated every pair twice was
// 523 ms of a 1025 ms pass (5.4e11 warp instructions, issue bound).
constexpr int POT_T = 128;
// per-tile partial sums stay in float64 like the reference's accumulator (grav.c:31-32): the loosely bound small
// groups sit at E ~ 0, where a float32 partial sum (1e-7 relative) flips removals - measured on the 2^21 box:
// 560 672 particles unbound instead of the reference's 567 192 (float64: 567 19x)
typedef double pot_acc_t;

struct PotArgs {
	int nGroup;
	const uint32_t *tileStart; // [nGroup+1] exclusive scan of tiles per group (index 0 unused = 0 tiles)
	const int *gStart;         // [nGroup+1] member offsets into q arrays (relative to n0)
	const float4 *qr, *qv;
	double *pot;
	// scoop
	const uint32_t *scStart; // [nGroup+1]
	const uint32_t *scList;
	const float4 *posS;      // (x,y,z,mass) of the scoop tree's points
	const float *softS;      // their softening, or ...
	const int *srcIdx;       // ... non-null: softening = soft[srcIdx[point]] (scoop from the kNN tree)
	const float *soft;
	const skidgpu_pgroup *cat;
	float L[3];
	float G;
	int iSoftType;
	int nMaxMembers;
};

__device__ __forceinline__ float pair_dir(float d2, float twoh, bool spline, int iSoftType)
{
	float dir = far_dir(d2);
	if (!(spline && d2 >= fmaxf(twoh * twoh, 1.0e-30f))) dir = soft_dir(d2, twoh, iSoftType); // inside the softening, or Plummer
	return dir;
}

__global__ void __launch_bounds__(POT_T) k_group_pot(const PotArgs a)
{
	__shared__ float4 s_r[POT_T];
	__shared__ float s_m[POT_T];
	__shared__ double s_col[POT_T / 32][POT_T];
	__shared__ int s_g;
	const int tid = threadIdx.x, w = tid >> 5;
	if (tid == 0) {
		// find the group of this tile: largest g with tileStart[g] <= blockIdx.x
		int lo = 1, hi = a.nGroup - 1;
		while (lo < hi) {
			int mid = (lo + hi + 1) >> 1;
			if (a.tileStart[mid] <= blockIdx.x) lo = mid;
			else hi = mid - 1;
		}
		s_g = lo;
	}
	__syncthreads();
	const int g = s_g;
	const int beg = a.gStart[g], n = a.gStart[g + 1] - beg;
	if (n >= a.nMaxMembers) return; // kd.c:1330
	const int I = (int)(blockIdx.x - a.tileStart[g]);
	const int i = I * POT_T + tid;
	const bool act = i < n;
	const bool spline = a.iSoftType != SKIDGPU_PLUMMER;
	float4 ri = make_float4(0, 0, 0, 0);
	float gmi = 0.0f;
	if (act) {
		ri = a.qr[beg + i];
		gmi = __fmul_rn(a.G, a.qv[beg + i].w); // kd->G*p[i].fMass (grav.c:31-32), float
	}
	double pot = 0.0;
	const int nT = (n + POT_T - 1) / POT_T;
	for (int J = I; J < nT; ++J) {
		const int j = J * POT_T + tid;
		if (j < n) {
			s_r[tid] = a.qr[beg + j];
			s_m[tid] = __fmul_rn(a.G, a.qv[beg + j].w);
		} else {
			s_r[tid] = make_float4(3.0e18f, 3.0e18f, 3.0e18f, 0.0f); // far away, massless
			s_m[tid] = 0.0f;
		}
#pragma unroll
		for (int c = 0; c < POT_T / 32; ++c) s_col[w][c * 32 + (tid & 31)] = 0.0;
		__syncthreads();
		pot_acc_t rowp = 0;
		if (J == I) { // diagonal tile: every member against every other member of the same tile
			const int lim = n - J * POT_T < POT_T ? n - J * POT_T : POT_T;
			if (act) {
#pragma unroll 4
				for (int t = 0; t < lim; ++t) {
					const float4 rj = s_r[t];
					const float dx = __fsub_rn(ri.x, rj.x), dy = __fsub_rn(ri.y, rj.y), dz = __fsub_rn(ri.z, rj.z);
					const float d2 = dist2_rn(dx, dy, dz);
					const float dir = t == tid ? 0.0f : pair_dir(d2, __fadd_rn(ri.w, rj.w), spline, a.iSoftType);
					rowp += (pot_acc_t)__fmul_rn(s_m[t], dir);
				}
			}
		} else { // off-diagonal tile: each pair once, both sides credited
			double *col = s_col[w];
#pragma unroll 4
			for (int t = 0; t < POT_T; ++t) {
				const int c = (tid + t) & (POT_T - 1);
				const float4 rj = s_r[c];
				const float dx = __fsub_rn(ri.x, rj.x), dy = __fsub_rn(ri.y, rj.y), dz = __fsub_rn(ri.z, rj.z);
				const float d2 = dist2_rn(dx, dy, dz);
				const float dir = pair_dir(d2, __fadd_rn(ri.w, rj.w), spline, a.iSoftType);
				rowp += (pot_acc_t)__fmul_rn(s_m[c], dir);        // padding columns are massless
				col[c] += (double)__fmul_rn(gmi, dir);            // inactive rows are massless
				__syncwarp();
			}
		}
		pot += (double)rowp;
		__syncthreads();
		if (J != I) {
			double cs = 0.0;
#pragma unroll
			for (int ww = 0; ww < POT_T / 32; ++ww) cs += s_col[ww][tid];
			if (j < n) atomicAdd(&a.pot[beg + j], cs);
		}
		__syncthreads();
	}
	// scoop sources (grav.c:107-131)
	const uint32_t sb = a.scStart[g], sn = a.scStart[g + 1] - sb;
	const float *rel = a.cat[g].rel;
	for (uint32_t j0 = 0; j0 < sn; j0 += POT_T) {
		uint32_t j = j0 + tid;
		if (j < sn) {
			uint32_t sidx = a.scList[sb + j];
			float4 p = a.posS[sidx];
			float nx = __fsub_rn(p.x, rel[0]), ny = __fsub_rn(p.y, rel[1]), nz = __fsub_rn(p.z, rel[2]);
			nx = wrap_del(nx, a.L[0]);
			ny = wrap_del(ny, a.L[1]);
			nz = wrap_del(nz, a.L[2]);
			s_r[tid] = make_float4(nx, ny, nz, a.srcIdx ? a.soft[a.srcIdx[sidx]] : a.softS[sidx]);
			s_m[tid] = __fmul_rn(a.G, p.w);
		}
		__syncthreads();
		int lim = (int)(sn - j0 < (uint32_t)POT_T ? sn - j0 : (uint32_t)POT_T);
		if (act) {
			pot_acc_t rowp = 0;
#pragma unroll 4
			for (int t = 0; t < lim; ++t) {
				const float4 rj = s_r[t];
				const float dx = __fsub_rn(rj.x, ri.x), dy = __fsub_rn(rj.y, ri.y), dz = __fsub_rn(rj.z, ri.z);
				const float d2 = dist2_rn(dx, dy, dz);
				rowp += (pot_acc_t)__fmul_rn(s_m[t], pair_dir(d2, __fadd_rn(rj.w, ri.w), spline, a.iSoftType));
			}
			pot += (double)rowp;
		}
		__syncthreads();
	}
	if (act) atomicAdd(&a.pot[beg + i], pot); // other blocks add their column sums to the same entry
}

// The removal loop of kdUnbind (kd.c:1360-1457).  One block per group.
struct UnbArgs {
	int nGroup;
	const int *gStart;
	float4 *qr, *qv;
	int *qord;
	double *pot;
	int *gid;
	skidgpu_pgroup *cat;
	int *gN;
	float fShift, fCosmo, z, G;
	float hx;
	int iSoftType, bNoUnbind, nMaxMembers, bSubPot;
	unsigned int *nUnbound;
	unsigned long long *nPairs;
	int rank, nranks;
	int nSmallMax; // groups of this many members or more are left to k_unbind_cl
};

constexpr int UNB_TINY = 128;   // < this many members: 64 threads per group
constexpr int UNB_MID = 1024;   // >= this many members: one block of UCL_T threads per group (k_unbind_cl<1>)
constexpr int UNB_BIG = 16384;  // >= this many: a thread-block cluster of UCL_NB blocks per group (k_unbind_cl<UCL_NB>)
constexpr int UCL_T = 1024;
constexpr int UCL_NB = 8;

// T threads per group; groups of nLo <= n < nHi members; SM = capacity of the shared-memory copy.  Most groups hold
// a few dozen members: with 256 threads for each, six of eight warps only ran the reductions (5e9 warp
// instructions at 2^24); they now get 64.
template <int T, int SM> __global__ void __launch_bounds__(T) k_unbind(const UnbArgs a, const int nLo, const int nHi)
{
	const int g = blockIdx.x + 1;
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const int beg = a.gStart[g];
	int n = a.gStart[g + 1] - beg;
	if (n >= a.nMaxMembers || n <= 0 || n < nLo || n >= nHi || (g % a.nranks) != a.rank) return;
	float4 *qr = a.qr + beg, *qv = a.qv + beg;
	int *qord = a.qord + beg;
	double *pot = a.pot + beg;
	// Small groups (almost all of them) are unbound out of shared memory: the swap and centre-of-mass update of a
	// removal are a chain of dependent accesses by one thread, ~1.5 us per removal out of global memory.
	__shared__ float4 s_qr[SM], s_qv[SM];
	__shared__ double s_pot[SM];
	__shared__ int s_ord[SM];
	if (n <= SM) {
		for (int i = tid; i < n; i += T) {
			s_qr[i] = qr[i];
			s_qv[i] = qv[i];
			s_pot[i] = pot[i];
			s_ord[i] = qord[i];
		}
		__syncthreads();
		qr = s_qr;
		qv = s_qv;
		pot = s_pot;
		qord = s_ord;
	}
	__shared__ double s_red[T / 32][7];
	__shared__ double s_cm[7]; // dMass, rcm[3], vcm[3]
	__shared__ float s_best[T / 32], s_least[T / 32];
	__shared__ int s_bi[T / 32], s_li[T / 32];
	__shared__ int s_iBig, s_iMin, s_stop;

	// centre of mass (kd.c:1360-1375), float64 sums of float32 products
	double acc[7] = {0, 0, 0, 0, 0, 0, 0};
	for (int i = tid; i < n; i += T) {
		float4 r = qr[i], v = qv[i];
		acc[0] += (double)v.w;
		acc[1] += (double)__fmul_rn(v.w, r.x);
		acc[2] += (double)__fmul_rn(v.w, r.y);
		acc[3] += (double)__fmul_rn(v.w, r.z);
		acc[4] += (double)__fmul_rn(v.w, v.x);
		acc[5] += (double)__fmul_rn(v.w, v.y);
		acc[6] += (double)__fmul_rn(v.w, v.z);
	}
#pragma unroll
	for (int k = 0; k < 7; ++k) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(SK_FULL, acc[k], o);
		if (lane == 0) s_red[w][k] = acc[k];
	}
	__syncthreads();
	if (tid == 0) {
		double t[7] = {0, 0, 0, 0, 0, 0, 0};
		for (int ww = 0; ww < T / 32; ++ww)
			for (int k = 0; k < 7; ++k) t[k] += s_red[ww][k];
		s_cm[0] = t[0];
		for (int k = 1; k < 7; ++k) s_cm[k] = t[k] / t[0];
	}
	__syncthreads();

	int nRemoved = 0;
	int iMinFinal = 0;
	while (true) {
		// energy scan (kd.c:1389-1408)
		const double rcx = s_cm[1], rcy = s_cm[2], rcz = s_cm[3], vcx = s_cm[4], vcy = s_cm[5], vcz = s_cm[6];
		float best = -1.0f, least = 1.0f;
		int bi = 0x7fffffff, li = 0x7fffffff;
		for (int i = tid; i < n; i += T) {
			float4 r = qr[i], v = qv[i];
			float dvx = (float)((double)a.fShift * ((double)v.x - vcx) + (double)a.fCosmo * ((double)r.x - rcx));
			float dvy = (float)((double)a.fShift * ((double)v.y - vcy) + (double)a.fCosmo * ((double)r.y - rcy));
			float dvz = (float)((double)a.fShift * ((double)v.z - vcz) + (double)a.fCosmo * ((double)r.z - rcz));
			float dv2 = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dvx, dvx)), __fmul_rn(dvy, dvy)),
			                      __fmul_rn(dvz, dvz));
			float fTot = (float)(0.5 * (double)dv2 - pot[i] * (1.0 + (double)a.z));
			if (fTot > best) {
				best = fTot;
				bi = i;
			}
			if (fTot < least) {
				least = fTot;
				li = i;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			float ob = __shfl_xor_sync(SK_FULL, best, o);
			int obi = __shfl_xor_sync(SK_FULL, bi, o);
			if (ob > best || (ob == best && obi < bi)) {
				best = ob;
				bi = obi;
			}
			float ol = __shfl_xor_sync(SK_FULL, least, o);
			int oli = __shfl_xor_sync(SK_FULL, li, o);
			if (ol < least || (ol == least && oli < li)) {
				least = ol;
				li = oli;
			}
		}
		if (lane == 0) {
			s_best[w] = best;
			s_bi[w] = bi;
			s_least[w] = least;
			s_li[w] = li;
		}
		__syncthreads();
		if (tid == 0) {
			float b = s_best[0], l = s_least[0];
			int ib = s_bi[0], il = s_li[0];
			for (int ww = 1; ww < T / 32; ++ww) {
				if (s_best[ww] > b || (s_best[ww] == b && s_bi[ww] < ib)) {
					b = s_best[ww];
					ib = s_bi[ww];
				}
				if (s_least[ww] < l || (s_least[ww] == l && s_li[ww] < il)) {
					l = s_least[ww];
					il = s_li[ww];
				}
			}
			if (ib == 0x7fffffff) ib = 0; // nothing exceeded the initial -1.0 (kd.c:1389-1390)
			if (il == 0x7fffffff) il = 0;
			s_iBig = ib;
			s_iMin = il;
			int stop = (b < 0.0f || a.bNoUnbind) ? 1 : 0; // kd.c:1409
			if (!stop) {
				// unbind particle iBig (kd.c:1413-1440)
				int nn = n - 1;
				a.gid[qord[ib]] = 0;
				if (nn == 0) {
					s_cm[0] = 0.0;
					s_cm[4] = s_cm[5] = s_cm[6] = 0.0;
					stop = 2;
				} else {
					float4 r = qr[ib], v = qv[ib];
					double dM = s_cm[0] - (double)v.w;
					s_cm[0] = dM;
					double f = (double)v.w / dM;
					s_cm[1] += f * (s_cm[1] - (double)r.x);
					s_cm[2] += f * (s_cm[2] - (double)r.y);
					s_cm[3] += f * (s_cm[3] - (double)r.z);
					s_cm[4] += f * (s_cm[4] - (double)v.x);
					s_cm[5] += f * (s_cm[5] - (double)v.y);
					s_cm[6] += f * (s_cm[6] - (double)v.z);
					// swap iBig <-> last
					float4 tr = qr[nn], tv = qv[nn];
					int to = qord[nn];
					double tp = pot[nn];
					qr[nn] = r;
					qv[nn] = v;
					qord[nn] = qord[ib];
					pot[nn] = pot[ib];
					qr[ib] = tr;
					qv[ib] = tv;
					qord[ib] = to;
					pot[ib] = tp;
				}
			}
			s_stop = stop;
		}
		__syncthreads();
		const int stop = s_stop;
		iMinFinal = s_iMin;
		if (stop == 1) break;
		--n;
		++nRemoved;
		if (stop == 2) break;
		if (a.bSubPot) { // kdSubPot (grav.c:39-60), only for pure dark / pure star inputs (kd.c:1441)
			float4 rs = qr[n];
			float ms = qv[n].w;
			for (int i = tid; i < n; i += T) {
				float4 r = qr[i];
				float dx = __fsub_rn(rs.x, r.x), dy = __fsub_rn(rs.y, r.y), dz = __fsub_rn(rs.z, r.z);
				float d2 = dist2_rn(dx, dy, dz);
				float twoh = __fadd_rn(rs.w, r.w);
				float dir = far_dir(d2);
				if (!(a.iSoftType != SKIDGPU_PLUMMER && d2 >= fmaxf(twoh * twoh, 1.0e-30f))) dir = soft_dir(d2, twoh, a.iSoftType);
				pot[i] -= (double)__fmul_rn(__fmul_rn(a.G, ms), dir);
			}
		}
		__syncthreads();
	}
	if (tid == 0) {
		skidgpu_pgroup *pg = &a.cat[g];
		pg->fMass = (float)s_cm[0];
		pg->vcm[0] = (float)s_cm[4];
		pg->vcm[1] = (float)s_cm[5];
		pg->vcm[2] = (float)s_cm[6];
		float4 rb = qr[iMinFinal];
		float rr[3] = {rb.x, rb.y, rb.z};
		float t2 = __fmul_rn(2.0f, a.hx);
		for (int j = 0; j < 3; ++j) { // kd.c:1452-1457 (hx for all three axes, as the reference)
			float dx = __fadd_rn(rr[j], pg->rel[j]);
			if (dx > a.hx) dx = __fsub_rn(dx, t2);
			if (dx <= -a.hx) dx = __fadd_rn(dx, t2);
			pg->rBound[j] = dx;
		}
		pg->nMembers = n;
		a.gN[g] = n;
		if (nRemoved) {
			atomicAdd(a.nUnbound, (unsigned int)nRemoved);
			atomicAdd(&a.gN[0], nRemoved);
		}
	}
}

// The same loop for large groups.  One block per group is a single SM chasing memory latency (measured: 96 us
// per removal in a 39 k-member group); here a group gets UCL_T threads, or a thread-block CLUSTER of NB blocks
// whose partial arg-max / arg-min meet through distributed shared memory, and a removal costs ONE pass over the
// members: the removed particle's potential is subtracted (kdSubPot) and the new energy is formed in the same
// sweep, with rcm/vcm already updated (they only depend on the removed particle, kd.c:1430-1434).  The order of
// removals, the tie rule (first index of the largest energy, kd.c:1400-1403) and the swap-to-end bookkeeping are
// the reference's.  Members are re-read through L2 (__ldcg): the swap is written by another SM of the cluster.
template <int NB> __device__ __forceinline__ float4 ucl_ld4(const float4 *p) { return NB > 1 ? __ldcg(p) : *p; }
template <int NB> __device__ __forceinline__ double ucl_ldd(const double *p) { return NB > 1 ? __ldcg(p) : *p; }

template <int NB> __global__ void __launch_bounds__(UCL_T) k_unbind_cl(const UnbArgs a, const int *list)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cl = cg::this_cluster();
	const int brank = NB > 1 ? (int)cl.block_rank() : 0;
	const int g = list[blockIdx.x / NB];
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	constexpr int NW = UCL_T / 32;
	const int NT = NB * UCL_T, t0 = brank * UCL_T + tid;
	const int beg = a.gStart[g];
	int n = a.gStart[g + 1] - beg;
	float4 *qr = a.qr + beg, *qv = a.qv + beg;
	int *qord = a.qord + beg;
	double *pot = a.pot + beg;
	__shared__ double s_red[NW][7];
	__shared__ double s_part[7];
	__shared__ double s_cm[7]; // dMass, rcm[3], vcm[3]
	__shared__ float s_wv[2][NW];
	__shared__ int s_wi[2][NW];
	__shared__ float s_xv[2][2]; // [parity][best, least] of this block, read by the other blocks of the cluster
	__shared__ int s_xi[2][2];
	__shared__ float s_fv;
	__shared__ int s_fi[2];
	__shared__ float4 s_rem[2];

	// centre of mass (kd.c:1360-1375), float64 sums of float32 products
	{
		double acc[7] = {0, 0, 0, 0, 0, 0, 0};
		for (int i = t0; i < n; i += NT) {
			const float4 r = qr[i], v = qv[i];
			acc[0] += (double)v.w;
			acc[1] += (double)__fmul_rn(v.w, r.x);
			acc[2] += (double)__fmul_rn(v.w, r.y);
			acc[3] += (double)__fmul_rn(v.w, r.z);
			acc[4] += (double)__fmul_rn(v.w, v.x);
			acc[5] += (double)__fmul_rn(v.w, v.y);
			acc[6] += (double)__fmul_rn(v.w, v.z);
		}
#pragma unroll
		for (int k = 0; k < 7; ++k) {
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(SK_FULL, acc[k], o);
			if (lane == 0) s_red[w][k] = acc[k];
		}
		__syncthreads();
		if (tid < 7) {
			double t = 0.0;
			for (int ww = 0; ww < NW; ++ww) t += s_red[ww][tid];
			s_part[tid] = t;
		}
		if (NB > 1) cl.sync();
		else __syncthreads();
		if (tid == 0) {
			double t[7] = {0, 0, 0, 0, 0, 0, 0};
			for (int r = 0; r < NB; ++r) {
				const double *rp = NB > 1 ? cl.map_shared_rank(s_part, r) : s_part;
				for (int k = 0; k < 7; ++k) t[k] += rp[k];
			}
			s_cm[0] = t[0];
			for (int k = 1; k < 7; ++k) s_cm[k] = t[k] / t[0];
		}
		__syncthreads();
	}

	int nRemoved = 0, iMinFinal = 0, par = 0;
	bool sub = false;
	float4 rs = make_float4(0, 0, 0, 0);
	float gms = 0.0f;
	while (true) {
		// one sweep: (kdSubPot of the particle removed last, grav.c:39-60) + energy scan (kd.c:1389-1408)
		const double rcx = s_cm[1], rcy = s_cm[2], rcz = s_cm[3], vcx = s_cm[4], vcy = s_cm[5], vcz = s_cm[6];
		float best = -1.0f, least = 1.0f;
		int bi = 0x7fffffff, li = 0x7fffffff;
		for (int i = t0; i < n; i += NT) {
			const float4 r = ucl_ld4<NB>(qr + i), v = ucl_ld4<NB>(qv + i);
			double p = ucl_ldd<NB>(pot + i);
			if (sub) {
				const float dx = __fsub_rn(rs.x, r.x), dy = __fsub_rn(rs.y, r.y), dz = __fsub_rn(rs.z, r.z);
				const float d2 = dist2_rn(dx, dy, dz);
				p -= (double)__fmul_rn(gms, pair_dir(d2, __fadd_rn(rs.w, r.w), a.iSoftType != SKIDGPU_PLUMMER, a.iSoftType));
				pot[i] = p;
			}
			const float dvx = (float)((double)a.fShift * ((double)v.x - vcx) + (double)a.fCosmo * ((double)r.x - rcx));
			const float dvy = (float)((double)a.fShift * ((double)v.y - vcy) + (double)a.fCosmo * ((double)r.y - rcy));
			const float dvz = (float)((double)a.fShift * ((double)v.z - vcz) + (double)a.fCosmo * ((double)r.z - rcz));
			const float dv2 = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dvx, dvx)), __fmul_rn(dvy, dvy)), __fmul_rn(dvz, dvz));
			const float fTot = (float)(0.5 * (double)dv2 - p * (1.0 + (double)a.z));
			if (fTot > best) {
				best = fTot;
				bi = i;
			}
			if (fTot < least) {
				least = fTot;
				li = i;
			}
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const float ob = __shfl_xor_sync(SK_FULL, best, o);
			const int obi = __shfl_xor_sync(SK_FULL, bi, o);
			if (ob > best || (ob == best && obi < bi)) {
				best = ob;
				bi = obi;
			}
			const float ol = __shfl_xor_sync(SK_FULL, least, o);
			const int oli = __shfl_xor_sync(SK_FULL, li, o);
			if (ol < least || (ol == least && oli < li)) {
				least = ol;
				li = oli;
			}
		}
		if (lane == 0) {
			s_wv[0][w] = best;
			s_wi[0][w] = bi;
			s_wv[1][w] = least;
			s_wi[1][w] = li;
		}
		__syncthreads();
		if (w == 0) {
			best = s_wv[0][lane];
			bi = s_wi[0][lane];
			least = s_wv[1][lane];
			li = s_wi[1][lane];
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const float ob = __shfl_xor_sync(SK_FULL, best, o);
				const int obi = __shfl_xor_sync(SK_FULL, bi, o);
				if (ob > best || (ob == best && obi < bi)) {
					best = ob;
					bi = obi;
				}
				const float ol = __shfl_xor_sync(SK_FULL, least, o);
				const int oli = __shfl_xor_sync(SK_FULL, li, o);
				if (ol < least || (ol == least && oli < li)) {
					least = ol;
					li = oli;
				}
			}
			if (lane == 0) {
				s_xv[par][0] = best;
				s_xi[par][0] = bi;
				s_xv[par][1] = least;
				s_xi[par][1] = li;
			}
		}
		if (NB > 1) {
			cl.sync();
			if (w == 0) { // the blocks' partial results through distributed shared memory
				const int r = lane < NB ? lane : 0;
				best = *cl.map_shared_rank(&s_xv[par][0], r);
				bi = *cl.map_shared_rank(&s_xi[par][0], r);
				least = *cl.map_shared_rank(&s_xv[par][1], r);
				li = *cl.map_shared_rank(&s_xi[par][1], r);
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) {
					const float ob = __shfl_xor_sync(SK_FULL, best, o);
					const int obi = __shfl_xor_sync(SK_FULL, bi, o);
					if (ob > best || (ob == best && obi < bi)) {
						best = ob;
						bi = obi;
					}
					const float ol = __shfl_xor_sync(SK_FULL, least, o);
					const int oli = __shfl_xor_sync(SK_FULL, li, o);
					if (ol < least || (ol == least && oli < li)) {
						least = ol;
						li = oli;
					}
				}
			}
		}
		if (tid == 0) {
			if (bi == 0x7fffffff) bi = 0; // nothing exceeded the initial -1.0 (kd.c:1389-1390)
			if (li == 0x7fffffff) li = 0;
			s_fv = best;
			s_fi[0] = bi;
			s_fi[1] = li;
			if (!(best < 0.0f || a.bNoUnbind)) { // every block keeps the removed particle before it is swapped away
				s_rem[0] = ucl_ld4<NB>(qr + bi);
				s_rem[1] = ucl_ld4<NB>(qv + bi);
			}
		}
		__syncthreads();
		const float fBig = s_fv;
		const int iBig = s_fi[0];
		iMinFinal = s_fi[1];
		if (fBig < 0.0f || a.bNoUnbind) break; // kd.c:1409
		rs = s_rem[0];
		const float4 vs = s_rem[1];
		gms = __fmul_rn(a.G, vs.w);
		if (NB > 1) cl.sync(); // all blocks have read particle iBig
		const int nn = n - 1;
		if (brank == 0 && tid == 0) { // unbind particle iBig (kd.c:1413-1440): label, swap with the last member
			a.gid[qord[iBig]] = 0;
			if (nn > 0 && iBig != nn) {
				const float4 tr = qr[nn], tv = qv[nn];
				const int to = qord[nn];
				const double tp = pot[nn];
				qr[nn] = rs;
				qv[nn] = vs;
				qord[nn] = qord[iBig];
				pot[nn] = pot[iBig];
				qr[iBig] = tr;
				qv[iBig] = tv;
				qord[iBig] = to;
				pot[iBig] = tp;
			}
			__threadfence();
		}
		if (tid == 0) { // every block updates its copy of the centre of mass with the same arithmetic
			if (nn == 0) {
				s_cm[0] = 0.0;
				s_cm[4] = s_cm[5] = s_cm[6] = 0.0;
			} else {
				const double dM = s_cm[0] - (double)vs.w;
				s_cm[0] = dM;
				const double f = (double)vs.w / dM;
				s_cm[1] += f * (s_cm[1] - (double)rs.x);
				s_cm[2] += f * (s_cm[2] - (double)rs.y);
				s_cm[3] += f * (s_cm[3] - (double)rs.z);
				s_cm[4] += f * (s_cm[4] - (double)vs.x);
				s_cm[5] += f * (s_cm[5] - (double)vs.y);
				s_cm[6] += f * (s_cm[6] - (double)vs.z);
			}
		}
		n = nn;
		++nRemoved;
		par ^= 1;
		sub = a.bSubPot != 0; // kdSubPot only for pure dark / pure star inputs (kd.c:1441)
		if (NB > 1) cl.sync(); // the swap is visible to every block
		else __syncthreads();
		if (nn == 0) break;
	}
	if (brank == 0 && tid == 0) {
		skidgpu_pgroup *pg = &a.cat[g];
		pg->fMass = (float)s_cm[0];
		pg->vcm[0] = (float)s_cm[4];
		pg->vcm[1] = (float)s_cm[5];
		pg->vcm[2] = (float)s_cm[6];
		const float4 rb = qr[iMinFinal];
		const float rr[3] = {rb.x, rb.y, rb.z};
		const float t2 = __fmul_rn(2.0f, a.hx);
		for (int j = 0; j < 3; ++j) { // kd.c:1452-1457 (hx for all three axes, as the reference)
			float dx = __fadd_rn(rr[j], pg->rel[j]);
			if (dx > a.hx) dx = __fsub_rn(dx, t2);
			if (dx <= -a.hx) dx = __fadd_rn(dx, t2);
			pg->rBound[j] = dx;
		}
		pg->nMembers = n;
		a.gN[g] = n;
		if (nRemoved) {
			atomicAdd(a.nUnbound, (unsigned int)nRemoved);
			atomicAdd(&a.gN[0], nRemoved);
		}
	}
	if (NB > 1) cl.sync(); // nobody leaves while its shared memory may still be read
}

// groups by size class for the removal loop: list[0 ..) = one block each, list[nGroup ..) = one cluster each
__global__ void __launch_bounds__(256)
    k_unbind_classes(int nGroup, const int *gN, int nMax, int rank, int nranks, int *list, unsigned int *counts)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup || g == 0 || (g % nranks) != rank) return;
	const int n = gN[g];
	if (n >= nMax || n < UNB_MID) return;
	if (n >= UNB_BIG) list[nGroup + atomicAdd(&counts[2], 1u)] = g;
	else list[atomicAdd(&counts[1], 1u)] = g;
}

// kdTooSmall (kd.c:1251-1293)
__global__ void __launch_bounds__(256) k_small_flags(int nGroup, const int *gN, int nMin, uint32_t *flags)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup) return;
	flags[g] = (g > 0 && gN[g] >= nMin) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
    k_small_map(int nGroup, const uint32_t *flags, const uint32_t *scan, const skidgpu_pgroup *cat, skidgpu_pgroup *out,
                int *map)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup) return;
	if (g == 0) {
		map[0] = 0;
		return;
	}
	if (flags[g]) {
		int ng = (int)scan[g] + 1;
		map[g] = ng;
		out[ng] = cat[g];
	} else map[g] = 0;
}
__global__ void __launch_bounds__(256) k_small_remap(int n, int *gid, const int *map, unsigned int *cnt0)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	bool zero = false;
	if (i < n) {
		int g = map[gid[i]];
		gid[i] = g;
		zero = g == 0;
	}
	uint32_t b = __ballot_sync(SK_FULL, zero);
	if ((threadIdx.x & 31) == 0 && b) atomicAdd(cnt0, (unsigned int)__popc(b));
}

// group radius for the .gtp (kd.c:1627-1644)
__global__ void __launch_bounds__(256)
    k_group_radius(int n, const int *gid, const float *x, const float *y, const float *z, float Lx, float Ly,
                   float Lz, skidgpu_pgroup *cat)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int g = gid[i];
	if (g <= 0) return;
	const float *rc = cat[g].rCenter;
	float dx = wrap_del(__fsub_rn(x[i], rc[0]), Lx);
	float dy = wrap_del(__fsub_rn(y[i], rc[1]), Ly);
	float dz = wrap_del(__fsub_rn(z[i], rc[2]), Lz);
	float f2 = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
	float fr = (float)sqrt((double)f2);
	atomicMax((unsigned int *)&cat[g].fRadius, __float_as_uint(fr));
}

// multi-GPU: group g is unbound by rank g % nranks; everybody else contributes nothing to the merge
__global__ void __launch_bounds__(256) k_tiles_per_group(int nGroup, const int *gN, int nMax, uint32_t *tiles,
                                                         int rank, int nranks)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup) return;
	int n = gN[g];
	tiles[g] = (g == 0 || n >= nMax || (g % nranks) != rank) ? 0u : (uint32_t)((n + POT_T - 1) / POT_T);
}

// catalogue rows that unbinding changes, packed for the cross-rank sum: owner's values, zeros elsewhere
__global__ void __launch_bounds__(256) k_cat_pack(int nGroup, const skidgpu_pgroup *cat, const int *gN, int rank,
                                                  int nranks, float *f)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup) return;
	float *o = f + (size_t)g * 8;
	bool own = g > 0 && (g % nranks) == rank;
	o[0] = own ? cat[g].fMass : 0.0f;
	for (int j = 0; j < 3; ++j) {
		o[1 + j] = own ? cat[g].vcm[j] : 0.0f;
		o[4 + j] = own ? cat[g].rBound[j] : 0.0f;
	}
	o[7] = 0.0f;
}
// member counts travel as integers (a float row entry is exact only up to 2^24 members)
__global__ void __launch_bounds__(256) k_count_pack(int nGroup, const int *gN, int rank, int nranks, int *out)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup) return;
	out[g] = (g > 0 && (g % nranks) == rank) ? gN[g] : 0;
}
__global__ void __launch_bounds__(256) k_cat_unpack(int nGroup, skidgpu_pgroup *cat, int *gN, const float *f, const int *cnt)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup || g == 0) return;
	const float *o = f + (size_t)g * 8;
	cat[g].fMass = o[0];
	for (int j = 0; j < 3; ++j) {
		cat[g].vcm[j] = o[1 + j];
		cat[g].rBound[j] = o[4 + j];
	}
	gN[g] = cnt[g];
	cat[g].nMembers = gN[g];
}

__global__ void __launch_bounds__(256) k_copy_counts(int nGroup, const int *gN, uint32_t *out)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g < nGroup) out[g] = (uint32_t)gN[g];
}

__global__ void __launch_bounds__(256) k_gstart_rel(int nGroup, const uint32_t *scan, int n0, int *gStart)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g <= nGroup) gStart[g] = (int)scan[g] - n0; // members of group g at q[gStart[g] .. gStart[g+1])
}

__global__ void __launch_bounds__(256)
    k_gather_scoop_src(int n0, const uint32_t *order, const float *x, const float *y, const float *z, float *ox,
                       float *oy, float *oz)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n0) return;
	uint32_t j = order[i];
	ox[i] = x[j];
	oy[i] = y[j];
	oz[i] = z[j];
}
__global__ void __launch_bounds__(256)
    k_gather_scoop_sorted(int n0, const uint32_t *perm, const uint32_t *order, const float *x, const float *y,
                          const float *z, const float *mass, const float *soft, float4 *posS, float *softS)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n0) return;
	uint32_t j = order[perm[i]];
	posS[i] = make_float4(x[j], y[j], z[j], mass[j]);
	softS[i] = soft[j];
}

// catalogue bookkeeping fields (pStart/pCurr, kd.c:1213-1219) from the exclusive scan of the member counts
__global__ void __launch_bounds__(256) k_set_pstart(int nGroup, const int *gN, const uint32_t *scan, skidgpu_pgroup *cat)
{
	int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= nGroup) return;
	cat[g].nMembers = gN[g];
	cat[g].pStart = (int)scan[g];
	cat[g].pCurr = (int)scan[g] + gN[g];
}

__global__ void __launch_bounds__(256) k_count_members(int n, const int *gid, int *gN)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int g = gid[i];
	// group 0 is by far the largest: aggregate it per warp
	uint32_t z = __ballot_sync(__activemask(), g == 0);
	if (g == 0) {
		if ((threadIdx.x & 31) == __ffs(z) - 1) atomicAdd(&gN[0], __popc(z));
	} else atomicAdd(&gN[g], 1);
}

void stage_unbind(skidgpu_ctx &c, float fG, float z, double fCosmoD, int iSoftType, float fScoop, int bNoUnbind,
                  int nMaxMembers, int nMinMembers, int *nUnboundOut, int *nGroupBeforeOut)
{
	cudaStream_t s = c.stream;
	const int n = c.n;
	if (!c.haveCenters) throw SkidError("skidgpu_unbind: skidgpu_centers / skidgpu_set_groups has not run");
	StageTimer tm(c, 5);
	const int G = c.nGroup;
	if (nGroupBeforeOut) *nGroupBeforeOut = G - 1;
	UnbindScratch &S = c.unbS; // persistent scratch (ctx.cuh)
	auto &order = S.order, &tiles = S.tiles, &tileStart = S.tileStart, &scCnt = S.scCnt, &scStart = S.scStart, &scList = S.scList,
	     &cntU = S.cntU;
	auto &keys = S.keys;
	auto &gStart = S.gStart, &qord = S.qord, &map = S.map;
	auto &qr = S.qr, &qv = S.qv, &posS = S.posS;
	auto &softS = S.softS, &sx = S.sx, &sy = S.sy, &sz = S.sz;
	auto &pot = S.pot;
	auto &dCnt = S.dCnt;
	auto &cat2 = S.cat2;
	BoxTree &treeS = S.treeS;
	unsigned int hUnbound = 0;

	if (G > 1) {
		// ---- kdGroupOrder: members of each group contiguous.  Only the grouped particles are sorted (by label,
		// stable: ascending iOrder within a group); most particles are ungrouped and keep their order.
		cntU.alloc(G + 2);
		uint32_t *scan = c.scan.alloc((size_t)(G > n ? G : n) + 64);
		uint32_t *flags = c.flags.alloc((size_t)n + 1);
		SK_LAUNCH(k_copy_counts, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gN.p, cntU.p);
		exclusive_scan_u32(cntU.p, scan, G, c.ws, s);
		int n0 = 0;
		CK(cudaMemcpyAsync(&n0, c.gN.p, sizeof(int), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		gStart.alloc(G + 2);
		SK_LAUNCH(k_gstart_rel, (unsigned)ceil_div(G + 1, 256), 256, 0, s, G, scan, n0, gStart.p);
		const int nm = n - n0; // grouped particles
		keys.alloc(nm > 0 ? nm : 1);
		order.alloc(nm > 0 ? nm : 1);
		if (nm > 0) {
			SK_LAUNCH(k_label_flags, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, 1, flags);
			exclusive_scan_u32(flags, scan, n, c.ws, s);
			SK_LAUNCH(k_label_compact, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, flags, scan, keys.p, order.p);
			int bits = 1;
			while ((1ll << bits) < (long long)G) ++bits;
			dist_sort_pairs(c, keys.p, order.p, nm, bits); // replicated input: shared between the ranks
		}

		// ---- the scoop sources (kd.c:1324-1325: a tree over the ungrouped particles).  When the kNN tree of the
		// density stage holds every particle (dark-only inputs, -gd) it serves as it is, with the label test at
		// the leaves; otherwise (other species mixes, the -unbind restart) a tree over the ungrouped is built.
		const bool reuseA = c.nAct == n && c.treeA.n == n && c.posA.p && c.iordA.p;
		if (!reuseA) {
			posS.alloc(n0 > 0 ? n0 : 1);
			softS.alloc(n0 > 0 ? n0 : 1);
		}
		if (!reuseA && n0 > 0) {
			auto &idx0 = S.idx0;
			idx0.alloc(n0);
			sx.alloc(n0);
			sy.alloc(n0);
			sz.alloc(n0);
			SK_LAUNCH(k_label_flags, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, 0, flags);
			exclusive_scan_u32(flags, scan, n, c.ws, s);
			SK_LAUNCH(k_label_compact, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, flags, scan, (uint64_t *)nullptr, idx0.p);
			SK_LAUNCH(k_gather_scoop_src, (unsigned)ceil_div(n0, 256), 256, 0, s, n0, idx0.p, c.x.p, c.y.p, c.z.p, sx.p,
			          sy.p, sz.p);
			tree_sort_points(treeS, sx.p, sy.p, sz.p, n0, c.ws, s, &c);
			SK_LAUNCH(k_gather_scoop_sorted, (unsigned)ceil_div(n0, 256), 256, 0, s, n0, treeS.perm.p, idx0.p, c.x.p,
			          c.y.p, c.z.p, c.mass.p, c.soft.p, posS.p, softS.p);
			tree_build_boxes(treeS, posS.p, nullptr, nullptr, n0, s);
		}
		ScoopArgs sc;
		sc.tv = tree_view(reuseA ? c.treeA : treeS);
		sc.posS = reuseA ? c.posA.p : posS.p;
		sc.nS = reuseA ? n : n0;
		sc.iord = reuseA ? c.iordA.p : nullptr;
		sc.gid = c.gid.p;
		sc.nGroup = G;
		sc.cat = c.gCat.p;
		for (int d = 0; d < 3; ++d) {
			sc.L[d] = c.L[d];
			sc.hL[d] = 0.5f * c.L[d];
		}
		sc.fBall2 = fScoop * fScoop; // grav.c:84
		sc.rank = c.rank;
		sc.nranks = c.nranks;
		scCnt.alloc(G + 2);
		scStart.alloc(G + 2);
		CK(cudaMemsetAsync(scCnt.p, 0, sizeof(uint32_t) * (G + 2), s));
		sc.cnt = scCnt.p;
		sc.start = nullptr;
		sc.list = nullptr;
		SK_LAUNCH(k_scoop<0>, (unsigned)ceil_div((size_t)(G - 1) * 32, 256), 256, 0, s, sc);
		exclusive_scan_u32(scCnt.p, scStart.p, G + 1, c.ws, s);
		uint32_t nScoop = 0;
		CK(cudaMemcpyAsync(&nScoop, scStart.p + G, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		scList.alloc(nScoop > 0 ? nScoop : 1);
		sc.start = scStart.p;
		sc.list = scList.p;
		if (nScoop > 0) SK_LAUNCH(k_scoop<1>, (unsigned)ceil_div((size_t)(G - 1) * 32, 256), 256, 0, s, sc);

		// ---- members in group order, relative coordinates
		qr.alloc(nm > 0 ? nm : 1);
		qv.alloc(nm > 0 ? nm : 1);
		qord.alloc(nm > 0 ? nm : 1);
		pot.alloc(nm > 0 ? nm : 1);
		const float hx = (float)(0.5 * (double)c.L[0]), hy = (float)(0.5 * (double)c.L[1]),
		            hz = (float)(0.5 * (double)c.L[2]);
		if (nm > 0) {
			MemArgs ma;
			ma.n = n;
			ma.n0 = n0;
			ma.order = order.p;
			ma.gid = c.gid.p;
			ma.cat = c.gCat.p;
			ma.x = c.x.p;
			ma.y = c.y.p;
			ma.z = c.z.p;
			ma.vx = c.vx.p;
			ma.vy = c.vy.p;
			ma.vz = c.vz.p;
			ma.mass = c.mass.p;
			ma.soft = c.soft.p;
			ma.hx = hx;
			ma.hy = hy;
			ma.hz = hz;
			ma.qr = qr.p;
			ma.qv = qv.p;
			ma.qord = qord.p;
			ma.fEps = -1.0f;
			ma.rank = c.rank;
			ma.nranks = c.nranks;
			SK_LAUNCH(k_members, (unsigned)ceil_div(nm, 256), 256, 0, s, ma);
			CK(cudaMemsetAsync(pot.p, 0, sizeof(double) * nm, s));

			// ---- potentials
			tiles.alloc(G + 2);
			tileStart.alloc(G + 2);
			SK_LAUNCH(k_tiles_per_group, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gN.p, nMaxMembers, tiles.p, c.rank,
			          c.nranks);
			exclusive_scan_u32(tiles.p, tileStart.p, G, c.ws, s);
			uint32_t nTiles = 0;
			CK(cudaMemcpyAsync(&nTiles, tileStart.p + G, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
			CK(cudaStreamSynchronize(s));
			PotArgs pa;
			pa.nGroup = G;
			pa.tileStart = tileStart.p;
			pa.gStart = gStart.p;
			pa.qr = qr.p;
			pa.qv = qv.p;
			pa.pot = pot.p;
			pa.scStart = scStart.p;
			pa.scList = scList.p;
			pa.posS = reuseA ? c.posA.p : posS.p;
			pa.softS = softS.p;
			pa.srcIdx = reuseA ? c.iordA.p : nullptr;
			pa.soft = c.soft.p;
			pa.cat = c.gCat.p;
			for (int d = 0; d < 3; ++d) pa.L[d] = c.L[d];
			pa.G = fG;
			pa.iSoftType = iSoftType;
			pa.nMaxMembers = nMaxMembers;
			if (nTiles > 0) SK_LAUNCH(k_group_pot, nTiles, POT_T, 0, s, pa);

			// ---- removal loop
			dCnt.alloc(4);
			CK(cudaMemsetAsync(dCnt.p, 0, sizeof(unsigned int) * 4, s));
			UnbArgs ua;
			ua.nGroup = G;
			ua.gStart = gStart.p;
			ua.qr = qr.p;
			ua.qv = qv.p;
			ua.qord = qord.p;
			ua.pot = pot.p;
			ua.gid = c.gid.p;
			ua.cat = c.gCat.p;
			ua.gN = c.gN.p;
			ua.fShift = (float)(1.0 / (1.0 + (double)z)); // kd.c:1317
			ua.fCosmo = (float)fCosmoD;                   // kd.c:1318 (float variable)
			ua.z = z;
			ua.G = fG;
			ua.hx = hx;
			ua.iSoftType = iSoftType;
			ua.bNoUnbind = bNoUnbind;
			ua.nMaxMembers = nMaxMembers;
			ua.bSubPot = (c.inType == SKIDGPU_DARK || c.inType == SKIDGPU_STAR) ? 1 : 0; // kd.c:1441
			ua.nUnbound = dCnt.p;
			ua.nPairs = nullptr;
			ua.rank = c.rank;
			ua.nranks = c.nranks;
			ua.nSmallMax = UNB_MID;
			// size classes: small groups one 256-thread block each, large ones 1024 threads, the largest a cluster
			auto &clsList = S.clsList;
			clsList.alloc(2 * (size_t)G + 2);
			SK_LAUNCH(k_unbind_classes, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gN.p, nMaxMembers, c.rank, c.nranks,
			          clsList.p, dCnt.p);
			unsigned int hCls[4] = {0, 0, 0, 0};
			CK(cudaMemcpyAsync(hCls, dCnt.p, sizeof hCls, cudaMemcpyDeviceToHost, s));
			CK(cudaStreamSynchronize(s));
			SK_LAUNCH((k_unbind<64, UNB_TINY>), (unsigned)(G - 1), 64, 0, s, ua, 1, UNB_TINY);
			SK_LAUNCH((k_unbind<256, 256>), (unsigned)(G - 1), 256, 0, s, ua, UNB_TINY, UNB_MID);
			if (hCls[1] > 0) SK_LAUNCH(k_unbind_cl<1>, hCls[1], UCL_T, 0, s, ua, (const int *)clsList.p);
			if (hCls[2] > 0) {
				cudaLaunchConfig_t cfg = {};
				cfg.gridDim = dim3(hCls[2] * UCL_NB);
				cfg.blockDim = dim3(UCL_T);
				cfg.stream = s;
				cudaLaunchAttribute at[1];
				at[0].id = cudaLaunchAttributeClusterDimension;
				at[0].val.clusterDim.x = UCL_NB;
				at[0].val.clusterDim.y = 1;
				at[0].val.clusterDim.z = 1;
				cfg.attrs = at;
				cfg.numAttrs = 1;
				CK(cudaLaunchKernelEx(&cfg, k_unbind_cl<UCL_NB>, ua, (const int *)(clsList.p + G)));
				++g_skid_launches;
			}
			if (c.nranks > 1) { // merge the shards: labels (owner wrote 0 for unbound members), rows, count
				auto &pack = S.pack;
				auto &cpack = S.cpack;
				pack.alloc((size_t)G * 8);
				cpack.alloc((size_t)G + 1);
				sk_reduce(c, c.gid.p, n, SK_I32, SK_MIN);
				SK_LAUNCH(k_cat_pack, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gCat.p, c.gN.p, c.rank, c.nranks, pack.p);
				SK_LAUNCH(k_count_pack, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gN.p, c.rank, c.nranks, cpack.p);
				sk_reduce(c, pack.p, (long long)G * 8, SK_F32, SK_SUM);
				sk_reduce(c, cpack.p, G, SK_I32, SK_SUM);
				SK_LAUNCH(k_cat_unpack, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gCat.p, c.gN.p, pack.p, cpack.p);
				sk_reduce(c, dCnt.p, 1, SK_I32, SK_SUM);
				CK(cudaStreamSynchronize(s));
			}
			CK(cudaMemcpyAsync(&hUnbound, dCnt.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
			CK(cudaStreamSynchronize(s));
		}
	}
	if (nUnboundOut) *nUnboundOut = (int)hUnbound;

	// ---- kdTooSmall (kd.c:1251-1293)
	{
		uint32_t *flags = c.flags.alloc((size_t)(G > n ? G : n) + 1);
		uint32_t *scan = c.scan.alloc((size_t)(G > n ? G : n) + 64);
		SK_LAUNCH(k_small_flags, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gN.p, nMinMembers, flags);
		exclusive_scan_u32(flags, scan, G, c.ws, s);
		uint32_t keep = 0;
		CK(cudaMemcpyAsync(&keep, scan + G, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		const int G2 = (int)keep + 1;
		cat2.alloc(G2 + 1);
		map.alloc(G + 1);
		CK(cudaMemsetAsync(cat2.p, 0, sizeof(skidgpu_pgroup) * (G2 + 1), s));
		SK_LAUNCH(k_small_map, (unsigned)ceil_div(G, 256), 256, 0, s, G, flags, scan, c.gCat.p, cat2.p, map.p);
		dCnt.alloc(4);
		CK(cudaMemsetAsync(dCnt.p, 0, sizeof(unsigned int) * 4, s));
		SK_LAUNCH(k_small_remap, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, map.p, dCnt.p);
		// new catalogue
		c.gCat.alloc(G2 + 1);
		CK(cudaMemcpyAsync(c.gCat.p, cat2.p, sizeof(skidgpu_pgroup) * G2, cudaMemcpyDeviceToDevice, s));
		c.nGroup = G2;
		// member counts of the compacted catalogue
		int *gN = c.gN.alloc(G2 + 1);
		CK(cudaMemsetAsync(gN, 0, sizeof(int) * (G2 + 1), s));
		double *acc = c.gAcc.alloc((size_t)(G2 + 1) * GA_STRIDE); // scratch only
		(void)acc;
		SK_LAUNCH(k_count_members, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, gN);
		exclusive_scan_u32((const uint32_t *)gN, scan, G2, c.ws, s); // counts are non-negative
		SK_LAUNCH(k_set_pstart, (unsigned)ceil_div(G2, 256), 256, 0, s, G2, gN, scan, c.gCat.p);
		SK_LAUNCH(k_group_radius, (unsigned)ceil_div(n, 256), 256, 0, s, n, c.gid.p, c.x.p, c.y.p, c.z.p, c.L[0], c.L[1],
		          c.L[2], c.gCat.p);
	}
	tm.stop();
}

