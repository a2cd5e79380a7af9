// Group statistics: the numbers behind the .stat file, replacing kdOutStats (kd.c:1703-1839; SURVEY §8f
// row 1).  The reference, per group: copies the members, makes their coordinates relative to rCenter
// (min-image), qsorts them by squared radius (CmpRadius kd.c:1690-1700) and runs ONE sequential float32
// loop (total/gas/star mass, max and half-mass circular velocity, velocity dispersion).
//
// Here: one 64-bit key per particle, (group << 32) | bits(r^2), sorted with the library's stable LSD radix
// sort (ties keep ascending iOrder), then one warp per group.  Float addition is not associative and the
// reference's sums are sequential, so the sums stay sequential - but only the ADDs: a chunk of 32 sorted
// members is loaded by the 32 lanes at once (gathers, square roots, divisions and the per-member terms
// are lane-parallel), the running total is carried through the chunk with one shuffle + one FADD per
// member, and the two data-dependent selections (max circular velocity, first crossing of the half mass)
// are decided by ballots, entering a sequential replay only for chunks in which something can change.
// Every expression keeps the reference's types: float products and sums without FMA contraction, double
// where a double literal (0.5, 4.0, 3.0) or sqrt() promotes the sub-expression.
#include "ctx.cuh"

namespace {

struct StatArgs {
	int n, nGroup;
	const uint32_t *start; // start[g]: offset of group g's members in the sorted arrays (group 0 first)
	const uint64_t *keys;  // sorted
	const uint32_t *vals;  // sorted: particle index (iOrder)
	const int *gid;
	const float *x, *y, *z, *vx, *vy, *vz, *mass, *soft, *temp, *rho; // rho may be null (-unbind restart)
	const skidgpu_pgroup *cat;
	int nGas, nDark;
	float hx, hy, hz;
	float G, fExp, fExpHub, fDensMin, fTempMax;
	skidgpu_stat_row *rows;
};

// kd.c:1750-1758: float compare against the float half period, float 2*h
__device__ __forceinline__ float stat_wrap(float d, float h)
{
	const float twoh = __fmul_rn(2.0f, h);
	if (d > h) d = __fsub_rn(d, twoh);
	if (d <= -h) d = __fadd_rn(d, twoh);
	return d;
}

__global__ void __launch_bounds__(256) k_stat_keys(const StatArgs a, uint64_t *keys, uint32_t *vals)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= a.n) return;
	const int g = a.gid[i];
	uint64_t key = 0;
	if (g > 0) {
		const float *rc = a.cat[g].rCenter;
		const float dx = stat_wrap(__fsub_rn(a.x[i], rc[0]), a.hx);
		const float dy = stat_wrap(__fsub_rn(a.y[i], rc[1]), a.hy);
		const float dz = stat_wrap(__fsub_rn(a.z[i], rc[2]), a.hz);
		// radius2 = 0.0; radius2 += r[k]*r[k] (kd.c:1770-1775)
		const float r2 = __fadd_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(dx, dx)), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
		key = ((uint64_t)(uint32_t)g << 32) | (uint64_t)__float_as_uint(r2); // r2 >= 0: bit order == value order
	}
	keys[i] = key;
	vals[i] = (uint32_t)i;
}

constexpr int ST_WARPS = 8;

__global__ void __launch_bounds__(ST_WARPS * 32) k_stat_groups(const StatArgs a)
{
	__shared__ double sCand[ST_WARPS][32];
	__shared__ double sVcD[ST_WARPS][32];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int g = 1 + blockIdx.x * ST_WARPS + w;
	if (g >= a.nGroup) return;
	const uint32_t s0 = a.start[g], s1 = a.start[g + 1];
	const int nm = (int)(s1 - s0);
	skidgpu_stat_row row;
	memset(&row, 0, sizeof row);
	if (nm <= 0) {
		if (lane == 0) a.rows[g] = row;
		return;
	}
	const float rcx = a.cat[g].rCenter[0], rcy = a.cat[g].rCenter[1], rcz = a.cat[g].rCenter[2];
	const float vcx = a.cat[g].vcm[0], vcy = a.cat[g].vcm[1], vcz = a.cat[g].vcm[2];

	// ---- "Quick, we need the half mass first" (kd.c:1780-1781): fHalfMass += 0.5*m, sequential.
	// 0.5*m is exact and the double sum of two floats rounds to the same float as a float add.
	float fHalf = 0.0f;
	for (int base = 0; base < nm; base += 32) {
		const int cnt = min(32, nm - base);
		float hm = 0.0f;
		if (lane < cnt) hm = __fmul_rn(0.5f, a.mass[a.vals[s0 + base + lane]]);
		for (int j = 0; j < cnt; ++j) fHalf = __fadd_rn(fHalf, __shfl_sync(SK_FULL, hm, j));
	}

	float fTot = 0.0f, fGas = 0.0f, fStar = 0.0f, fVdisp = 0.0f;
	float fVcirc = 0.0f, fmVcirc = 0.0f, fRVmax = 0.0f, fRhmass = 0.0f;
	double curD = 0.0; // (double)fVcirc
	float r2Last = 0.0f;
	for (int base = 0; base < nm; base += 32) {
		const int cnt = min(32, nm - base);
		const bool live = lane < cnt;
		float m = 0.0f, r2 = 0.0f, d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
		bool isGas = false, isStar = false, outside = false;
		double sq = 1.0;
		if (live) {
			const uint32_t p = a.vals[s0 + base + lane];
			r2 = __uint_as_float((uint32_t)(a.keys[s0 + base + lane] & 0xffffffffull));
			m = a.mass[p];
			const float so = a.soft[p];
			const float dx = stat_wrap(__fsub_rn(a.x[p], rcx), a.hx);
			const float dy = stat_wrap(__fsub_rn(a.y[p], rcy), a.hy);
			const float dz = stat_wrap(__fsub_rn(a.z[p], rcz), a.hz);
			// dv = fExp*(v - vcm) + fExpHub*r; fVdisp += dv*dv (kd.c:1800-1805)
			const float e0 = __fadd_rn(__fmul_rn(a.fExp, __fsub_rn(a.vx[p], vcx)), __fmul_rn(a.fExpHub, dx));
			const float e1 = __fadd_rn(__fmul_rn(a.fExp, __fsub_rn(a.vy[p], vcy)), __fmul_rn(a.fExpHub, dy));
			const float e2 = __fadd_rn(__fmul_rn(a.fExp, __fsub_rn(a.vz[p], vcz)), __fmul_rn(a.fExpHub, dz));
			d0 = __fmul_rn(e0, e0);
			d1 = __fmul_rn(e1, e1);
			d2 = __fmul_rn(e2, e2);
			isGas = (int)p < a.nGas && (a.rho ? a.rho[p] : 0.0f) >= a.fDensMin && a.temp[p] <= a.fTempMax;
			isStar = (int)p >= a.nGas + a.nDark;
			outside = (double)r2 > 4.0 * (double)so * (double)so; // kd.c:1787
			sq = sqrt((double)r2);
		}
		// running total through the chunk: lane j keeps the value after member j was added
		float myTot = 0.0f;
		for (int j = 0; j < cnt; ++j) {
			fTot = __fadd_rn(fTot, __shfl_sync(SK_FULL, m, j));
			if (lane == j) myTot = fTot;
		}
		const uint32_t gasMask = __ballot_sync(SK_FULL, live && isGas);
		const uint32_t starMask = __ballot_sync(SK_FULL, live && isStar);
		for (uint32_t mk = gasMask; mk; mk &= mk - 1) fGas = __fadd_rn(fGas, __shfl_sync(SK_FULL, m, __ffs(mk) - 1));
		for (uint32_t mk = starMask; mk; mk &= mk - 1) fStar = __fadd_rn(fStar, __shfl_sync(SK_FULL, m, __ffs(mk) - 1));
		for (int j = 0; j < cnt; ++j) {
			fVdisp = __fadd_rn(fVdisp, __shfl_sync(SK_FULL, d0, j));
			fVdisp = __fadd_rn(fVdisp, __shfl_sync(SK_FULL, d1, j));
			fVdisp = __fadd_rn(fVdisp, __shfl_sync(SK_FULL, d2, j));
		}
		// lane-parallel candidates: G*fTotMass/sqrt(r2) in double for the test, the float pair that is stored
		const float gm = __fmul_rn(a.G, myTot);
		const float rv = (float)sq;                  // fRVmax / fRhmass = sqrt(fBall2) rounded to float
		const float vc = live ? __fdiv_rn(gm, rv) : 0.0f; // G*fTotMass/fRVmax, all float
		const double cand = live ? (double)gm / sq : 0.0;
		// max circular velocity outside 2 softenings (kd.c:1787-1791): sequential replay only if some
		// member of the chunk beats the current value (nothing can change otherwise)
		if (__any_sync(SK_FULL, live && outside && cand > curD)) {
			sCand[w][lane] = (live && outside) ? cand : -1.0;
			sVcD[w][lane] = (double)vc;
			__syncwarp();
			int jAcc = -1;
			for (int j = 0; j < cnt; ++j) {
				if (sCand[w][j] > curD) {
					curD = sVcD[w][j];
					jAcc = j;
				}
			}
			__syncwarp();
			if (jAcc >= 0) {
				fVcirc = __shfl_sync(SK_FULL, vc, jAcc);
				fRVmax = __shfl_sync(SK_FULL, rv, jAcc);
			}
		}
		// half-mass radius (kd.c:1796-1799): first member with fTotMass > fHalfMass while fmVcirc == 0
		if (fmVcirc == 0.0f) {
			uint32_t mk = __ballot_sync(SK_FULL, live && myTot > fHalf);
			while (mk && fmVcirc == 0.0f) {
				const int j = __ffs(mk) - 1;
				fRhmass = __shfl_sync(SK_FULL, rv, j);
				fmVcirc = __shfl_sync(SK_FULL, vc, j);
				mk &= mk - 1;
			}
		}
		r2Last = __shfl_sync(SK_FULL, r2, cnt - 1);
	}
	// outer circular velocity (kd.c:1807-1811)
	const double sqLast = sqrt((double)r2Last);
	const float flVcirc = (float)((double)__fmul_rn(a.G, fTot) / sqLast);
	if (fVcirc == 0.0f) {
		fVcirc = flVcirc;
		fRVmax = (float)sqLast;
	}
	if (lane == 0) {
		row.nMembers = nm;
		row.fTotMass = fTot;
		row.fGasMass = fGas;
		row.fStarMass = fStar;
		row.fVcirc = fVcirc;
		row.fmVcirc = fmVcirc;
		row.flVcirc = flVcirc;
		row.fRVmax = fRVmax;
		row.fRhmass = fRhmass;
		row.fRouter2 = r2Last;
		row.fVdispSum = fVdisp;
		a.rows[g] = row;
	}
}

__global__ void __launch_bounds__(256) k_stat_counts(int nGroup, const int *gN, uint32_t *out)
{
	const int g = blockIdx.x * blockDim.x + threadIdx.x;
	if (g < nGroup) out[g] = (uint32_t)gN[g];
}

} // namespace

void stage_stats(skidgpu_ctx &c, float fG, float z, double dExpHub, float fDensMin, float fTempMax,
                 skidgpu_stat_row *hostRows)
{
	cudaStream_t s = c.stream;
	const int n = c.n, G = c.nGroup;
	if (G < 1 || !c.gid.p || !c.gCat.p || !c.gN.p)
		throw SkidError("skidgpu_stats: no group catalogue (run skidgpu_unbind first)");
	if (G == 1 || n == 0) {
		if (hostRows) memset(hostRows, 0, sizeof(skidgpu_stat_row) * (size_t)G);
		return;
	}
	DevBuf<uint64_t> keys;
	DevBuf<uint32_t> vals, cnt, start;
	DevBuf<skidgpu_stat_row> rows;
	keys.alloc(n);
	vals.alloc(n);
	cnt.alloc(G + 2);
	start.alloc(G + 2);
	rows.alloc(G);
	StatArgs a;
	a.n = n;
	a.nGroup = G;
	a.gid = c.gid.p;
	a.x = c.x.p;
	a.y = c.y.p;
	a.z = c.z.p;
	a.vx = c.vx.p;
	a.vy = c.vy.p;
	a.vz = c.vz.p;
	a.mass = c.mass.p;
	a.soft = c.soft.p;
	a.temp = c.temp.p;
	// density stage did not run (-unbind restart): fDensity reads as 0; after an initial cut (-fic) the cut
	// scatterers read as 0 too (smooth1.c:463-470)
	a.rho = c.nAct > 0 ? (c.haveRhoStat ? c.rhoStat.p : c.rho.p) : nullptr;
	a.cat = c.gCat.p;
	a.nGas = c.nGas;
	a.nDark = c.nDark;
	a.hx = (float)(0.5 * (double)c.L[0]); // kd.c:1738-1740
	a.hy = (float)(0.5 * (double)c.L[1]);
	a.hz = (float)(0.5 * (double)c.L[2]);
	a.G = fG;
	a.fExp = (float)(1.0 / (1.0 + (double)z)); // kd.c:1730
	a.fExpHub = (float)dExpHub;                // kd.c:1731
	a.fDensMin = fDensMin;
	a.fTempMax = fTempMax;
	a.keys = keys.p;
	a.vals = vals.p;
	a.start = start.p;
	a.rows = rows.p;
	CK(cudaMemsetAsync(rows.p, 0, sizeof(skidgpu_stat_row) * (size_t)G, s));
	SK_LAUNCH(k_stat_keys, (unsigned)ceil_div(n, 256), 256, 0, s, a, keys.p, vals.p);
	int gbits = 1;
	while ((1ll << gbits) < (long long)G) ++gbits;
	radix_sort_pairs(keys.p, vals.p, n, 32 + gbits, c.ws, s);
	SK_LAUNCH(k_stat_counts, (unsigned)ceil_div(G, 256), 256, 0, s, G, c.gN.p, cnt.p);
	exclusive_scan_u32(cnt.p, start.p, G, c.ws, s);
	SK_LAUNCH(k_stat_groups, (unsigned)ceil_div(G - 1, ST_WARPS), ST_WARPS * 32, 0, s, a);
	if (hostRows)
		CK(cudaMemcpyAsync(hostRows, rows.p, sizeof(skidgpu_stat_row) * (size_t)G, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s)); // before the local buffers are released
}
