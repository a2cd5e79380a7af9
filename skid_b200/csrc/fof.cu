// Stage 4: friends-of-friends grouping of the converged mover positions.
//
// Replaces kdFoF (kd.c:802-917; INTERCONT kd.h:217-335): connected components of the graph
// {min-image d2 < tau^2, float32, strict} over ALL movers.  The reference walks a kd-tree with a
// BFS FIFO; converged clumps are extremely dense (hundreds of movers within fCvg), so edges are
// never enumerated here.  Movers are binned in a uniform grid of cell size <= 0.57 tau
// (< tau/sqrt(3)): two movers in one cell are always linked, so the union-find runs over CELLS,
// and two neighbouring cells (|offset| <= R, R = 2 normally) are united by the first mover pair
// found within tau.  Group ids are canonical: groups are numbered by ascending smallest member
// iOrder (the reference numbers them in mover-tree traversal order, an artefact - SURVEY D5).
#include "ctx.cuh"

struct GridSpec {
	double lo[3], inv[3];
	int nc[3];
	int wrap[3];
	int R;
};

__global__ void __launch_bounds__(256) k_cell_keys(int m, const float *x, const float *y, const float *z,
                                                   GridSpec g, uint64_t *keys, uint32_t *idx)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	float p[3] = {x[i], y[i], z[i]};
	long long c[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		long long v = (long long)floor(((double)p[d] - g.lo[d]) * g.inv[d]);
		if (g.wrap[d]) { // a mover outside the box (input not pre-wrapped, a wrong -c) belongs to the periodic image of its cell
			v %= g.nc[d];
			if (v < 0) v += g.nc[d];
		} else {
			if (v < 0) v = 0;
			if (v > g.nc[d] - 1) v = g.nc[d] - 1;
		}
		c[d] = v;
	}
	keys[i] = ((uint64_t)c[2] * (uint64_t)g.nc[1] + (uint64_t)c[1]) * (uint64_t)g.nc[0] + (uint64_t)c[0];
	idx[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_cell_heads(int m, const uint64_t *keys, uint32_t *flags)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// cellStart[c] = first sorted mover of cell c; cellKey[c]; moverCell[i] = cell of sorted mover i
__global__ void __launch_bounds__(256)
    k_cell_fill(int m, const uint64_t *keys, const uint32_t *flags, const uint32_t *scan, uint32_t *cellStart,
                uint64_t *cellKey, uint32_t *moverCell, int nCells)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t c = scan[i] + flags[i] - 1; // inclusive scan - 1
	moverCell[i] = c;
	if (flags[i]) {
		cellStart[c] = (uint32_t)i;
		cellKey[c] = keys[i];
	}
	if (i == 0) cellStart[nCells] = (uint32_t)m;
}

__global__ void __launch_bounds__(256)
    k_gather_sorted_movers(int m, const uint32_t *idx, const float *x, const float *y, const float *z,
                           const int *mOrd, float4 *spos)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t j = idx[i];
	spos[i] = make_float4(x[j], y[j], z[j], __int_as_float(mOrd[j]));
}

#define HASH_EMPTY 0xffffffffffffffffull
__device__ __forceinline__ uint32_t hash64(uint64_t k)
{
	k ^= k >> 33;
	k *= 0xff51afd7ed558ccdull;
	k ^= k >> 33;
	k *= 0xc4ceb9fe1a85ec53ull;
	k ^= k >> 33;
	return (uint32_t)k;
}

__global__ void __launch_bounds__(256) k_hash_insert(int nCells, const uint64_t *cellKey, uint64_t *hkeys,
                                                     uint32_t *hvals, uint32_t hmask)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nCells) return;
	uint64_t k = cellKey[c];
	uint32_t h = hash64(k) & hmask;
	while (true) {
		unsigned long long old = atomicCAS((unsigned long long *)&hkeys[h], HASH_EMPTY, (unsigned long long)k);
		if (old == HASH_EMPTY || old == k) {
			hvals[h] = (uint32_t)c;
			return;
		}
		h = (h + 1) & hmask;
	}
}

__device__ __forceinline__ int hash_lookup(uint64_t k, const uint64_t *hkeys, const uint32_t *hvals, uint32_t hmask)
{
	uint32_t h = hash64(k) & hmask;
	while (true) {
		uint64_t v = hkeys[h];
		if (v == k) return (int)hvals[h];
		if (v == HASH_EMPTY) return -1;
		h = (h + 1) & hmask;
	}
}

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x)
{
	while (true) {
		uint32_t p = ((volatile uint32_t *)parent)[x];
		if (p == x) return x;
		uint32_t gp = ((volatile uint32_t *)parent)[p];
		if (gp != p) parent[x] = gp; // path halving (benign race: only ever points to an ancestor)
		x = p;
	}
}

__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b)
{
	while (true) {
		a = uf_find(parent, a);
		b = uf_find(parent, b);
		if (a == b) return;
		uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
		uint32_t old = atomicCAS(&parent[hi], hi, lo);
		if (old == hi) return;
	}
}

__global__ void __launch_bounds__(256) k_uf_init(int n, uint32_t *parent)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) parent[i] = (uint32_t)i;
}

// Bounding box of every cell's movers (one warp per cell): lets k_link_cells decide most cell pairs without
// looking at a single mover pair.
__global__ void __launch_bounds__(256) k_cell_boxes(int nCells, const uint32_t *cellStart, const float4 *spos, float4 *cellBox)
{
	const int lane = threadIdx.x & 31;
	const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (c >= nCells) return;
	const uint32_t beg = cellStart[c], end = cellStart[c + 1];
	float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (uint32_t i = beg + lane; i < end; i += 32) {
		const float4 p = spos[i];
		lo[0] = fminf(lo[0], p.x), hi[0] = fmaxf(hi[0], p.x);
		lo[1] = fminf(lo[1], p.y), hi[1] = fmaxf(hi[1], p.y);
		lo[2] = fminf(lo[2], p.z), hi[2] = fmaxf(hi[2], p.z);
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			lo[d] = fminf(lo[d], __shfl_xor_sync(SK_FULL, lo[d], o));
			hi[d] = fmaxf(hi[d], __shfl_xor_sync(SK_FULL, hi[d], o));
		}
	if (lane == 0) {
		cellBox[2 * (size_t)c] = make_float4(lo[0], lo[1], lo[2], 0.0f);
		cellBox[2 * (size_t)c + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
	}
}

constexpr int FOF_UNROLL = 4; // rounds of 32 mover pairs in flight in the pair scan of k_link_cells (1: 6.6 ms, 4: 5.9 ms, 8: 5.8 ms FoF stage at 2^24)
struct LinkArgs {
	int nCells;
	const uint32_t *cellStart;
	const uint64_t *cellKey;
	const float4 *spos;
	const float4 *cellBox; // [2c] = lo, [2c+1] = hi of the movers of cell c
	const uint64_t *hkeys;
	const uint32_t *hvals;
	uint32_t hmask;
	uint32_t *parent;
	GridSpec g;
	float L[3], hL[3];
	float fTau2;
};

// One warp per cell A.  Half stencil: offsets (dx,dy,dz) lexicographically > 0 in (dz,dy,dx).
__global__ void __launch_bounds__(256) k_link_cells(const LinkArgs a)
{
	const int lane = threadIdx.x & 31;
	const int A = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (A >= a.nCells) return;
	const uint64_t key = a.cellKey[A];
	const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
	const int cx = (int)(key % (uint64_t)nx);
	const int cy = (int)((key / (uint64_t)nx) % (uint64_t)ny);
	const int cz = (int)(key / ((uint64_t)nx * (uint64_t)ny));
	const int R = a.g.R, W = 2 * R + 1;
	const int nOff = W * W * W;
	const uint32_t aBeg = a.cellStart[A], aEnd = a.cellStart[A + 1];
	const uint32_t nA = aEnd - aBeg;
	const float4 alo = a.cellBox[2 * (size_t)A], ahi = a.cellBox[2 * (size_t)A + 1];
	for (int base = nOff / 2 + 1; base < nOff; base += 32) {
		int o = base + lane;
		int B = -1;
		if (o < nOff) {
			int ox = o % W - R, oy = (o / W) % W - R, oz = o / (W * W) - R;
			int bx = cx + ox, by = cy + oy, bz = cz + oz;
			bool ok = true;
			if (a.g.wrap[0]) bx = ((bx % nx) + nx) % nx;
			else ok = ok && bx >= 0 && bx < nx;
			if (a.g.wrap[1]) by = ((by % ny) + ny) % ny;
			else ok = ok && by >= 0 && by < ny;
			if (a.g.wrap[2]) bz = ((bz % nz) + nz) % nz;
			else ok = ok && bz >= 0 && bz < nz;
			if (ok) {
				uint64_t bk = ((uint64_t)bz * (uint64_t)ny + (uint64_t)by) * (uint64_t)nx + (uint64_t)bx;
				B = hash_lookup(bk, a.hkeys, a.hvals, a.hmask);
				if (B == A) B = -1;
			}
		}
		// The boxes of the two cells' movers decide most pairs of cells: converged clumps are much smaller than
		// a cell, so two clumps farther apart than tau are rejected and two clumps well within tau are united
		// without a mover-pair test.  (ncu on the version that always scanned the pairs: 11.3 ms at 2^24 for
		// 4.7e8 warp instructions, 23 % warps active - a few warps scanning the nA x nB ~ 10^5 pairs of two
		// unlinked dense cells, through a 64-bit division per pair, were the whole kernel.)  The bounds are a
		// lower / upper bound of the min-image distance of any pair, with a 1e-5 margin on tau^2 for the
		// float rounding of the exact test (coordinate differences this small are exact in float32).
		// Every lane tests the box of the neighbour it found - 32 box fetches in flight instead of one after the
		// other - and only the cells that can be linked are taken one by one below.
		bool near = false, boxLinked = false;
		if (B >= 0) {
			const float4 blo = a.cellBox[2 * (size_t)B], bhi = a.cellBox[2 * (size_t)B + 1];
			float gl2 = 0.0f, gh2 = 0.0f;
			const float al[3] = {alo.x, alo.y, alo.z}, ah[3] = {ahi.x, ahi.y, ahi.z};
			const float bl[3] = {blo.x, blo.y, blo.z}, bh[3] = {bhi.x, bhi.y, bhi.z};
#pragma unroll
			for (int d = 0; d < 3; ++d) {
				const float direct = fmaxf(fmaxf(bl[d] - ah[d], al[d] - bh[d]), 0.0f);
				const float span = fmaxf(ah[d], bh[d]) - fminf(al[d], bl[d]);
				const float wrapped = fmaxf(a.L[d] - span, 0.0f); // the other way round the periodic axis
				const float gmin = fminf(direct, wrapped);
				const float sep = fmaxf(ah[d] - bl[d], bh[d] - al[d]);
				gl2 = fmaf(gmin, gmin, gl2);
				gh2 = fmaf(sep, sep, gh2);
			}
			near = gl2 <= a.fTau2 * 1.00001f;     // else: no mover of A is within tau of a mover of B
			boxLinked = gh2 < a.fTau2 * 0.99999f; // every mover of A is within tau of every mover of B
		}
		uint32_t have = __ballot_sync(SK_FULL, near);
		const uint32_t linkedByBox = __ballot_sync(SK_FULL, boxLinked);
		while (have) {
			int src = __ffs(have) - 1;
			have &= have - 1;
			int Bc = __shfl_sync(SK_FULL, B, src);
			// already in the same component?
			uint32_t ra = 0, rb = 1;
			if (lane == 0) {
				ra = uf_find(a.parent, (uint32_t)A);
				rb = uf_find(a.parent, (uint32_t)Bc);
			}
			ra = __shfl_sync(SK_FULL, ra, 0);
			rb = __shfl_sync(SK_FULL, rb, 0);
			if (ra == rb) continue;
			bool linked = (linkedByBox >> src) & 1u;
			if (!linked) {
				const uint32_t bBeg = a.cellStart[Bc], bEnd = a.cellStart[Bc + 1];
				const uint32_t nB = bEnd - bBeg;
				const unsigned long long nPairs = (unsigned long long)nA * nB;
				// four rounds of 32 pairs between two looks at the result: the loads of a round depend on nothing
				// but the pair number, so four of them are in flight at once
				for (unsigned long long pbase = 0; pbase < nPairs && !linked; pbase += FOF_UNROLL * 32) {
					bool hit = false;
#pragma unroll
					for (int u = 0; u < FOF_UNROLL; ++u) {
						const unsigned long long pi = pbase + u * 32 + lane;
						if (pi < nPairs) {
							uint32_t ia, ib;
							if (nPairs <= 0xffffffffull) ia = (uint32_t)pi / nB, ib = (uint32_t)pi - ia * nB; // 32-bit division
							else ia = (uint32_t)(pi / nB), ib = (uint32_t)(pi % nB);
							float4 pa = a.spos[aBeg + ia];
							float4 pb = a.spos[bBeg + ib];
							// kd.c:871-875 with the query shifted by +-L first (INTERCONT)
							float dx = minimg_dx(pa.x, __fadd_rn(pa.x, a.L[0]), __fsub_rn(pa.x, a.L[0]), a.hL[0], pb.x);
							float dy = minimg_dx(pa.y, __fadd_rn(pa.y, a.L[1]), __fsub_rn(pa.y, a.L[1]), a.hL[1], pb.y);
							float dz = minimg_dx(pa.z, __fadd_rn(pa.z, a.L[2]), __fsub_rn(pa.z, a.L[2]), a.hL[2], pb.z);
							hit = hit || dist2_rn(dx, dy, dz) < a.fTau2;
						}
					}
					linked = __any_sync(SK_FULL, hit);
				}
			}
			if (linked && lane == 0) uf_union(a.parent, (uint32_t)A, (uint32_t)Bc);
			__syncwarp();
		}
	}
}

// root per cell, and the smallest member iOrder per component
__global__ void __launch_bounds__(256)
    k_component_min(int m, const uint32_t *moverCell, const float4 *spos, uint32_t *parent, uint32_t *minOrd)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t r = uf_find(parent, moverCell[i]);
	atomicMin(&minOrd[r], (uint32_t)__float_as_int(spos[i].w));
}

__global__ void __launch_bounds__(256) k_mark_reps(int nCells, const uint32_t *parent, const uint32_t *minOrd,
                                                   uint32_t *repFlag)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nCells) return;
	if (parent[c] == (uint32_t)c) repFlag[minOrd[c]] = 1u;
}

__global__ void __launch_bounds__(256)
    k_assign_gid(int m, const uint32_t *moverCell, const float4 *spos, uint32_t *parent, const uint32_t *minOrd,
                 const uint32_t *repScan, int *gid, int *repOrd)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	uint32_t r = uf_find(parent, moverCell[i]);
	uint32_t mo = minOrd[r];
	int g = (int)repScan[mo] + 1;
	int ord = __float_as_int(spos[i].w);
	gid[ord] = g;
	if ((uint32_t)ord == mo) repOrd[g] = ord;
}

void stage_fof(skidgpu_ctx &c, float fTau, int *nGroupOut)
{
	cudaStream_t s = c.stream;
	const int n = c.n, m = c.nMove;
	StageTimer tm(c, 2);
	int *gid = c.gid.alloc(n > 0 ? n : 1);
	CK(cudaMemsetAsync(gid, 0, sizeof(int) * (size_t)n, s));
	c.haveCenters = false;
	if (m == 0) { // kd.c:815-823
		c.nGroup = 1;
		c.repOrd.alloc(1);
		if (nGroupOut) *nGroupOut = 1;
		tm.stop();
		return;
	}
	// ---- grid
	GridSpec g;
	std::vector<float> hb(6);
	const double cs = 0.57 * (double)fTau;
	if (!(cs > 0.0)) throw SkidError("skidgpu_fof: tau must be > 0");
	double cellMax = 0.0;
	if (c.bPeriodic) {
		for (int d = 0; d < 3; ++d) {
			double L = (double)c.L[d];
			double ncd = ceil(L / cs);
			if (ncd < 1) ncd = 1;
			if (ncd > 2097151.0) throw SkidError("skidgpu_fof: tau too small for the FoF grid (> 2^21 cells per axis)");
			g.nc[d] = (int)ncd;
			g.lo[d] = (double)c.C[d] - 0.5 * L;
			g.inv[d] = ncd / L;
			g.wrap[d] = 1;
			if (L / ncd > cellMax) cellMax = L / ncd;
		}
	} else {
		tree_bbox_only(c.treeM, c.mx.p, c.my.p, c.mz.p, m, s);
		CK(cudaMemcpyAsync(hb.data(), c.treeM.bbox.p, 6 * sizeof(float), cudaMemcpyDeviceToHost, s));
		CK(cudaStreamSynchronize(s));
		for (int d = 0; d < 3; ++d) {
			double ext = (double)hb[3 + d] - (double)hb[d];
			double ncd = floor(ext / cs) + 1.0;
			if (ncd > 2097151.0) throw SkidError("skidgpu_fof: tau too small for the FoF grid (> 2^21 cells per axis)");
			g.nc[d] = (int)ncd;
			g.lo[d] = (double)hb[d];
			g.inv[d] = 1.0 / cs;
			g.wrap[d] = 0;
		}
		cellMax = cs;
	}
	// smallest cell edge decides the stencil radius: offsets > R are farther than tau apart
	double cellMin = cellMax;
	if (c.bPeriodic)
		for (int d = 0; d < 3; ++d) {
			double e = (double)c.L[d] / g.nc[d];
			if (e < cellMin) cellMin = e;
		}
	g.R = (int)ceil((double)fTau * (1.0 + 1e-5) / cellMin);
	if (g.R < 1) g.R = 1;
	if (g.R > 4) throw SkidError("skidgpu_fof: internal: stencil radius > 4");
	int bits = 1;
	{
		double tot = (double)g.nc[0] * (double)g.nc[1] * (double)g.nc[2];
		while (bits < 63 && ldexp(1.0, bits) < tot) ++bits;
	}

	// ---- sort movers by cell
	uint64_t *keys = c.treeM.keys.alloc(m);
	uint32_t *idx = c.treeM.perm.alloc(m);
	SK_LAUNCH(k_cell_keys, (unsigned)ceil_div(m, 256), 256, 0, s, m, c.mx.p, c.my.p, c.mz.p, g, keys, idx);
	radix_sort_pairs(keys, idx, m, bits, c.ws, s);
	uint32_t *flags = c.flags.alloc((size_t)(m > n ? m : n) + 1);
	uint32_t *scan = c.scan.alloc((size_t)(m > n ? m : n) + 64);
	SK_LAUNCH(k_cell_heads, (unsigned)ceil_div(m, 256), 256, 0, s, m, keys, flags);
	exclusive_scan_u32(flags, scan, m, c.ws, s);
	uint32_t nCellsU = 0;
	CK(cudaMemcpyAsync(&nCellsU, scan + m, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	const int nCells = (int)nCellsU;

	FofScratch &S = c.fofS; // persistent scratch (ctx.cuh)
	auto &cellStart = S.cellStart, &moverCell = S.moverCell, &hvals = S.hvals, &parent = S.parent, &minOrd = S.minOrd;
	auto &cellKey = S.cellKey, &hkeys = S.hkeys;
	auto &spos = S.spos, &cellBox = S.cellBox;
	cellStart.alloc(nCells + 1);
	cellKey.alloc(nCells);
	moverCell.alloc(m);
	spos.alloc(m);
	SK_LAUNCH(k_cell_fill, (unsigned)ceil_div(m, 256), 256, 0, s, m, keys, flags, scan, cellStart.p, cellKey.p,
	          moverCell.p, nCells);
	SK_LAUNCH(k_gather_sorted_movers, (unsigned)ceil_div(m, 256), 256, 0, s, m, idx, c.mx.p, c.my.p, c.mz.p, c.mOrd.p,
	          spos.p);
	cellBox.alloc(2 * (size_t)nCells);
	SK_LAUNCH(k_cell_boxes, (unsigned)ceil_div((size_t)nCells * 32, 256), 256, 0, s, nCells, cellStart.p, spos.p, cellBox.p);
	uint32_t hcap = 64;
	while (hcap < 2u * (uint32_t)nCells) hcap <<= 1;
	hkeys.alloc(hcap);
	hvals.alloc(hcap);
	CK(cudaMemsetAsync(hkeys.p, 0xff, sizeof(uint64_t) * hcap, s));
	SK_LAUNCH(k_hash_insert, (unsigned)ceil_div(nCells, 256), 256, 0, s, nCells, cellKey.p, hkeys.p, hvals.p, hcap - 1);
	parent.alloc(nCells);
	SK_LAUNCH(k_uf_init, (unsigned)ceil_div(nCells, 256), 256, 0, s, nCells, parent.p);

	LinkArgs la;
	la.nCells = nCells;
	la.cellStart = cellStart.p;
	la.cellKey = cellKey.p;
	la.spos = spos.p;
	la.cellBox = cellBox.p;
	la.hkeys = hkeys.p;
	la.hvals = hvals.p;
	la.hmask = hcap - 1;
	la.parent = parent.p;
	la.g = g;
	for (int d = 0; d < 3; ++d) {
		la.L[d] = c.L[d];
		la.hL[d] = 0.5f * c.L[d];
	}
	la.fTau2 = fTau * fTau; // kd.c:830
	SK_LAUNCH(k_link_cells, (unsigned)ceil_div((size_t)nCells * 32, 256), 256, 0, s, la);

	// ---- canonical labels
	minOrd.alloc(nCells);
	CK(cudaMemsetAsync(minOrd.p, 0xff, sizeof(uint32_t) * nCells, s));
	SK_LAUNCH(k_component_min, (unsigned)ceil_div(m, 256), 256, 0, s, m, moverCell.p, spos.p, parent.p, minOrd.p);
	CK(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (size_t)n, s));
	SK_LAUNCH(k_mark_reps, (unsigned)ceil_div(nCells, 256), 256, 0, s, nCells, parent.p, minOrd.p, flags);
	exclusive_scan_u32(flags, scan, n, c.ws, s);
	uint32_t nG = 0;
	CK(cudaMemcpyAsync(&nG, scan + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	c.nGroup = (int)nG + 1;
	int *repOrd = c.repOrd.alloc(c.nGroup);
	SK_LAUNCH(k_assign_gid, (unsigned)ceil_div(m, 256), 256, 0, s, m, moverCell.p, spos.p, parent.p, minOrd.p, scan,
	          gid, repOrd);
	if (nGroupOut) *nGroupOut = c.nGroup;
	tm.stop();
}
