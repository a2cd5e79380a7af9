// Stage 1: spatial index build.  Replaces kdBuildTree (kd.c:371-460; kdSelectInit 225-253,
// UpPassInit 307-336, Combine 287-304).  The reference builds a balanced median-split binary
// kd-tree with 16-particle buckets; tree shape does not affect any result (only tie order), so
// the GPU index is designed for warp-wide traversal instead: points are sorted by a 48-bit Hilbert
// key, cut into buckets of 32 consecutive points (one coalesced 512 B float4 load per bucket),
// and every internal node has 32 children so a warp tests all child boxes of a node at once.
#include "ctx.cuh"

__global__ void k_bbox_init(float *bbox)
{
	if (threadIdx.x < 3) ((unsigned int *)bbox)[threadIdx.x] = 0xffffffffu;      // flipped +max
	else if (threadIdx.x < 6) ((unsigned int *)bbox)[threadIdx.x] = 0u;         // flipped -max
}

__global__ void __launch_bounds__(256) k_bbox(const float *x, const float *y, const float *z, int n,
                                              float *bbox)
{
	float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float p[3] = {x[i], y[i], z[i]};
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			lo[d] = fminf(lo[d], p[d]);
			hi[d] = fmaxf(hi[d], p[d]);
		}
	}
#pragma unroll
	for (int d = 0; d < 3; ++d) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			lo[d] = fminf(lo[d], __shfl_xor_sync(SK_FULL, lo[d], o));
			hi[d] = fmaxf(hi[d], __shfl_xor_sync(SK_FULL, hi[d], o));
		}
	}
	if ((threadIdx.x & 31) == 0) {
		unsigned int *b = (unsigned int *)bbox;
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			atomicMin(&b[d], float_flip(lo[d]));
			atomicMax(&b[3 + d], float_flip(hi[d]));
		}
	}
}

__global__ void k_bbox_finish(float *bbox)
{
	if (threadIdx.x < 6) bbox[threadIdx.x] = float_unflip(((unsigned int *)bbox)[threadIdx.x]);
}

// Sort key (TREE_KEY_BITS = 48 bits): Hilbert index of the point's cell, 16 bits per axis (cells of 1/65536 of the
// box, finer than the particle spacing in any realistic core; equal keys keep their input order).
__global__ void __launch_bounds__(256) k_sfc_keys(const float *x, const float *y, const float *z, int n,
                                                  const float *bbox, uint64_t *keys, uint32_t *perm)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float p[3] = {x[i], y[i], z[i]};
	uint32_t q[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) {
		double ext = (double)bbox[3 + d] - (double)bbox[d];
		double t = ext > 0.0 ? ((double)p[d] - (double)bbox[d]) / ext : 0.0;
		long long v = (long long)(t * 65536.0);
		if (v < 0) v = 0;
		if (v > 65535) v = 65535;
		q[d] = (uint32_t)v;
	}
	keys[i] = hilbert3(q[0], q[1], q[2], 16);
	perm[i] = (uint32_t)i;
}

// bounding box of n points into t.bbox (device: lo[3], hi[3])
void tree_bbox_only(BoxTree &t, const float *x, const float *y, const float *z, int n, cudaStream_t s)
{
	float *bbox = t.bbox.alloc(8);
	SK_LAUNCH(k_bbox_init, 1, 32, 0, s, bbox);
	int nb = (int)ceil_div(n > 0 ? n : 1, 256);
	if (nb > 1184) nb = 1184;
	SK_LAUNCH(k_bbox, nb, 256, 0, s, x, y, z, n, bbox);
	SK_LAUNCH(k_bbox_finish, 1, 32, 0, s, bbox);
}

// ------------------------------------------------------------------ distributed sort
// Every rank holds the same n (key, val) pairs (the snapshot is replicated), so the splitters need
// no communication: a regular sample of the keys is sorted by everybody, rank r keeps the pairs whose key
// lies in [split[r], split[r+1]) - in input order, so the stable local sort leaves equal keys in input
// order exactly like one global stable sort - sorts them and the pieces are all-gathered in rank order.
constexpr int DS_SAMPLES = 1 << 16;

__global__ void __launch_bounds__(256) k_ds_sample(const uint64_t *keys, size_t n, size_t stride, int ns, uint64_t *samp, uint32_t *sv)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= ns) return;
	samp[i] = keys[(size_t)i * stride];
	sv[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256)
    k_ds_flags(const uint64_t *keys, size_t n, const uint64_t *samp, int ns, int rank, int nranks, uint32_t *flags)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t lo = rank == 0 ? 0ull : samp[(size_t)ns * rank / nranks];
	const uint64_t k = keys[i];
	bool in = k >= lo;
	if (rank < nranks - 1) in = in && k < samp[(size_t)ns * (rank + 1) / nranks];
	flags[i] = in ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
    k_ds_compact(const uint64_t *keys, const uint32_t *vals, size_t n, const uint32_t *flags, const uint32_t *scan, uint64_t *ko,
                 uint32_t *vo, int *cnt, int rank)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) cnt[rank] = (int)scan[n];
	if (i >= n || !flags[i]) return;
	ko[scan[i]] = keys[i];
	vo[scan[i]] = vals[i];
}

void dist_sort_pairs(skidgpu_ctx &c, uint64_t *keys, uint32_t *vals, size_t n, int bits)
{
	cudaStream_t s = c.stream;
	Workspace &ws = c.ws;
	if (c.nranks <= 1 || !c.comm || n < (size_t)DS_SAMPLES * 4) { // small inputs: everybody sorts everything
		radix_sort_pairs(keys, vals, n, bits, ws, s);
		return;
	}
	const int ns = DS_SAMPLES;
	uint64_t *samp = ws.dsSamp.alloc(ns);
	uint32_t *sv = ws.dsSampV.alloc(ns);
	SK_LAUNCH(k_ds_sample, (unsigned)ceil_div(ns, 256), 256, 0, s, keys, n, n / ns, ns, samp, sv);
	radix_sort_pairs(samp, sv, ns, bits, ws, s);
	uint32_t *flags = ws.dsFlag.alloc(n), *scan = ws.dsScan.alloc(n + 64);
	SK_LAUNCH(k_ds_flags, (unsigned)ceil_div(n, 256), 256, 0, s, keys, n, samp, ns, c.rank, c.nranks, flags);
	exclusive_scan_u32(flags, scan, n, ws, s);
	// an upper bound of this rank's share without a round trip: the sample puts ~n/nranks in every range;
	// the buffers are sized for twice that and the exact counts are checked below
	const size_t cap = 2 * (n / c.nranks) + 4096;
	uint64_t *ko = ws.dsKey.alloc(cap);
	uint32_t *vo = ws.dsVal.alloc(cap);
	int *cnt = ws.dsCnt.alloc(c.nranks + 1);
	uint32_t mine = 0;
	CK(cudaMemcpyAsync(&mine, scan + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	if (mine > cap) { // pathological key distribution (many equal keys): grow
		ko = ws.dsKey.alloc(mine);
		vo = ws.dsVal.alloc(mine);
	}
	SK_LAUNCH(k_ds_compact, (unsigned)ceil_div(n, 256), 256, 0, s, keys, vals, n, flags, scan, ko, vo, cnt, c.rank);
	radix_sort_pairs(ko, vo, mine, bits, ws, s);
	sk_allgather(c, cnt, 1, SK_I32);
	std::vector<int> hc(c.nranks);
	CK(cudaMemcpyAsync(hc.data(), cnt, sizeof(int) * c.nranks, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	std::vector<long long> counts(c.nranks), offs(c.nranks);
	long long tot = 0;
	for (int r = 0; r < c.nranks; ++r) {
		counts[r] = hc[r];
		offs[r] = tot;
		tot += hc[r];
	}
	if (tot != (long long)n) throw SkidError("dist_sort_pairs: the ranks do not hold the same input (piece sizes do not add up)");
	sk_allgatherv(c, vo, vals, counts.data(), offs.data(), SK_I32);
	// only the order travels: no caller reads the sorted keys (keys[] is left unsorted on several ranks)
}

void tree_sort_points(BoxTree &t, const float *x, const float *y, const float *z, int n, Workspace &ws,
                      cudaStream_t s, skidgpu_ctx *dist)
{
	t.n = n;
	float *bbox = t.bbox.alloc(8);
	uint64_t *keys = t.keys.alloc(n > 0 ? n : 1);
	uint32_t *perm = t.perm.alloc(n > 0 ? n : 1);
	if (n == 0) return;
	tree_bbox_only(t, x, y, z, n, s);
	SK_LAUNCH(k_sfc_keys, (unsigned)ceil_div(n, 256), 256, 0, s, x, y, z, n, bbox, keys, perm);
	if (dist) dist_sort_pairs(*dist, keys, perm, n, TREE_KEY_BITS);
	else radix_sort_pairs(keys, perm, n, TREE_KEY_BITS, ws, s);
}

// Leaf boxes: `leaf` consecutive sorted points per leaf (leaf = 8, 16 or 32 lanes of a warp).
__global__ void __launch_bounds__(256) k_leaf_boxes(const float4 *pos4, const float *infl, const float *aux,
                                                    int n, float4 *box, int nBoxPadded, int leaf)
{
	long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long b = i / leaf;
	if (b >= nBoxPadded) return; // whole groups exit together: nBoxPadded*leaf is a multiple of 32
	float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	float a = -3.0e38f;
	if (i < n) {
		float4 p = pos4[i];
		float h = infl ? infl[i] : 0.0f;
		lo[0] = __fadd_rd(p.x, -h);
		lo[1] = __fadd_rd(p.y, -h);
		lo[2] = __fadd_rd(p.z, -h);
		hi[0] = __fadd_ru(p.x, h);
		hi[1] = __fadd_ru(p.y, h);
		hi[2] = __fadd_ru(p.z, h);
		if (aux) a = aux[i];
	}
	for (int o = leaf >> 1; o > 0; o >>= 1) {
#pragma unroll
		for (int d = 0; d < 3; ++d) {
			lo[d] = fminf(lo[d], __shfl_xor_sync(SK_FULL, lo[d], o));
			hi[d] = fmaxf(hi[d], __shfl_xor_sync(SK_FULL, hi[d], o));
		}
		a = fmaxf(a, __shfl_xor_sync(SK_FULL, a, o));
	}
	if ((i % leaf) == 0) {
		box[2 * b] = make_float4(lo[0], lo[1], lo[2], a);
		box[2 * b + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
	}
}

// Parent boxes: union of `fan` consecutive child boxes (fan = 8 or 32 lanes of a warp).
__global__ void __launch_bounds__(256) k_upper_boxes(const float4 *child, int nChild, float4 *box,
                                                     int nBoxPadded, int fan)
{
	long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	long long b = c / fan;
	if (b >= nBoxPadded) return;
	float4 lo = make_float4(3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f);
	float4 hi = make_float4(-3.0e38f, -3.0e38f, -3.0e38f, 0.0f);
	if (c < nChild) {
		lo = child[2 * c];
		hi = child[2 * c + 1];
	}
	for (int o = fan >> 1; o > 0; o >>= 1) {
		lo.x = fminf(lo.x, __shfl_xor_sync(SK_FULL, lo.x, o));
		lo.y = fminf(lo.y, __shfl_xor_sync(SK_FULL, lo.y, o));
		lo.z = fminf(lo.z, __shfl_xor_sync(SK_FULL, lo.z, o));
		lo.w = fmaxf(lo.w, __shfl_xor_sync(SK_FULL, lo.w, o));
		hi.x = fmaxf(hi.x, __shfl_xor_sync(SK_FULL, hi.x, o));
		hi.y = fmaxf(hi.y, __shfl_xor_sync(SK_FULL, hi.y, o));
		hi.z = fmaxf(hi.z, __shfl_xor_sync(SK_FULL, hi.z, o));
	}
	if ((c % fan) == 0) {
		box[2 * b] = lo;
		box[2 * b + 1] = hi;
	}
}

// Levels: level 0 = leaves of `leaf` points; a node of level l+1 has `fan` children of level l; the
// root is implicit (its children are the <= fan boxes of level top-1).  Every level is padded to a
// multiple of max(fan, 32/leaf...) boxes with empty boxes so traversals never read out of bounds.
void tree_build_boxes(BoxTree &t, const float4 *pos4, const float *infl, const float *aux, int n,
                      cudaStream_t s, int leaf, int fan)
{
	if ((leaf != 8 && leaf != 16 && leaf != 32) || (fan != 8 && fan != 32))
		throw SkidError("tree_build_boxes: unsupported leaf/fan");
	t.n = n;
	t.leaf = leaf;
	t.fan = fan;
	int cnt[SK_MAXLEV], pad[SK_MAXLEV];
	int top = 0;
	int c = (int)ceil_div(n > 0 ? n : 1, leaf);
	while (true) {
		if (top >= SK_MAXLEV) throw SkidError("tree_build_boxes: too many levels");
		cnt[top] = c;
		pad[top] = (int)ceil_div(c, 32) * 32; // multiple of 32 (hence of fan) boxes
		++top;
		if (c <= fan) break;
		c = (int)ceil_div(c, fan);
	}
	size_t total = 0;
	for (int l = 0; l < top; ++l) total += 2 * (size_t)pad[l];
	float4 *base = t.store.alloc(total);
	size_t off = 0;
	for (int l = 0; l < SK_MAXLEV; ++l) {
		t.box[l] = nullptr;
		t.cnt[l] = 0;
	}
	for (int l = 0; l < top; ++l) {
		t.box[l] = base + off;
		t.cnt[l] = cnt[l];
		off += 2 * (size_t)pad[l];
	}
	t.top = top;
	SK_LAUNCH(k_leaf_boxes, (unsigned)ceil_div((size_t)pad[0] * leaf, 256), 256, 0, s, pos4, infl, aux, n, t.box[0],
	          pad[0], leaf);
	for (int l = 1; l < top; ++l)
		SK_LAUNCH(k_upper_boxes, (unsigned)ceil_div((size_t)pad[l] * fan, 256), 256, 0, s, t.box[l - 1], cnt[l - 1],
		          t.box[l], pad[l], fan);
}
