// Device-wide exclusive scan and stable LSD radix sort (hand-written; no CUB/thrust).
#include "common.cuh"

thread_local long long g_skid_launches = 0;

// ------------------------------------------------------------------ scan
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total)
{
	__shared__ uint32_t wsum[SC_THREADS / 32];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(SK_FULL, inc, o);
		if (lane >= o) inc += t;
	}
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	if (w == 0) {
		uint32_t s = (lane < SC_THREADS / 32) ? wsum[lane] : 0;
#pragma unroll
		for (int o = 1; o < SC_THREADS / 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(SK_FULL, s, o);
			if (lane >= o) s += t;
		}
		if (lane < SC_THREADS / 32) wsum[lane] = s;
	}
	__syncthreads();
	uint32_t base = (w > 0) ? wsum[w - 1] : 0;
	*total = wsum[SC_THREADS / 32 - 1];
	__syncthreads();
	return base + inc - v;
}

// Per-tile local exclusive scan; tile totals to sums[].
__global__ void __launch_bounds__(SC_THREADS) k_scan_tiles(const uint32_t *in, uint32_t *out, size_t n,
                                                           uint32_t *sums)
{
	size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;
	uint32_t v[SC_ITEMS], tsum = 0;
#pragma unroll
	for (int i = 0; i < SC_ITEMS; ++i) {
		v[i] = (base + i < n) ? in[base + i] : 0;
		tsum += v[i];
	}
	uint32_t total;
	uint32_t ex = block_exclusive_scan(tsum, &total);
#pragma unroll
	for (int i = 0; i < SC_ITEMS; ++i) {
		if (base + i < n) out[base + i] = ex;
		ex += v[i];
	}
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_add(uint32_t *out, size_t n, const uint32_t *offs)
{
	size_t base = (size_t)blockIdx.x * SC_TILE + (size_t)threadIdx.x * SC_ITEMS;
	uint32_t o = offs[blockIdx.x];
#pragma unroll
	for (int i = 0; i < SC_ITEMS; ++i)
		if (base + i < n) out[base + i] += o;
}

__global__ void k_scan_total(const uint32_t *in, uint32_t *out, size_t n)
{
	out[n] = (n > 0) ? out[n - 1] + in[n - 1] : 0;
}

static void scan_rec(const uint32_t *in, uint32_t *out, size_t n, DevBuf<uint32_t> **bufs, int depth,
                     cudaStream_t s)
{
	size_t nb = ceil_div(n, SC_TILE);
	if (depth >= 3) throw SkidError("exclusive_scan_u32: input too large");
	uint32_t *sums = bufs[depth]->alloc(2 * nb + 2);
	SK_LAUNCH(k_scan_tiles, (unsigned)nb, SC_THREADS, 0, s, in, out, n, sums);
	if (nb > 1) {
		uint32_t *offs = sums + nb; // nb+1 entries
		scan_rec(sums, offs, nb, bufs, depth + 1, s);
		SK_LAUNCH(k_scan_add, (unsigned)nb, SC_THREADS, 0, s, out, n, offs);
	}
}

// Small inputs (the digit histograms of small sorts, the prune flags late in the move loop, the demo): one block
// walks the whole array, tile by tile, carrying the running total - one launch instead of six.
constexpr size_t SC_SMALL = 8 * SC_TILE; // 16384 elements (~12 us in one block; beyond that the multi-level scan wins)
__global__ void __launch_bounds__(SC_THREADS) k_scan_small(const uint32_t *in, uint32_t *out, size_t n)
{
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (size_t t0 = 0; t0 < n; t0 += SC_TILE) {
		const size_t base = t0 + (size_t)threadIdx.x * SC_ITEMS;
		uint32_t v[SC_ITEMS], tsum = 0;
#pragma unroll
		for (int i = 0; i < SC_ITEMS; ++i) {
			v[i] = (base + i < n) ? in[base + i] : 0;
			tsum += v[i];
		}
		uint32_t total;
		uint32_t ex = block_exclusive_scan(tsum, &total) + carry;
#pragma unroll
		for (int i = 0; i < SC_ITEMS; ++i) {
			if (base + i < n) out[base + i] = ex;
			ex += v[i];
		}
		__syncthreads();
		if (threadIdx.x == 0) carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) out[n] = carry;
}

void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, Workspace &ws, cudaStream_t s)
{
	if (n == 0) {
		CK(cudaMemsetAsync(out, 0, sizeof(uint32_t), s));
		return;
	}
	if (n <= SC_SMALL) {
		SK_LAUNCH(k_scan_small, 1, SC_THREADS, 0, s, in, out, n);
		return;
	}
	DevBuf<uint32_t> *bufs[3] = {&ws.scanA, &ws.scanB, &ws.scanC};
	scan_rec(in, out, n, bufs, 0, s);
	SK_LAUNCH(k_scan_total, 1, 1, 0, s, in, out, n);
}

// ------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_BINS = 256;

__global__ void __launch_bounds__(RS_THREADS) k_rs_count(const uint64_t *keys, size_t n, int shift,
                                                         size_t chunk, uint32_t *hist, int nb)
{
	__shared__ uint32_t h[RS_BINS];
	h[threadIdx.x] = 0;
	__syncthreads();
	size_t beg = (size_t)blockIdx.x * chunk;
	size_t end = beg + chunk < n ? beg + chunk : n;
	for (size_t i = beg + threadIdx.x; i < end; i += RS_THREADS)
		atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
	__syncthreads();
	hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
    k_rs_scatter(const uint64_t *kin, const uint32_t *vin, uint64_t *kout, uint32_t *vout, size_t n,
                 int shift, size_t chunk, const uint32_t *prefix, int nb)
{
	__shared__ uint32_t base[RS_BINS];
	__shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const uint32_t lt = (1u << lane) - 1u;
	base[tid] = prefix[(size_t)tid * nb + blockIdx.x];
	size_t beg = (size_t)blockIdx.x * chunk;
	size_t end = beg + chunk < n ? beg + chunk : n;
	for (size_t tile = beg; tile < end; tile += RS_TILE) {
#pragma unroll
		for (int i = 0; i < RS_WARPS; ++i) wcnt[i][tid] = 0;
		__syncthreads();
		uint64_t k[RS_ITEMS];
		uint32_t v[RS_ITEMS], rank[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
		for (int r = 0; r < RS_ITEMS; ++r) {
			size_t idx = tile + (size_t)w * (32 * RS_ITEMS) + r * 32 + lane;
			bool valid = idx < end;
			k[r] = valid ? kin[idx] : 0;
			v[r] = valid ? vin[idx] : 0;
			uint32_t d = valid ? ((unsigned)(k[r] >> shift) & 255u) : (256u + lane);
			dig[r] = d;
			uint32_t peers = __match_any_sync(SK_FULL, d);
			uint32_t before = __popc(peers & lt);
			uint32_t old = 0;
			if (valid) old = wcnt[w][d];
			__syncwarp();
			if (valid && before == 0) wcnt[w][d] = old + __popc(peers);
			__syncwarp();
			rank[r] = old + before;
		}
		__syncthreads();
		// exclusive prefix over warps for digit `tid`
		uint32_t run = 0;
#pragma unroll
		for (int i = 0; i < RS_WARPS; ++i) {
			uint32_t t = wcnt[i][tid];
			wcnt[i][tid] = run;
			run += t;
		}
		__syncthreads();
#pragma unroll
		for (int r = 0; r < RS_ITEMS; ++r) {
			if (dig[r] < 256u) {
				uint32_t pos = base[dig[r]] + wcnt[w][dig[r]] + rank[r];
				kout[pos] = k[r];
				vout[pos] = v[r];
			}
		}
		__syncthreads();
		base[tid] += run;
		__syncthreads();
	}
}

void radix_sort_pairs(uint64_t *keys, uint32_t *vals, size_t n, int bits, Workspace &ws, cudaStream_t s)
{
	if (n <= 1) return;
	if (n >= (1ull << 32)) throw SkidError("radix_sort_pairs: n too large");
	int passes = (bits + 7) / 8;
	if (passes < 1) passes = 1;
	size_t tiles = ceil_div(n, RS_TILE);
	int nb = (int)(tiles < 1184 ? tiles : 1184); // 148 SMs x 8 resident blocks
	size_t chunk = ceil_div(tiles, nb) * RS_TILE;
	nb = (int)ceil_div(n, chunk);
	uint32_t *hist = ws.hist.alloc((size_t)RS_BINS * nb);
	uint32_t *pref = ws.histScan.alloc((size_t)RS_BINS * nb + 1);
	uint64_t *kalt = ws.keyAlt.alloc(n);
	uint32_t *valt = ws.valAlt.alloc(n);
	uint64_t *kin = keys, *kout = kalt;
	uint32_t *vin = vals, *vout = valt;
	for (int p = 0; p < passes; ++p) {
		SK_LAUNCH(k_rs_count, nb, RS_THREADS, 0, s, kin, n, p * 8, chunk, hist, nb);
		exclusive_scan_u32(hist, pref, (size_t)RS_BINS * nb, ws, s);
		SK_LAUNCH(k_rs_scatter, nb, RS_THREADS, 0, s, kin, vin, kout, vout, n, p * 8, chunk, pref, nb);
		uint64_t *tk = kin;
		kin = kout;
		kout = tk;
		uint32_t *tv = vin;
		vin = vout;
		vout = tv;
	}
	if (kin != keys) {
		CK(cudaMemcpyAsync(keys, kin, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
		CK(cudaMemcpyAsync(vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
	}
}
