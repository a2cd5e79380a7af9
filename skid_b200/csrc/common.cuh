// Shared host/device helpers for the skidgpu library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <vector>

#define SK_FULL 0xffffffffu

struct SkidError : public std::runtime_error {
	explicit SkidError(const std::string &m) : std::runtime_error(m) {}
};

#define CK(call)                                                                              \
	do {                                                                                  \
		cudaError_t e_ = (call);                                                      \
		if (e_ != cudaSuccess) {                                                      \
			char b_[512];                                                         \
			snprintf(b_, sizeof b_, "%s:%d: %s: %s", __FILE__, __LINE__, #call,   \
			         cudaGetErrorString(e_));                                     \
			throw SkidError(b_);                                                  \
		}                                                                             \
	} while (0)

// Launch counter (skidgpu_counter(…,0)): every kernel launch goes through SK_LAUNCH.
extern thread_local long long g_skid_launches; // per host thread: one thread drives one context at a time
#define SK_LAUNCH(kern, grid, block, smem, stream, ...)                                       \
	do {                                                                                  \
		kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                     \
		++g_skid_launches;                                                            \
		CK(cudaGetLastError());                                                       \
	} while (0)

// Grow-only device buffer, stream-ordered: cudaMallocAsync on the calling context's stream (set per
// HOST THREAD by every API entry point, so several contexts can be driven from several threads - one
// thread per GPU in host/skid -gpus N) from the device's default pool, whose release threshold
// skidgpu_create raises so that freed blocks stay cached - steady state does no driver allocation.
// A buffer is freed on the stream it was allocated on.
extern thread_local cudaStream_t g_skid_stream;
template <class T> struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	cudaStream_t owner = 0;
	DevBuf() {}
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	~DevBuf() { release(); }
	void release()
	{
		if (p) cudaFreeAsync(p, owner);
		p = nullptr;
		cap = 0;
	}
	T *alloc(size_t n)
	{
		if (n > cap) {
			release();
			size_t want = n + n / 16 + 64;
			owner = g_skid_stream;
			CK(cudaMallocAsync((void **)&p, want * sizeof(T), owner));
			cap = want;
		}
		return p;
	}
};

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------
// primitives.cu
// ---------------------------------------------------------------------------------------
struct Workspace {
	DevBuf<uint32_t> scanA, scanB, scanC; // block sums for the multi-level scan
	DevBuf<uint32_t> hist;               // radix sort digit histograms
	DevBuf<uint32_t> histScan;
	DevBuf<uint64_t> keyAlt;
	DevBuf<uint32_t> valAlt;
	// distributed sort (tree.cu: dist_sort_pairs)
	DevBuf<uint64_t> dsSamp, dsKey;
	DevBuf<uint32_t> dsSampV, dsVal, dsFlag, dsScan;
	DevBuf<int> dsCnt;
};

// out[0..n] = exclusive prefix sum of in[0..n) ; out has n+1 entries (out[n] = total).
void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, Workspace &ws, cudaStream_t s);
// Stable LSD radix sort of (key,val) pairs on the low `bits` bits of the key.  Result in keys/vals.
void radix_sort_pairs(uint64_t *keys, uint32_t *vals, size_t n, int bits, Workspace &ws, cudaStream_t s);

// ---------------------------------------------------------------------------------------
// tree.cu : 32-wide bucket tree over Morton-ordered points
// ---------------------------------------------------------------------------------------
#define SK_MAXLEV 10
constexpr int TREE_KEY_BITS = 48; // sorted bits of the space-filling-curve keys (tree.cu: k_sfc_keys; move.cu: k_mover_keys)

// Hilbert-curve index of the cell (x, y, z), `bits` bits per axis (Skilling's transposition: undo the excess
// work of the Gray-coded octant path top-down, Gray-encode, interleave).  Consecutive indices are face
// neighbours, so a run of 32 consecutive sorted points is one connected blob: measured on the clustered 2^20
// box, the k=64 ball of a query meets 14.0 leaf boxes of 32 points instead of 22.2 with Morton (Z-order) keys,
// whose runs straddle the jumps of the curve, and 4.7 instead of 8.3 boxes of the level above.
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline uint64_t hilbert3(uint32_t x, uint32_t y, uint32_t z, int bits)
{
	uint32_t X[3] = {x, y, z};
	for (uint32_t Q = 1u << (bits - 1); Q > 1u; Q >>= 1) {
		const uint32_t P = Q - 1u;
		for (int i = 0; i < 3; ++i) {
			if (X[i] & Q) X[0] ^= P;
			else {
				const uint32_t t = (X[0] ^ X[i]) & P;
				X[0] ^= t;
				X[i] ^= t;
			}
		}
	}
	X[1] ^= X[0];
	X[2] ^= X[1];
	uint32_t t = 0;
	for (uint32_t Q = 1u << (bits - 1); Q > 1u; Q >>= 1)
		if (X[2] & Q) t ^= Q - 1u;
	uint64_t key = 0;
	for (int i = 0; i < 3; ++i) {
		uint64_t v = (X[i] ^ t) & 0x1fffffu; // spread the bits three apart
		v = (v | (v << 32)) & 0x1f00000000ffffull;
		v = (v | (v << 16)) & 0x1f0000ff0000ffull;
		v = (v | (v << 8)) & 0x100f00f00f00f00full;
		v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
		v = (v | (v << 2)) & 0x1249249249249249ull;
		key |= v << (2 - i);
	}
	return key;
}
struct BoxTree {
	int n = 0;          // points
	int leaf = 32, fan = 32; // points per leaf box, children per node
	int top = 0;        // number of box levels; level 0 = buckets of 32 points
	int cnt[SK_MAXLEV]; // boxes per level
	float4 *box[SK_MAXLEV]; // box[l][2*j] = (lo.xyz, aux), box[l][2*j+1] = (hi.xyz, 0)
	DevBuf<float4> store;   // backing store of all levels
	DevBuf<uint64_t> keys;
	DevBuf<uint32_t> perm; // sorted position -> input index
	DevBuf<float> bbox;    // 6 floats lo[3], hi[3]
};

// Sort n points (x,y,z arrays) by their 48-bit Hilbert key (tree.cu); fills t.perm.  The bounding box is reduced
// on the device.
// dist (nullable): a context with nranks > 1 whose ranks all hold the same points - the sort is then shared
// between the ranks (every rank sorts one key range, the pieces are all-gathered); same result.
struct skidgpu_ctx;
void tree_sort_points(BoxTree &t, const float *x, const float *y, const float *z, int n,
                      Workspace &ws, cudaStream_t s, skidgpu_ctx *dist = nullptr);
void tree_bbox_only(BoxTree &t, const float *x, const float *y, const float *z, int n, cudaStream_t s);
// Build the box levels over sorted points pos4[0..n) (xyz used).  infl (nullable): per-point
// inflation radius (sorted order); aux (nullable): per-point value whose max goes to lo.w.
void tree_build_boxes(BoxTree &t, const float4 *pos4, const float *infl, const float *aux, int n,
                      cudaStream_t s, int leaf = 32, int fan = 32);

// Device-side view passed by value to kernels.
struct TreeView {
	const float4 *box[SK_MAXLEV];
	int cnt[SK_MAXLEV];
	int top;
	int n;
	int leaf, fan;
};
static inline TreeView tree_view(const BoxTree &t)
{
	TreeView v;
	for (int l = 0; l < SK_MAXLEV; ++l) {
		v.box[l] = t.box[l];
		v.cnt[l] = t.cnt[l];
	}
	v.top = t.top;
	v.n = t.n;
	v.leaf = t.leaf;
	v.fan = t.fan;
	return v;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// (dx*dx + dy*dy) + dz*dz in IEEE float32, round-to-nearest, never contracted into FMAs:
// the reference's arithmetic (smooth1.c:75,107,368; gcc x86-64 SSE2, no FMA).
__device__ __forceinline__ float dist2_rn(float dx, float dy, float dz)
{
	return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Min-image query coordinate: the reference shifts the QUERY by +-L first (sx = x +- lx,
// kd.h:136,148) and then subtracts the particle (smooth1.c:104).  x0/xp/xm = x, x+L, x-L in f32.
__device__ __forceinline__ float minimg_dx(float x0, float xp, float xm, float hL, float px)
{
	float d = __fsub_rn(x0, px);
	if (d > hL) d = __fsub_rn(xm, px);
	else if (d < -hL) d = __fsub_rn(xp, px);
	return d;
}

// Distance (>=0) from coordinate s to interval [lo,hi].
__device__ __forceinline__ float axis_gap(float s, float lo, float hi)
{
	return fmaxf(fmaxf(__fsub_rn(lo, s), __fsub_rn(s, hi)), 0.0f);
}
// Periodic gap: min over the three images of the query.  Monotone w.r.t. minimg_dx above, so a
// box is never pruned while one of its points is inside the ball (same rounding, same order).
__device__ __forceinline__ float axis_gap_periodic(float x0, float xp, float xm, float lo, float hi)
{
	return fminf(axis_gap(x0, lo, hi), fminf(axis_gap(xp, lo, hi), axis_gap(xm, lo, hi)));
}

__device__ __forceinline__ unsigned int float_flip(float f)
{ // order-preserving map float -> uint
	unsigned int u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unflip(unsigned int u)
{
	return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
#endif
