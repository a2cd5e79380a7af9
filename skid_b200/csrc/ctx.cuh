// Context of the skidgpu library: device-resident state shared by the stage files.
#pragma once
#include "common.cuh"
#include "../../include/skidgpu.h"

// Device time per kernel family, measured with CUDA events on the context's stream WITHOUT synchronising the
// host: spans are recorded while the work is enqueued and resolved (cudaEventElapsedTime) at the end of the
// stage.  Enabled by skidgpu_set_profile (bench.py); off by default, the product path records nothing.
enum { KF_TILE_STEP = 0, KF_KNN = 1, KF_BUILD = 2, KF_FALLBACK = 3, KF_PRUNE = 4, KF_COUNT = 5 };
struct KernelSpans {
	struct Span {
		int fam;
		size_t a, b;
	};
	std::vector<cudaEvent_t> ev;
	std::vector<Span> spans;
	size_t used = 0;
	bool on = false;
	cudaEvent_t next()
	{
		if (used == ev.size()) {
			cudaEvent_t e;
			CK(cudaEventCreate(&e));
			ev.push_back(e);
		}
		return ev[used++];
	}
	void begin(int fam, cudaStream_t s)
	{
		if (!on) return;
		spans.push_back({fam, used, 0});
		CK(cudaEventRecord(next(), s));
	}
	void end(cudaStream_t s)
	{
		if (!on) return;
		spans.back().b = used;
		CK(cudaEventRecord(next(), s));
	}
	// after the stream has been synchronised: add the spans to ms[fam] / launches[fam]
	void resolve(double *ms, int *cnt)
	{
		for (const Span &sp : spans) {
			float t = 0;
			if (cudaEventElapsedTime(&t, ev[sp.a], ev[sp.b]) == cudaSuccess) {
				ms[sp.fam] += t;
				++cnt[sp.fam];
			}
		}
		spans.clear();
		used = 0;
	}
	~KernelSpans()
	{
		for (cudaEvent_t e : ev) cudaEventDestroy(e);
	}
};

// Scratch of the FoF and unbinding stages.  These used to be locals of the stage functions (allocated from the
// stream-ordered pool and freed on every call); on the massive-halo box the pool then sometimes needed
// 0.1 - 0.7 s to carve the FoF buffers out of what the unbinding stage of the previous pass had left behind
// (measured with events around the section; the kernels themselves are stable).  They live as long as the
// context and only grow.
struct FofScratch {
	DevBuf<uint32_t> cellStart, moverCell, hvals, parent, minOrd;
	DevBuf<uint64_t> cellKey, hkeys;
	DevBuf<float4> spos, cellBox;
};
struct UnbindScratch {
	DevBuf<uint32_t> order, tiles, tileStart, scCnt, scStart, scList, cntU, idx0;
	DevBuf<uint64_t> keys;
	DevBuf<int> gStart, qord, map, clsList, cpack;
	DevBuf<float4> qr, qv, posS;
	DevBuf<float> softS, sx, sy, sz, pack;
	DevBuf<double> pot;
	DevBuf<unsigned int> dCnt;
	DevBuf<skidgpu_pgroup> cat2;
	BoxTree treeS;
};

struct skidgpu_ctx {
	int device = 0;
	cudaStream_t stream = 0;
	std::string err;
	float L[3], C[3];
	int bPeriodic = 0, bDiag = 0;
	int rank = 0, nranks = 1;
	void *comm = nullptr; // ncclComm_t of this context (dist.cu); null on one GPU
	long long commBytes = 0, commCalls = 0;
	skidgpu_reduce_cb reduceCb = nullptr; // test shim: exchanges through a host callback when no communicator is set
	void *reduceUser = nullptr;

	// ---- particles, SoA by iOrder (file order: gas, dark, star; kd.c:113-119)
	int n = 0, nGas = 0, nDark = 0, nStar = 0, inType = 0;
	int reservedFor = 0; // largest particle count the memory pool was pre-grown for (api.cu set_counts)
	DevBuf<float> x, y, z, vx, vy, vz, mass, soft, temp;
	DevBuf<float> rho, ball2; // by iOrder; 0 for non scatter-active
	DevBuf<float> rhoStat;    // rho with the scatterers cut at step 0 zeroed (what kdOutStats reads after -fic)
	bool haveRhoStat = false;
	DevBuf<skidgpu_pinit> aos; // staging for the AoS upload
	Workspace ws;

	// ---- scatter-active set + kNN tree (stage 1/2)
	int nAct = 0, nSmooth = 0, bGasAndDark = 0, bGasOnly = 0;
	DevBuf<uint32_t> flags, scan;
	DevBuf<uint32_t> actIdx;  // compacted file indices of scatter-active particles
	DevBuf<float> ax_, ay_, az_; // gathered positions of the active set (tree input)
	BoxTree treeA;
	DevBuf<float4> posA;   // sorted (x,y,z,mass)
	DevBuf<int> iordA;     // sorted position -> file index
	DevBuf<float> ball2A;  // sorted
	DevBuf<double> rho64A; // sorted, f64 accumulators
	DevBuf<float> rhoA;    // sorted, final f32 density
	bool keepNbr = false;
	DevBuf<int> nbr;
	DevBuf<float> nbrD2;

	// ---- scatterer entities = active originals + periodic replicas (smooth1.c:278-332)
	int nEnt = 0, nExtra = 0;
	DevBuf<float> ex, ey, ez, eInfl, eRhoSorted;
	DevBuf<float4> entPosU;  // unsorted (x,y,z,ball2)
	DevBuf<float4> entNRU;   // unsorted (4/fBall2, fNorm, rho, 0)
	DevBuf<uint32_t> entSrcU; // unsorted: sorted-A index | 0x80000000 for a replica
	DevBuf<float4> entPos;   // sorted (x,y,z,ball2)
	DevBuf<float4> entNR;    // sorted (4/fBall2, fNorm, rhoEff, 0): rhoEff = 0 once cut at step 0
	DevBuf<float4> entRec;   // sorted, interleaved copy: rec[2e] = entPos[e], rec[2e+1] = entNR[e]
	DevBuf<uint32_t> entSrc;
	DevBuf<uint8_t> entTouched;
	BoxTree treeE;

	// ---- movers (kd.c:630-666), Morton order of their initial positions
	int nMove = 0, nActive = 0, bNoPrune = 0;
	long long moverSteps = 0;
	DevBuf<float> mx, my, mz, rox, roy, roz;
	DevBuf<int> mOrd; // mover id -> iOrder
	DevBuf<uint32_t> actList, actList2;
	// tiles: TILE consecutive entries of the position-sorted active list share one scatterer list (move.cu)
	int nTiles = 0, tileStepsLeft = 0, tileWindow = 5, tileBuilds = 0, superCap = 2048;
	DevBuf<uint64_t> tKeys;
	DevBuf<uint32_t> tList, tOff;
	uint32_t bigBase = 0, nBig = 0;
	DevBuf<float4> tPos;
	DevBuf<int> tCnt;
	DevBuf<uint8_t> tPend;
	DevBuf<uint32_t> supList;
	DevBuf<int> supCnt;
	DevBuf<uint32_t> tileQueue, shortQueue;
	bool tileFresh = false;
	int moveKernel = 0;    // test hook (skidgpu_debug_move_kernel): 0 tiles, 1 a tree walk per mover and step
	int actPar = 0;        // which of the two device-side active counts (dT[8], dT[9]) is current
	int nActiveBound = 0;  // host-side upper bound of the device-side active count (grid sizes)
	uint32_t *hLog = nullptr; // pinned host mirror of the per-block log slots
	DevBuf<uint32_t> dLog;
	std::vector<cudaEvent_t> logEv;
	DevBuf<float> tmpx, tmpy, tmpz;
	BoxTree treeM;
	DevBuf<uint32_t> dT; // [0] = T used this step (float bits), [1] = min rho of hit entities this step
	DevBuf<uint32_t> dCount;
	bool keepStep0 = false;
	DevBuf<float> a0x, a0y, a0z;
	DevBuf<uint8_t> aliveByOrd;
	int shardLo = 0, shardHi = 0;
	bool cyclic = false; // movers owned block-cyclically (move.cu) instead of [shardLo, shardHi)
	int nOwned = 0;
	DevBuf<float> mxyz; // packed x|y|z blocks for the multi-GPU exchange of mover positions

	// ---- groups
	int nGroup = 0; // groups + 1 (kd->nGroup)
	DevBuf<int> gid;    // by iOrder
	DevBuf<int> repOrd; // per group: iOrder of its reference member (rel)
	DevBuf<int> gN;
	DevBuf<double> gAcc; // per group accumulators
	DevBuf<skidgpu_pgroup> gCat;
	std::vector<skidgpu_pgroup> hCat;
	bool haveCenters = false;

	FofScratch fofS;
	UnbindScratch unbS;

	// ---- timing / counters
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double stage_ms[6] = {0, 0, 0, 0, 0, 0};
	long long nQueries = 0, nPairs = 0;
	// kernel-family timing (skidgpu_kernel_ms)
	KernelSpans spans;
	double kernel_ms[KF_COUNT] = {0, 0, 0, 0, 0};
	int kernel_launches[KF_COUNT] = {0, 0, 0, 0, 0};
};

// multi-GPU exchange points (dist.cu); no-ops on one rank
void sk_reduce(skidgpu_ctx &c, void *dev, long long count, int dtype, int op);
void sk_allgather(skidgpu_ctx &c, void *buf, long long per, int dtype);
void sk_allgatherv(skidgpu_ctx &c, const void *send, void *recv, const long long *counts, const long long *offs, int dtype);
void dist_unique_id(void *id128);
void dist_comm_init(skidgpu_ctx &c, const void *id128, int rank, int nranks);
void dist_comm_destroy(skidgpu_ctx &c);
// stable sort of (key, val) pairs that every rank holds identically: each rank sorts one key range, the pieces
// are all-gathered (tree.cu).  Only vals[] is sorted on return (keys[] is scratch).  radix_sort_pairs on one rank.
void dist_sort_pairs(skidgpu_ctx &c, uint64_t *keys, uint32_t *vals, size_t n, int bits);
#define SK_I32 0
#define SK_U8 1
#define SK_F32 2
#define SK_F64 3
#define SK_MIN 0
#define SK_MAX 1
#define SK_SUM 2

constexpr int MOVE_OWN_BLOCK = 4096; // movers are owned in blocks of this many consecutive movers (move.cu)

// stage entry points (each in its own .cu)
void stage_density(skidgpu_ctx &c, int nSmooth, int bGasAndDark, int bGasOnly, int *nExtraScat);
void stage_move(skidgpu_ctx &c, float fDensMin, float fTempMax, float fMassMax, float fCvg, float fStep,
                int bForceInitialCut, int bNoPrune, skidgpu_log_cb cb, void *user, int *nMove, int *nIttr);
void stage_microstep(skidgpu_ctx &c, int nSteps, float fStep, skidgpu_log_cb cb, void *user);
void stage_fof(skidgpu_ctx &c, float fTau, int *nGroup);
void stage_centers(skidgpu_ctx &c);
void stage_set_groups(skidgpu_ctx &c, const int *piGroup, int nGroup, const skidgpu_pgroup *centres);
void stage_unbind(skidgpu_ctx &c, float fG, float z, double fCosmo, int iSoftType, float fScoop,
                  int bNoUnbind, int nMaxMembers, int nMinMembers, int *nUnbound, int *nGroupBefore);

void stage_stats(skidgpu_ctx &c, float fG, float z, double dExpHub, float fDensMin, float fTempMax,
                 skidgpu_stat_row *hostRows);

struct StageTimer {
	skidgpu_ctx &c;
	int stage;
	StageTimer(skidgpu_ctx &c_, int st) : c(c_), stage(st) { CK(cudaEventRecord(c.ev0, c.stream)); }
	void stop()
	{
		CK(cudaEventRecord(c.ev1, c.stream));
		CK(cudaEventSynchronize(c.ev1));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
		c.stage_ms[stage] = ms;
	}
};
