// Context of the skidgpu library: device-resident state shared by the stage files.
#pragma once
#include "common.cuh"
#include "../../include/skidgpu.h"

struct skidgpu_ctx {
	int device = 0;
	cudaStream_t stream = 0;
	std::string err;
	float L[3], C[3];
	int bPeriodic = 0, bDiag = 0;
	int rank = 0, nranks = 1;
	skidgpu_reduce_cb reduceCb = nullptr;
	void *reduceUser = nullptr;

	// ---- particles, SoA by iOrder (file order: gas, dark, star; kd.c:113-119)
	int n = 0, nGas = 0, nDark = 0, nStar = 0, inType = 0;
	int reservedFor = 0; // largest particle count the memory pool was pre-grown for (api.cu set_counts)
	DevBuf<float> x, y, z, vx, vy, vz, mass, soft, temp;
	DevBuf<float> rho, ball2; // by iOrder; 0 for non scatter-active
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
	DevBuf<uint32_t> mList;            // candidate lists, LIST_CAP per mover (move.cu)
	DevBuf<float> lx0, ly0, lz0, ldelta, lhmin;
	DevBuf<int> lcnt;
	float listInitFactor = 0.3f;
	DevBuf<uint32_t> mQueue; // movers that refresh their list this step
	// tiles: TILE consecutive entries of the position-sorted active list share one scatterer list (move.cu)
	int nTiles = 0, tileStepsLeft = 0, tileWindow = 5, tileBuilds = 0, superCap = 2048;
	DevBuf<uint64_t> tKeys;
	DevBuf<uint32_t> tList, tOff;
	uint32_t bigBase = 0, nBig = 0;
	DevBuf<float4> tPos;
	DevBuf<int> tCnt;
	DevBuf<uint8_t> tPend;
	int tileOverlap = 0;
	cudaStream_t stream2 = 0; // k_tile_walk beside k_tile_step at a rebuild step (move.cu)
	cudaEvent_t evWalk0 = nullptr, evWalk1 = nullptr;
	DevBuf<uint32_t> supList;
	DevBuf<int> supCnt;
	DevBuf<uint32_t> tileQueue, shortQueue;
	bool tileFresh = false;
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
	DevBuf<float> mxyz; // contiguous x|y|z copy for the multi-GPU exchange

	// ---- groups
	int nGroup = 0; // groups + 1 (kd->nGroup)
	DevBuf<int> gid;    // by iOrder
	DevBuf<int> repOrd; // per group: iOrder of its reference member (rel)
	DevBuf<int> gN;
	DevBuf<double> gAcc; // per group accumulators
	DevBuf<skidgpu_pgroup> gCat;
	std::vector<skidgpu_pgroup> hCat;
	bool haveCenters = false;

	// ---- timing / counters
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double stage_ms[6] = {0, 0, 0, 0, 0, 0};
	long long nQueries = 0, nPairs = 0;
	// dominant-kernel timing (skidgpu_kernel_ms)
	cudaEvent_t evk0 = nullptr, evk1 = nullptr;
	double kernel_ms[2] = {0, 0};
	int kernel_launches[2] = {0, 0};
};

// Brackets a group of launches of one dominant kernel with events; stop() must be called after a
// stream synchronisation point has been reached (it synchronises on the stop event itself).
struct KernelTimer {
	skidgpu_ctx &c;
	int which;
	bool open = false;
	KernelTimer(skidgpu_ctx &c_, int w) : c(c_), which(w) {}
	void start()
	{
		CK(cudaEventRecord(c.evk0, c.stream));
		open = true;
	}
	void stop(int launches)
	{
		if (!open) return;
		CK(cudaEventRecord(c.evk1, c.stream));
		CK(cudaEventSynchronize(c.evk1));
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, c.evk0, c.evk1));
		c.kernel_ms[which] += ms;
		c.kernel_launches[which] += launches;
		open = false;
	}
};

// multi-GPU agreement point (no-op on one rank)
static inline void sk_reduce(skidgpu_ctx &c, void *dev, long long count, int dtype, int op)
{
	if (c.nranks <= 1) return;
	if (!c.reduceCb) throw SkidError("nranks > 1 but no reduce callback set (skidgpu_set_reduce_cb)");
	if (c.reduceCb(c.reduceUser, dev, count, dtype, op) != 0) throw SkidError("reduce callback failed");
}
#define SK_I32 0
#define SK_U8 1
#define SK_F32 2
#define SK_F64 3
#define SK_MIN 0
#define SK_MAX 1
#define SK_SUM 2

// stage entry points (each in its own .cu)
void stage_density(skidgpu_ctx &c, int nSmooth, int bGasAndDark, int bGasOnly, int *nExtraScat);
void stage_move(skidgpu_ctx &c, float fDensMin, float fTempMax, float fMassMax, float fCvg, float fStep,
                int bForceInitialCut, int bNoPrune, skidgpu_log_cb cb, void *user, int *nMove, int *nIttr);
void stage_microstep(skidgpu_ctx &c, int nSteps, float fStep, skidgpu_log_cb cb, void *user);
void move_mask_unowned(skidgpu_ctx &c); // zero the positions of movers other ranks own (before the sum-exchange)
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
