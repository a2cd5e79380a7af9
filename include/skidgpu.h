/*
 * skidgpu.h - C-ABI of the B200-native SKID group-finding hot path.
 *
 * SKID (reference: /root/reference, v1.4.1) has no plugin/FFI interface; its only seam is
 * the stage API that main.c calls in a fixed order (prototypes kd.h:338-365,
 * smooth1.h:94-99, grav.h:34-36; call sites main.c:347-495).  Every entry point below
 * replaces one group of those calls.  A flag-compatible C driver (host/skid_main.c) and
 * the reference's own main.c can both bind to it; see INTEGRATION.md.
 *
 * Conventions: opaque context, int status (0 = ok, !=0 = error, text via
 * skidgpu_last_error), caller-owned HOST buffers unless a name says "dev", no C++
 * or torch types.  All arrays indexed "by iOrder" use the particle's position in the
 * input file (PINIT.iOrder, kd.c:165).  There is NO CPU fallback: every compute entry
 * point fails if no CUDA device is usable.
 */
#ifndef SKIDGPU_H
#define SKIDGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define SKIDGPU_OK 0
#define SKIDGPU_ERR 1

/* particle species bits, kd.h:17-19 */
#define SKIDGPU_DARK 1
#define SKIDGPU_GAS 2
#define SKIDGPU_STAR 4
/* softening types, kd.h:24-25 */
#define SKIDGPU_PLUMMER 1
#define SKIDGPU_SPLINE 2

/* Mirrors PINIT (kd.h:27-36), 48 bytes. */
typedef struct skidgpu_pinit {
	float r[3];
	float v[3];
	float fMass;
	float fSoft;
	float fTemp;
	float fBall2;
	float fDensity;
	int iOrder;
} skidgpu_pinit;

/* Mirrors PGROUP (kd.h:45-55), 68 bytes. */
typedef struct skidgpu_pgroup {
	float rel[3];
	float rCenter[3];
	float rBound[3];
	float vcm[3];
	float fMass;
	float fRadius;
	int nMembers;
	int pStart;
	int pCurr;
} skidgpu_pgroup;

typedef struct skidgpu_ctx skidgpu_ctx;

/* Progress callback: replaces the "Ittr:%d nActive:%d nScatter:%d" / "Microstep:%d
 * nScatter:%d" prints of main.c:402,416,436.  kind 0 = move block, 1 = micro step. */
typedef void (*skidgpu_log_cb)(void *user, int kind, int iter, int nActive, int nScatter);

/* kdInit (kd.c:46, main.c:347) / kdFinish (kd.c:1842, main.c:495).
 * device: CUDA ordinal.  fPeriod = FLT_MAX per axis when not periodic (main.c:125-128). */
int skidgpu_create(skidgpu_ctx **pctx, int device, const float fPeriod[3],
                   const float fCenter[3], int bPeriodic, int bDiag);
void skidgpu_destroy(skidgpu_ctx *ctx);
const char *skidgpu_last_error(skidgpu_ctx *ctx);

/* Optional hint (no reference equivalent): grow the context's device memory pool by `bytes` ahead of
 * time, e.g. from a helper thread while the snapshot is still being read (the particle count is in the
 * TIPSY header).  The hot path needs about 700 bytes per particle; a first pass otherwise pays for the
 * pool growth on its critical path.  Clamped to half of the free device memory; never changes results. */
int skidgpu_reserve(skidgpu_ctx *ctx, unsigned long long bytes);

/* Multi-GPU (no reference equivalent - the reference is serial).  One context per GPU, each driven
 * by its own host thread (host/skid -gpus N) or its own process (bench.py under torchrun); EVERY context gets
 * the same snapshot (skidgpu_set_particles) and makes the same stage calls.  Particles, trees and scatterers
 * are replicated; the sorts behind the tree builds, the kNN queries, the movers and the groups to unbind are
 * shared between the ranks.  The library issues the few exchanges itself with NCCL over NVLink, on the context's
 * stream: all-gather of fBall2 and all-reduce of the f64 density partials, max of the step-0 "touched" flags,
 * min of fScatDens every step, sum of the active-mover / scatterer counts every 5 steps, all-gather of the
 * converged mover positions before grouping, min of labels + sum of the changed catalogue rows after unbinding.
 *   skidgpu_comm_unique_id   rank 0 makes the 128-byte NCCL id and hands it to the others (any transport)
 *   skidgpu_comm_init        collective over all nranks contexts (ncclCommInitRank); implies skidgpu_set_shard
 *   skidgpu_comm_bytes       payload bytes / number of collectives issued by this context since create
 * NCCL (libnccl.so.2) is loaded at the first of these calls; a single-GPU run never needs it. */
#define SKIDGPU_UNIQUE_ID_BYTES 128
int skidgpu_comm_unique_id(void *id128);
int skidgpu_comm_init(skidgpu_ctx *ctx, const void *id128, int rank, int nranks);
long long skidgpu_comm_bytes(skidgpu_ctx *ctx, long long *nCalls);
/* rank/nranks without a communicator: exchanges then go through the callback below (test shim). */
int skidgpu_set_shard(skidgpu_ctx *ctx, int rank, int nranks);

/* What kdReadTipsy (kd.c:122-222, main.c:348) leaves in kd->pInit: n = nGas+nDark+nStar
 * particles in file order (gas, dark, star), p[i].iOrder == i.  fTime = header time.
 * With a communicator (skidgpu_comm_init, nranks > 1) the call is collective: every rank passes the same
 * snapshot, uploads only its 1/nranks slice of it and the slices are all-gathered between the GPUs. */
int skidgpu_set_particles(skidgpu_ctx *ctx, const skidgpu_pinit *p, int n, int nGas,
                          int nDark, int nStar);

/* Same, from already device-resident SoA float arrays (r: 3 arrays of n, v: 3 arrays of n).
 * Used by bench.py's device-resident ("value") leg. */
int skidgpu_set_particles_dev(skidgpu_ctx *ctx, const float *dx, const float *dy,
                              const float *dz, const float *dvx, const float *dvy,
                              const float *dvz, const float *dmass, const float *dsoft,
                              const float *dtemp, int n, int nGas, int nDark, int nStar);

/* kdSetSoft (kd.c:103-110, main.c:464): override every softening. */
int skidgpu_set_soft(skidgpu_ctx *ctx, float fEps);

/* Test shim for the exchanges above when no communicator is set (skidgpu_set_shard only): the library calls
 * `cb` instead of NCCL.  The buffer is DEVICE memory of this context, the call is made after all producing work
 * has been enqueued on the context's stream (skidgpu_stream) and the reduced result must be visible to work
 * enqueued on that stream afterwards.  dtype: 0 int32, 1 uint8, 2 float32, 3 float64.  op: 0 min, 1 max, 2 sum.
 * All-gathers are expressed as zero-fill + sum.  Return 0 on success. */
typedef int (*skidgpu_reduce_cb)(void *user, void *dev, long long count, int dtype, int op);
int skidgpu_set_reduce_cb(skidgpu_ctx *ctx, skidgpu_reduce_cb cb, void *user);

/* kdScatterActive + kdBuildTree + smInit + smDensityInit (main.c:374-378):
 * tree over the scatter-active species, exact periodic k-nearest (k = nSmooth, self
 * included), fBall2 = k-th distance^2 (bitwise as the reference), symmetric
 * gather+scatter spline density, then the periodic replica scatterers (smooth1.c:278-332).
 * rho_by_iOrder / ball2_by_iOrder: optional (NULL ok) n floats; density of non
 * scatter-active particles is 0 (kd.c:166), their ball2 is left 0. */
int skidgpu_density(skidgpu_ctx *ctx, int nSmooth, int bGasAndDark, int bGasOnly,
                    float *rho_by_iOrder, float *ball2_by_iOrder, int *nExtraScat);

/* Test hook: neighbour lists of the last skidgpu_density call.  For scatter-active
 * particle with file index i: nbr[i*nSmooth + e] = iOrder of e-th neighbour (ascending
 * by (d2, tree index)), d2[i*nSmooth + e] its squared distance.  Rows of non-active
 * particles are filled with -1.  Must be enabled BEFORE skidgpu_density. */
int skidgpu_keep_neighbors(skidgpu_ctx *ctx, int bKeep);
int skidgpu_get_neighbors(skidgpu_ctx *ctx, int *nbr, float *d2);

/* kdInitMove + the whole "flow" loop (main.c:394-419): select movers (CutCriterion,
 * kd.c:555-597), then step 0 with the initial scatterer cut (bInitial = dark-only input
 * || bForceInitialCut), then blocks of 5 steps + kdPruneInactive until no mover is active.
 * bNoPrune = the "-nsp" extension: never remove scatterers.
 * Outputs: *nMove movers, *nIttr = number of "Ittr" lines printed (blocks + 1).
 * The loop is driven from the device: the active count stays in device memory, the host enqueues blocks a few
 * ahead of the counts it has read back, so `cb` is called a block or two after the block it reports (in order). */
int skidgpu_move(skidgpu_ctx *ctx, float fDensMin, float fTempMax, float fMassMax,
                 float fCvg, float fStep, int bForceInitialCut, int bNoPrune,
                 skidgpu_log_cb cb, void *user, int *nMove, int *nIttr);

/* Test hook: accelerations (density gradient, smAccDensity smooth1.c:408-518) of step 0.
 * Must be enabled BEFORE skidgpu_move.  iOrder[nMove], a[3*nMove]. */
int skidgpu_keep_step0(skidgpu_ctx *ctx, int bKeep);
int skidgpu_get_step0(skidgpu_ctx *ctx, int *iOrder, float *a3, unsigned char *scat_alive_by_iOrder);

/* kdFoF (kd.c:802-917, main.c:425): components of {min-image d2 < tau^2} over all movers
 * at their converged positions.  *nGroup = number of groups + 1 (kd->nGroup). */
int skidgpu_fof(skidgpu_ctx *ctx, float fTau, int *nGroup);

/* kdReactivateMove + micro steps (main.c:431-438). */
int skidgpu_microstep(skidgpu_ctx *ctx, int nSteps, float fStep, skidgpu_log_cb cb, void *user);

/* Moved positions for kdOutVector (kd.c:1550-1608): iOrder[nMove] ascending, r3[3*nMove]. */
int skidgpu_get_moved(skidgpu_ctx *ctx, int *iOrder, float *r3);

/* kdInitpGroup + kdCalcCenter (main.c:452-453).  piGroup_by_iOrder: n ints (NULL ok).
 * g: nGroup entries (entry 0 = the non-group; NULL ok). */
int skidgpu_centers(skidgpu_ctx *ctx, int *piGroup_by_iOrder, skidgpu_pgroup *g);

/* The -unbind restart path (main.c:349-373): kdInGroup + kdInitpGroup + kdReadCenter.
 * centres: nGroup entries with rCenter/vcm filled from a .gtp, or NULL for the
 * centre-of-mass fallback (kd.c:1160-1193). */
int skidgpu_set_groups(skidgpu_ctx *ctx, const int *piGroup_by_iOrder, int nGroup,
                       const skidgpu_pgroup *centres);

/* kdUnbind + kdTooSmall (main.c:469-471).  fG = gravitational constant, z = redshift,
 * fCosmo = a*H(a) (kd.c:1317-1318; cosmology stays on the host), inType-dependent
 * potential update and the always-on scoop potential exactly as kd.c:1379-1446.
 * Outputs: final labels by iOrder (n ints), catalogue g (*nGroup entries, entry 0 =
 * non-group), *nGroup (= groups + 1), *nUnbound, *nGroupBefore (= "Groups before Unbind"). */
int skidgpu_unbind(skidgpu_ctx *ctx, float fG, float z, double fCosmo, int iSoftType,
                   float fScoop, int bNoUnbind, int nMaxMembers, int nMinMembers,
                   int *piGroup_by_iOrder, skidgpu_pgroup *g, int *nGroup, int *nUnbound,
                   int *nGroupBefore);

/* kdOutStats (kd.c:1703-1839; main.c:482-484 "-stats"): per final group, the members sorted by distance
 * from rCenter and the reference's sequential float32 accumulation, on the device.  Call after
 * skidgpu_unbind.  rows: nGroup entries (entry 0 = non-group, zeroed).  fExpHub = a*H(a) as passed to
 * skidgpu_unbind (kd.c:1731); fDensMin/fTempMax = the -d / -t cuts that define "gas mass" (kd.c:1792-1794).
 * The writer prints, per group ig >= 1 (kd.c:1822-1836, every value with "%g"):
 *   ig nMembers fTotMass fGasMass fStarMass sqrt(fVcirc) sqrt(fmVcirc) sqrt(flVcirc) fRVmax fRhmass
 *   sqrt(fRouter2) (float)sqrt(fVdispSum/(3.0*nMembers)) rCenter[3] vcm[3] rBound[3]
 * with the square roots taken in double on the float stored here. */
typedef struct {
	int nMembers;
	float fTotMass, fGasMass, fStarMass;
	float fVcirc, fmVcirc, flVcirc; /* G*M(<r)/r at the maximum, at the half-mass radius, at the outermost member */
	float fRVmax, fRhmass;          /* radius of the maximum, half-mass radius */
	float fRouter2;                 /* squared radius of the outermost member */
	float fVdispSum;                /* sum over members and axes of dv^2 */
} skidgpu_stat_row;
int skidgpu_stats(skidgpu_ctx *ctx, float fG, float z, double fExpHub, float fDensMin, float fTempMax,
                  skidgpu_stat_row *rows);

/* Per-stage device time of the last call of each stage, in milliseconds (CUDA events on the
 * context's stream).  stage: 0 tree+density, 1 move, 2 fof, 3 microstep, 4 centres, 5 unbind. */
double skidgpu_stage_ms(skidgpu_ctx *ctx, int stage);

/* Counters for bench.py: kernels launched since create, mover-steps (sum over steps of
 * active movers), entity-hit interactions are not counted on the device. */
long long skidgpu_counter(skidgpu_ctx *ctx, int which); /* 0 launches, 1 mover-steps, 2 kNN queries, 3 unbind pair evaluations */

/* Device time (ms, CUDA events on the context's stream, recorded without synchronising the host) spent in
 * each kernel family during the last call of its stage; recorded only after skidgpu_set_profile(ctx, 1).
 * which: 0 = k_tile_step (gradient walk + move, one launch per step), 1 = k_knn_density, 2 = tile-list builds
 * (re-sort, k_super_walk, k_tile_filter, k_tile_walk), 3 = per-step fallbacks (short-tile walks, own-walk movers),
 * 4 = prune + counts.  *nLaunches (nullable) = number of spans covered. */
double skidgpu_kernel_ms(skidgpu_ctx *ctx, int which, int *nLaunches);
int skidgpu_set_profile(skidgpu_ctx *ctx, int bOn);

/* Test hook: 0 = tile kernels (default), 1 = the v1 kernel (a tree walk per mover and step) for regression
 * comparisons.  Set before skidgpu_move. */
int skidgpu_debug_move_kernel(skidgpu_ctx *ctx, int which);

/* The CUDA stream (cudaStream_t) all work of this context is issued on, for callers that want to
 * bracket calls with their own events. */
void *skidgpu_stream(skidgpu_ctx *ctx);

/* Test hooks for the hand-written device primitives (stable LSD radix sort of (key,val) pairs on
 * the low `bits` key bits; exclusive prefix sum with out[n] = total).  Host arrays in and out. */
int skidgpu_debug_sort(skidgpu_ctx *ctx, unsigned long long *keys, unsigned int *vals, long long n, int bits);
int skidgpu_debug_scan(skidgpu_ctx *ctx, const unsigned int *in, unsigned int *out, long long n);

#ifdef __cplusplus
}
#endif
#endif
