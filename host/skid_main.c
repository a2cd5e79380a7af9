/*
 * skid (B200): drop-in driver with SKID v1.4.1's command line, stdin TIPSY input and
 * .grp/.gtp/.den/.ray/.stat outputs (reference main.c:17-499), calling the GPU stages through
 * include/skidgpu.h.  Flags are the reference's (main.c:140-335) plus two extensions:
 *   -nsp       never prune scatterers (README:29-32 describes it; main.c never implemented it)
 *   -gpu <n>   CUDA device ordinal (default 0)
 *   -gpus <N>  share the run between N GPUs (devices n .. n+N-1) of this node: one context and one host thread per
 *              GPU, every context gets the snapshot and makes the same stage calls, the library exchanges what the
 *              ranks must agree on with NCCL over NVLink (include/skidgpu.h: skidgpu_comm_init); rank 0 writes the
 *              outputs.  Results are those of one GPU.
 */
#include <float.h>
#include <limits.h>
#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "skid_host.h"

#define PRUNE_STEPS 5
#define MICRO_STEP 0.1

static void usage(void)
{
	fputs("USAGE:\n"
	      "skid -tau <fLinkLength> [OPTIONAL ARGUMENTS]\n"
	      "	 reads TIPSY BINARY input file from stdin\n"
	      "     [-std]\n"
	      "COSMOLOGY and UNITS arguments:\n"
	      "     [-z <fRedShift>] [-O <fOmega>]\n"
	      "     [-G <fGravConst>] [-H <fHubble>]\n"
	      "     [-Lambda <fLambda>] [-Q <fQuintessense]\n"
	      "GROUP FINDING arguments (see man page!):\n"
	      "     [-s <nSmooth>] [-d <fMinDensity>] [-t <fMaxTemp>]\n"
	      "     [-cvg <fConvergeRadius>] [-scoop <fScoopRadius>]\n"
	      "     [-m <nMinMembers>] [-nu] [-gd] [-unbind <GroupName>[.grp]]\n"
	      "     [-M <fMaxMass>] [-fic] [-go] [-maxgroup nMaxMembers] [-nsp]\n"
	      "GRAVITATIONAL SOFTENING arguments:\n"
	      "     [-spline] [-plummer] [-e <fSoft>]\n"
	      "PERIODIC BOX specification:\n"
	      "     [-p <xyzPeriod>]\n"
	      "     [-c <xyzCenter>]\n"
	      "     [-cx <xCenter>] [-cy <yCenter>] [-cz <zCenter>]\n"
	      "OUTPUT arguments:\n"
	      "     [-o <Output Name>] [-ray] [-den] [-stats] [-diag] [-gpu <device>] [-gpus <nGpu>]\n"
	      "\nSee man page skid(1).\n",
	      stderr);
	exit(1);
}

typedef struct {
	int reserved0; /* keeps every real field at a non-zero offset: offset 0 means "no flag field" in g_opts */
	int bTau, bCvg, bScoop, bEps, bPeriodic, bStandard;
	int nSmooth, nMembers, nMaxMembers, iSoftType;
	int bNoUnbind, bGasAndDark, bGasOnly, bUnbindOnly, bForceInitialCut, bNoPrune;
	int bOutRay, bOutDens, bOutStats, bOutDiag, iDevice, nGpus;
	float fTau, z, Omega0, Lambda, fQuintess, G, H0;
	float fDensMin, fTempMax, fMassMax, fCvg, fScoop, fEps;
	float fPeriod[3], fCenter[3];
	char achGroup[256], achName[256];
} options;

enum { A_FLAG, A_FLOAT, A_INT, A_STR };
typedef struct {
	const char *name;
	int kind;
	size_t off;     /* value field */
	size_t off_set; /* "was given" flag field or 0 */
	int flagval;
} optdef;
#define OFF(f) offsetof(options, f)

static const optdef g_opts[] = {
    {"-tau", A_FLOAT, OFF(fTau), OFF(bTau), 0},
    {"-z", A_FLOAT, OFF(z), 0, 0},
    {"-O", A_FLOAT, OFF(Omega0), 0, 0},
    {"-Lambda", A_FLOAT, OFF(Lambda), 0, 0},
    {"-Q", A_FLOAT, OFF(fQuintess), 0, 0},
    {"-G", A_FLOAT, OFF(G), 0, 0},
    {"-H", A_FLOAT, OFF(H0), 0, 0},
    {"-s", A_INT, OFF(nSmooth), 0, 0},
    {"-d", A_FLOAT, OFF(fDensMin), 0, 0},
    {"-t", A_FLOAT, OFF(fTempMax), 0, 0},
    {"-M", A_FLOAT, OFF(fMassMax), 0, 0},
    {"-fic", A_FLAG, OFF(bForceInitialCut), 0, 1},
    {"-cvg", A_FLOAT, OFF(fCvg), OFF(bCvg), 0},
    {"-scoop", A_FLOAT, OFF(fScoop), OFF(bScoop), 0},
    {"-m", A_INT, OFF(nMembers), 0, 0},
    {"-maxgroup", A_INT, OFF(nMaxMembers), 0, 0},
    {"-nu", A_FLAG, OFF(bNoUnbind), 0, 1},
    {"-gd", A_FLAG, OFF(bGasAndDark), 0, 1},
    {"-go", A_FLAG, OFF(bGasOnly), 0, 1},
    {"-nsp", A_FLAG, OFF(bNoPrune), 0, 1},
    {"-unbind", A_STR, OFF(achGroup), OFF(bUnbindOnly), 0},
    {"-spline", A_FLAG, OFF(iSoftType), 0, SKIDGPU_SPLINE},
    {"-plummer", A_FLAG, OFF(iSoftType), 0, SKIDGPU_PLUMMER},
    {"-e", A_FLOAT, OFF(fEps), OFF(bEps), 0},
    {"-cx", A_FLOAT, OFF(fCenter[0]), 0, 0},
    {"-cy", A_FLOAT, OFF(fCenter[1]), 0, 0},
    {"-cz", A_FLOAT, OFF(fCenter[2]), 0, 0},
    {"-o", A_STR, OFF(achName), 0, 0},
    {"-ray", A_FLAG, OFF(bOutRay), 0, 1},
    {"-den", A_FLAG, OFF(bOutDens), 0, 1},
    {"-stats", A_FLAG, OFF(bOutStats), 0, 1},
    {"-diag", A_FLAG, OFF(bOutDiag), 0, 1},
    {"-std", A_FLAG, OFF(bStandard), 0, 1},
    {"-gpu", A_INT, OFF(iDevice), 0, 0},
    {"-gpus", A_INT, OFF(nGpus), 0, 0},
};

static void parse_args(int argc, char **argv, options *o)
{
	int i = 1, j;
	size_t k;
	memset(o, 0, sizeof *o);
	/* defaults, main.c:86-136 */
	o->Omega0 = 1.0f;
	o->G = 1.0f;
	o->nSmooth = 64;
	o->fTempMax = FLT_MAX;
	o->fMassMax = FLT_MAX;
	o->nMembers = 8;
	o->nMaxMembers = INT_MAX;
	o->iSoftType = SKIDGPU_SPLINE;
	o->nGpus = 1;
	for (j = 0; j < 3; ++j) o->fPeriod[j] = FLT_MAX;
	strcpy(o->achName, "skid");
	while (i < argc) {
		const char *a = argv[i];
		int matched = 0;
		/* -p and -c set all three axes at once (main.c:273-289) */
		if (!strcmp(a, "-p") || !strcmp(a, "-c")) {
			float v;
			if (++i >= argc) usage();
			v = atof(argv[i++]);
			for (j = 0; j < 3; ++j) (a[1] == 'p' ? o->fPeriod : o->fCenter)[j] = v;
			if (a[1] == 'p') o->bPeriodic = 1;
			continue;
		}
		for (k = 0; k < sizeof g_opts / sizeof g_opts[0]; ++k) {
			const optdef *d = &g_opts[k];
			char *base = (char *)o;
			if (strcmp(a, d->name)) continue;
			matched = 1;
			++i;
			if (d->kind == A_FLAG) {
				*(int *)(base + d->off) = d->flagval;
			} else {
				if (i >= argc) usage();
				if (d->kind == A_FLOAT) *(float *)(base + d->off) = atof(argv[i]);
				else if (d->kind == A_INT) *(int *)(base + d->off) = atoi(argv[i]);
				else {
					strncpy(base + d->off, argv[i], 255);
					(base + d->off)[255] = 0;
				}
				++i;
			}
			if (d->off_set) *(int *)(base + d->off_set) = 1;
			break;
		}
		if (!matched) usage();
	}
	if (!o->bTau) usage(); /* main.c:339 */
	if (o->nGpus < 1) usage();
	if (!o->bCvg) o->fCvg = 0.5 * o->fTau;
	if (!o->bScoop) o->fScoop = 2.0 * o->fTau;
}

static void log_cb(void *user, int kind, int iter, int nActive, int nScatter)
{
	(void)user;
	if (kind == 0) printf("Ittr:%d nActive:%d nScatter:%d\n", iter, nActive, nScatter);
	else printf("Microstep:%d nScatter:%d\n", iter, nScatter);
	fflush(stdout);
}

static void die(skidgpu_ctx *ctx, const char *what)
{
	fprintf(stderr, "ERROR: %s: %s\n", what, skidgpu_last_error(ctx));
	exit(1);
}

static void print_time(const char *label, double ms)
{
	long us = (long)(ms * 1000.0 + 0.5);
	printf("%s%ld.%06ld\n", label, us / 1000000, us % 1000000);
}

/* The ASCII writers run on a helper thread while the GPU works on the next stage (the .den is formatted
 * during the move loop, the .ray during unbinding): their inputs are final by then. */
typedef struct {
	pthread_t th;
	int running, rc;
	int (*fn)(void *);
	void *arg;
} bg_job;

static void *bg_tramp(void *p)
{
	bg_job *j = (bg_job *)p;
	j->rc = j->fn(j->arg);
	return NULL;
}

static void bg_start(bg_job *j, int (*fn)(void *), void *arg)
{
	j->fn = fn;
	j->arg = arg;
	j->rc = 0;
	j->running = pthread_create(&j->th, NULL, bg_tramp, j) == 0;
	if (!j->running) j->rc = fn(arg);
}

static int bg_wait(bg_job *j)
{
	if (j->running) {
		pthread_join(j->th, NULL);
		j->running = 0;
	}
	return j->rc;
}

typedef struct {
	char path[300];
	const snapshot *s;
	const float *rho, *mvR, *fPeriod;
	const int *mvOrder;
	int nMove;
} write_job;

static int job_density(void *p)
{
	const write_job *w = (const write_job *)p;
	return out_density(w->path, w->s->n, w->rho);
}

static int job_vector(void *p)
{
	const write_job *w = (const write_job *)p;
	return out_vector(w->path, w->s, w->nMove, w->mvOrder, w->mvR, w->fPeriod);
}

static double wall(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* lap timer for SKID_HOST_TIMING: wall time of every host-side phase, in call order */
static struct {
	const char *name[32];
	double sec[32];
	int n;
	double last;
} g_lap;

static void lap(const char *name)
{
	const double t = wall();
	if (g_lap.n < 32) {
		g_lap.name[g_lap.n] = name;
		g_lap.sec[g_lap.n++] = t - g_lap.last;
	}
	g_lap.last = t;
}

typedef struct {
	skidgpu_ctx **pctx;
	const options *o;
} create_job;

static int job_create(void *p)
{
	const create_job *c = (const create_job *)p;
	int rc = skidgpu_create(c->pctx, c->o->iDevice, c->o->fPeriod, c->o->fCenter, c->o->bPeriodic, c->o->bOutDiag);
	if (!rc && getenv("SKID_PREALLOC_GB")) { /* experiment: pool growth off the critical path */
		double t = wall();
		skidgpu_reserve(*c->pctx, (unsigned long long)(atof(getenv("SKID_PREALLOC_GB")) * 1073741824.0));
		if (getenv("SKID_HOST_TIMING")) fprintf(stderr, "{\"reserve_s\": %.3f}\n", wall() - t);
	}
	return rc;
}

/* -gpus N: ranks 1 .. N-1.  Each runs the stage script of main() on its own context and device from its own host
 * thread, with no outputs (rank 0 = the main thread owns those); the library keeps the ranks in step. */
typedef struct {
	pthread_t th;
	const options *o;
	const snapshot *s;
	int rank;
	unsigned char id[SKIDGPU_UNIQUE_ID_BYTES];
	const int *piGroup; /* -unbind restart: the catalogue read by the main thread */
	int nGroup;
	const skidgpu_pgroup *centres;
	float fStep;
	double fCosmo;
} gpu_worker;

static void *worker_main(void *p)
{
	gpu_worker *w = (gpu_worker *)p;
	const options *o = w->o;
	skidgpu_ctx *ctx = NULL;
	int nx = 0, nMove = 0, nIttr = 0, nGroup = 1, nUnbound = 0, nBefore = 0;
	if (skidgpu_create(&ctx, o->iDevice + w->rank, o->fPeriod, o->fCenter, o->bPeriodic, 0)) die(NULL, "skidgpu_create (worker)");
	if (skidgpu_comm_init(ctx, w->id, w->rank, o->nGpus)) die(ctx, "skidgpu_comm_init (worker)");
	if (skidgpu_set_particles(ctx, w->s->p, w->s->n, w->s->nGas, w->s->nDark, w->s->nStar)) die(ctx, "skidgpu_set_particles (worker)");
	if (o->bUnbindOnly) {
		if (skidgpu_set_groups(ctx, w->piGroup, w->nGroup, w->centres)) die(ctx, "skidgpu_set_groups (worker)");
	} else {
		if (skidgpu_density(ctx, o->nSmooth, o->bGasAndDark, o->bGasOnly, NULL, NULL, &nx)) die(ctx, "skidgpu_density (worker)");
		if (skidgpu_move(ctx, o->fDensMin, o->fTempMax, o->fMassMax, o->fCvg, w->fStep, o->bForceInitialCut, o->bNoPrune, NULL,
		                 NULL, &nMove, &nIttr))
			die(ctx, "skidgpu_move (worker)");
		if (skidgpu_fof(ctx, o->fTau, &nGroup)) die(ctx, "skidgpu_fof (worker)");
		if (skidgpu_microstep(ctx, PRUNE_STEPS, MICRO_STEP * w->fStep, NULL, NULL)) die(ctx, "skidgpu_microstep (worker)");
		if (skidgpu_centers(ctx, NULL, NULL)) die(ctx, "skidgpu_centers (worker)");
	}
	if (o->bEps && skidgpu_set_soft(ctx, o->fEps)) die(ctx, "skidgpu_set_soft (worker)");
	if (skidgpu_unbind(ctx, o->G, o->z, w->fCosmo, o->iSoftType, o->fScoop, o->bNoUnbind, o->nMaxMembers, o->nMembers, NULL, NULL,
	                   &nGroup, &nUnbound, &nBefore))
		die(ctx, "skidgpu_unbind (worker)");
	skidgpu_destroy(ctx);
	return NULL;
}

static void start_workers(gpu_worker *w, const options *o, const snapshot *s, const int *piGroup, int nGroup,
                          const skidgpu_pgroup *centres, float fStep)
{
	int r;
	const float fShift = 1.0 / (1.0 + o->z);
	const double fCosmo = fShift * cosmo_exp2hub(fShift, o->H0, o->Omega0, o->Lambda, 0.0, o->fQuintess);
	if (skidgpu_comm_unique_id(w[0].id)) die(NULL, "skidgpu_comm_unique_id");
	for (r = 0; r < o->nGpus; ++r) {
		w[r].o = o;
		w[r].s = s;
		w[r].rank = r;
		if (r) memcpy(w[r].id, w[0].id, sizeof w[0].id);
		w[r].piGroup = piGroup;
		w[r].nGroup = nGroup;
		w[r].centres = centres;
		w[r].fStep = fStep;
		w[r].fCosmo = fCosmo;
		if (r && pthread_create(&w[r].th, NULL, worker_main, &w[r])) {
			fprintf(stderr, "ERROR: could not start the host thread of GPU %d\n", r);
			exit(1);
		}
	}
}

int main(int argc, char **argv)
{
	options o;
	snapshot s;
	skidgpu_ctx *ctx = NULL;
	skidgpu_pgroup *cat = NULL;
	int *piGroup = NULL;
	float *rho = NULL;
	float fStep;
	int nGroup = 1, nMove = 0, nIttr = 0, nUnbound = 0, nBefore = 0, nExtra = 0;
	char achFile[300];
	bg_job denJob = {0}, rayJob = {0}, createJob = {0};
	write_job denArgs, rayArgs;
	create_job createArgs;
	int *mvOrder = NULL;
	float *mvR = NULL;
	double t0 = wall(), tRead, tInit, tStages, tEnd; /* host wall clock, reported with SKID_HOST_TIMING=1 */
	gpu_worker *workers = NULL;
	int gtpRc = 0;

	g_lap.last = t0;
	printf("SKID v1.4.1 (B200 GPU hot path): group finder compatible with SKID v1.4.1, Stadel 2000\n");
	parse_args(argc, argv, &o);
	fStep = 0.5 * o.fCvg; /* main.c:345 */

	/* the CUDA context comes up on a helper thread while stdin is being read and decoded */
	createArgs.pctx = &ctx;
	createArgs.o = &o;
	bg_start(&createJob, job_create, &createArgs);
	if (tipsy_read(stdin, o.bStandard, &s)) {
		fprintf(stderr, "ERROR: could not read a TIPSY %s binary from stdin\n", o.bStandard ? "standard" : "native");
		return 1;
	}
	printf("nDark:%d nGas:%d nStar:%d\n", s.nDark, s.nGas, s.nStar);
	fflush(stdout);
	tRead = wall();
	lap("read_input");

	if (bg_wait(&createJob)) die(NULL, "skidgpu_create");
	lap("wait_context");
	piGroup = (int *)calloc((size_t)s.n, sizeof(int));
	rho = (float *)calloc((size_t)s.n, sizeof(float));
	if (o.bUnbindOnly) {
		/* main.c:349-373: strip a trailing .grp, read <name>.grp and, if present, <name>.gtp */
		size_t len = strlen(o.achGroup);
		if (len >= 4 && !strcmp(o.achGroup + len - 4, ".grp")) o.achGroup[len - 4] = 0;
		snprintf(achFile, sizeof achFile, "%s.grp", o.achGroup);
		nGroup = grp_read(achFile, s.n, piGroup);
		if (nGroup < 0) return 1;
		cat = (skidgpu_pgroup *)calloc((size_t)nGroup + 1, sizeof(skidgpu_pgroup));
		snprintf(achFile, sizeof achFile, "%s.gtp", o.achGroup);
		gtpRc = gtp_read(achFile, o.bStandard, nGroup, cat);
		if (gtpRc < 0) return 1;
	}
	if (o.nGpus > 1) { /* the other GPUs: one host thread each; this thread is rank 0 */
		workers = (gpu_worker *)calloc((size_t)o.nGpus, sizeof(gpu_worker));
		start_workers(workers, &o, &s, piGroup, nGroup, gtpRc ? cat : NULL, fStep);
		if (skidgpu_comm_init(ctx, workers[0].id, 0, o.nGpus)) die(ctx, "skidgpu_comm_init");
		lap("comm_init");
	}
	if (skidgpu_set_particles(ctx, s.p, s.n, s.nGas, s.nDark, s.nStar)) die(ctx, "skidgpu_set_particles");
	tInit = wall();
	lap("upload");

	if (o.bUnbindOnly) {
		if (skidgpu_set_groups(ctx, piGroup, nGroup, gtpRc ? cat : NULL)) die(ctx, "skidgpu_set_groups");
	} else {
		if (skidgpu_density(ctx, o.nSmooth, o.bGasAndDark, o.bGasOnly, rho, NULL, &nExtra)) die(ctx, "skidgpu_density");
		lap("density");
		if (o.bPeriodic) printf("nExtraScat:%d\n", nExtra);
		if (o.bOutDens) {
			snprintf(denArgs.path, sizeof denArgs.path, "%s.den", o.achName);
			denArgs.s = &s;
			denArgs.rho = rho;
			bg_start(&denJob, job_density, &denArgs);
		}
		if (skidgpu_move(ctx, o.fDensMin, o.fTempMax, o.fMassMax, o.fCvg, fStep, o.bForceInitialCut, o.bNoPrune,
		                 log_cb, NULL, &nMove, &nIttr))
			die(ctx, "skidgpu_move");
		lap("move");
		if (skidgpu_fof(ctx, o.fTau, &nGroup)) die(ctx, "skidgpu_fof");
		if (skidgpu_microstep(ctx, PRUNE_STEPS, MICRO_STEP * fStep, log_cb, NULL)) die(ctx, "skidgpu_microstep");
		lap("fof_microstep");
		if (o.bOutRay) {
			mvOrder = (int *)malloc((size_t)(nMove ? nMove : 1) * sizeof(int));
			mvR = (float *)malloc((size_t)(nMove ? nMove : 1) * 3 * sizeof(float));
			if (skidgpu_get_moved(ctx, mvOrder, mvR)) die(ctx, "skidgpu_get_moved");
			lap("get_moved");
			bg_wait(&denJob); /* one set of writer threads at a time */
			lap("wait_den_writer");
			snprintf(rayArgs.path, sizeof rayArgs.path, "%s.ray", o.achName);
			rayArgs.s = &s;
			rayArgs.nMove = nMove;
			rayArgs.mvOrder = mvOrder;
			rayArgs.mvR = mvR;
			rayArgs.fPeriod = o.fPeriod;
			bg_start(&rayJob, job_vector, &rayArgs);
		}
		cat = (skidgpu_pgroup *)calloc((size_t)nGroup + 1, sizeof(skidgpu_pgroup));
		if (skidgpu_centers(ctx, NULL, NULL)) die(ctx, "skidgpu_centers");
		lap("centers");
	}
	/* kdSetUniverse / kdSetSoft / kdUnbind / kdTooSmall (main.c:459-471) */
	if (o.bEps && skidgpu_set_soft(ctx, o.fEps)) die(ctx, "skidgpu_set_soft");
	{
		const float fShift = 1.0 / (1.0 + o.z);
		const double dHub = cosmo_exp2hub(fShift, o.H0, o.Omega0, o.Lambda, 0.0, o.fQuintess);
		const double fCosmo = fShift * dHub;
		if (skidgpu_unbind(ctx, o.G, o.z, fCosmo, o.iSoftType, o.fScoop, o.bNoUnbind, o.nMaxMembers, o.nMembers,
		                   piGroup, cat, &nGroup, &nUnbound, &nBefore))
			die(ctx, "skidgpu_unbind");
		lap("unbind");
		printf("Groups before Unbind:%d\n", nBefore);
		printf("Number of particles Unbound:%d\n", nUnbound);
		printf("Number of Groups:%d\n", nGroup - 1);
		fflush(stdout);
		tStages = wall();
		if (bg_wait(&denJob) | bg_wait(&rayJob)) fprintf(stderr, "WARNING: could not write the .den/.ray file\n");
		lap("wait_ray_writer");
		snprintf(achFile, sizeof achFile, "%s.grp", o.achName);
		out_group(achFile, s.n, piGroup);
		snprintf(achFile, sizeof achFile, "%s.gtp", o.achName);
		out_gtp(achFile, o.bStandard, s.time, nGroup, cat);
		lap("write_grp_gtp");
		if (o.bOutStats) {
			/* kdOutStats, main.c:482-484 */
			skidgpu_stat_row *rows = (skidgpu_stat_row *)calloc((size_t)nGroup + 1, sizeof(skidgpu_stat_row));
			if (skidgpu_stats(ctx, o.G, o.z, fCosmo, o.fDensMin, o.fTempMax, rows)) die(ctx, "skidgpu_stats");
			snprintf(achFile, sizeof achFile, "%s.stat", o.achName);
			out_stats(achFile, nGroup, cat, rows);
			free(rows);
			lap("stats");
		}
	}
	printf("SKID GPU Time:\n");
	if (!o.bUnbindOnly) {
		print_time("   Initial Density:    ", skidgpu_stage_ms(ctx, 0));
		print_time("   Moving Particles:   ", skidgpu_stage_ms(ctx, 1));
		print_time("   Friends of Friends: ", skidgpu_stage_ms(ctx, 2));
		print_time("   Microstepping:      ", skidgpu_stage_ms(ctx, 3));
	}
	if (!o.bNoUnbind) print_time("   Unbinding:          ", skidgpu_stage_ms(ctx, 5));
	fflush(stdout);
	tEnd = wall();
	if (getenv("SKID_HOST_TIMING")) {
		int k;
		fprintf(stderr, "{\"host_laps_s\": {");
		for (k = 0; k < g_lap.n; ++k) fprintf(stderr, "%s\"%s\": %.3f", k ? ", " : "", g_lap.name[k], g_lap.sec[k]);
		fprintf(stderr, "}}\n");
		fprintf(stderr,
		        "{\"host_wall_s\": {\"read\": %.3f, \"create_upload\": %.3f, \"stages_with_overlapped_writers\": %.3f, "
		        "\"final_writers\": %.3f, \"total\": %.3f}, \"n\": %d, \"threads\": %d}\n",
		        tRead - t0, tInit - tRead, tStages - tInit, tEnd - tStages, tEnd - t0, s.n, host_threads());
	}
	if (workers) {
		int r;
		for (r = 1; r < o.nGpus; ++r) pthread_join(workers[r].th, NULL);
		free(workers);
	}
	skidgpu_destroy(ctx);
	free(mvOrder);
	free(mvR);
	free(piGroup);
	free(rho);
	free(cat);
	free(s.p);
	return 0;
}
