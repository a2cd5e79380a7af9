/*
 * TEST INFRASTRUCTURE ONLY - see skid_oracle.h.  CPU restatement of SKID v1.4.1's hot path.
 * Compile with -ffp-contract=off (Makefile): the reference arithmetic is plain float32 without
 * FMA contraction (gcc x86-64 SSE2), and kNN radii are compared bitwise.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "skid_oracle.h"

#define ORC_PLUMMER 1

/* ---- shared float32 helpers ------------------------------------------------------------ */

/* min-image separation the way the reference forms it: the QUERY is shifted by +-L first
 * (sx = x +- lx, INTERSECT kd.h:136,148), then the particle is subtracted (smooth1.c:104). */
static float minimg(float x, float px, float L)
{
	float d = x - px;
	if (L > 0.0f) {
		float h = 0.5f * L;
		if (d > h) {
			float sx = x - L;
			d = sx - px;
		} else if (d < -h) {
			float sx = x + L;
			d = sx - px;
		}
	}
	return d;
}

static float dist2f(float dx, float dy, float dz)
{
	float a = dx * dx, b = dy * dy, c = dz * dz; /* smooth1.c:107: dx*dx + dy*dy + dz*dz */
	float s = a + b;
	return s + c;
}

/* ---- kNN + density ---------------------------------------------------------------------- */

typedef struct {
	float d2;
	int j;
} cand;

static int cmp_cand(const void *a, const void *b)
{
	const cand *x = (const cand *)a, *y = (const cand *)b;
	if (x->d2 < y->d2) return -1;
	if (x->d2 > y->d2) return 1;
	return (x->j > y->j) - (x->j < y->j);
}

void orc_knn_density(int n, const float *pos, const float *mass, int k, float period, float *ball2,
                     float *rho, int *nbr, float *nbrd2)
{
	cand *c = (cand *)malloc((size_t)n * sizeof(cand));
	cand *best = (cand *)malloc((size_t)n * (size_t)k * sizeof(cand));
	int i, j, e;
	for (i = 0; i < n; ++i) rho[i] = 0.0f;
	/* brute force: all n distances per query, keep the k smallest by (d2, index) */
	for (i = 0; i < n; ++i) {
		const float x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
		int m = 0;
		float worst = FLT_MAX;
		for (j = 0; j < n; ++j) {
			float d2 = dist2f(minimg(x, pos[3 * j], period), minimg(y, pos[3 * j + 1], period),
			                  minimg(z, pos[3 * j + 2], period));
			if (m < 2 * k || d2 <= worst) {
				c[m].d2 = d2;
				c[m].j = j;
				++m;
				if (m == n || m >= 8 * k) { /* compact: keep the k best, tighten the cut */
					qsort(c, (size_t)m, sizeof(cand), cmp_cand);
					m = k;
					worst = c[k - 1].d2;
				}
			}
		}
		qsort(c, (size_t)m, sizeof(cand), cmp_cand);
		memcpy(best + (size_t)i * k, c, (size_t)k * sizeof(cand));
		ball2[i] = c[k - 1].d2; /* smooth1.c:243: fBall2 = key of the farthest of the k */
	}
	/* density (smooth1.c:249-263): every neighbour except the farthest, symmetric gather + scatter */
	for (i = 0; i < n; ++i) {
		const cand *b = best + (size_t)i * k;
		float h2 = ball2[i];
		float ih2 = 4.0 / h2;
		float fNorm = 0.5 * M_1_PI * sqrt(ih2) * ih2;
		for (e = 0; e < k - 1; ++e) {
			float r2 = b[e].d2 * ih2;
			float rs = 2.0 - sqrt(r2);
			if (r2 < 1.0) rs = (1.0 - 0.75 * rs * r2);
			else rs = 0.25 * rs * rs * rs;
			rs *= fNorm;
			rho[i] += rs * mass[b[e].j];
			rho[b[e].j] += rs * mass[i];
		}
		if (nbr)
			for (e = 0; e < k; ++e) {
				nbr[(size_t)i * k + e] = b[e].j;
				if (nbrd2) nbrd2[(size_t)i * k + e] = b[e].d2;
			}
	}
	free(c);
	free(best);
}

/* ---- replicas ---------------------------------------------------------------------------- */

int orc_replicas(int n, const float *pos, const float *ball2, float period, const float *center, int cap,
                 int *rep_src, float *rep_pos)
{
	float lo[3], hi[3];
	int i, ix, iy, iz, d, cnt = 0;
	for (d = 0; d < 3; ++d) { /* smooth1.c:283-286 */
		lo[d] = center[d] - 0.5 * period;
		hi[d] = center[d] + 0.5 * period;
	}
	for (i = 0; i < n; ++i)
		for (ix = -1; ix <= 1; ++ix)
			for (iy = -1; iy <= 1; ++iy)
				for (iz = -1; iz <= 1; ++iz) {
					float q[3], d2 = 0.0f;
					if (!ix && !iy && !iz) continue;
					q[0] = pos[3 * i] + ix * period;
					q[1] = pos[3 * i + 1] + iy * period;
					q[2] = pos[3 * i + 2] + iz * period;
					for (d = 0; d < 3; ++d) { /* INTERSECTNP kd.h:102-119 */
						float a = lo[d] - q[d], b = q[d] - hi[d];
						if (a > 0.0f) d2 += a * a;
						else if (b > 0.0f) d2 += b * b;
					}
					if (d2 < ball2[i]) {
						if (rep_src && cnt < cap) {
							rep_src[cnt] = i;
							rep_pos[3 * cnt] = q[0];
							rep_pos[3 * cnt + 1] = q[1];
							rep_pos[3 * cnt + 2] = q[2];
						}
						++cnt;
					}
				}
	return cnt;
}

/* ---- gradient, gather form, brute force ---------------------------------------------------- */

/* kernel-gradient weight of smAccDensity (smooth1.c:447-454) for one (scatterer, mover) hit */
static float grad_weight(float d2, float ball2, float m)
{
	float ih2 = 4.0 / ball2;
	float fNorm = M_1_PI * ih2 * ih2 * sqrt(ih2) * m;
	float r2 = d2 * ih2;
	float rs = sqrt(r2);
	if (r2 < 1.0) rs = -3.0 + 2.25 * rs;
	else rs = -3.0 / rs + 3.0 - 0.75 * rs;
	rs *= fNorm;
	return rs;
}

float orc_gradient(int nEnt, const float *epos, const float *eball2, const float *emass, const float *erho,
                   const unsigned char *ent_alive, int nMove, const float *mpos, float *acc,
                   unsigned char *touched)
{
	float fScatDens = 0.0f;
	int e, m;
	memset(acc, 0, (size_t)nMove * 3 * sizeof(float));
	if (touched) memset(touched, 0, (size_t)nEnt);
	/* scatterer-major like the reference (smooth1.c:436-471) so that float32 sums add in scatterer order */
	for (e = 0; e < nEnt; ++e) {
		int any = 0;
		if (ent_alive && !ent_alive[e]) continue;
		for (m = 0; m < nMove; ++m) {
			float dx = epos[3 * e] - mpos[3 * m], dy = epos[3 * e + 1] - mpos[3 * m + 1],
			      dz = epos[3 * e + 2] - mpos[3 * m + 2];
			float d2 = dist2f(dx, dy, dz);
			if (d2 < eball2[e]) {
				float w = grad_weight(d2, eball2[e], emass[e]);
				acc[3 * m] += dx * w;
				acc[3 * m + 1] += dy * w;
				acc[3 * m + 2] += dz * w;
				any = 1;
			}
		}
		if (any) {
			if (touched) touched[e] = 1;
			if (fScatDens == 0.0f || erho[e] < fScatDens) fScatDens = erho[e];
		}
	}
	return fScatDens;
}

/* Sum over the hits of a mover of the MAGNITUDES of the gradient terms, S_m = sum |x_e - x_m| |w| : the scale
 * against which the float32 summation noise of smAccDensity (smooth1.c:447-459) is judged - SURVEY 8c:
 * |a_gpu - a_ref| <= 1e-5 S_m, where |a| itself can be arbitrarily small by cancellation. */
void orc_gradient_abs(int nEnt, const float *epos, const float *eball2, const float *emass, int nMove,
                      const float *mpos, double *sabs)
{
	int e, m;
	for (m = 0; m < nMove; ++m) sabs[m] = 0.0;
	for (e = 0; e < nEnt; ++e)
		for (m = 0; m < nMove; ++m) {
			float dx = epos[3 * e] - mpos[3 * m], dy = epos[3 * e + 1] - mpos[3 * m + 1],
			      dz = epos[3 * e + 2] - mpos[3 * m + 2];
			float d2 = dist2f(dx, dy, dz);
			if (d2 < eball2[e]) sabs[m] += sqrt((double)d2) * fabs((double)grad_weight(d2, eball2[e], emass[e]));
		}
}

/* ---- the flow loop: scatter form over a uniform grid of the ACTIVE movers ------------------- */

typedef struct {
	int ng;        /* cells per axis */
	float lo, inv; /* grid origin and 1/cell */
	int *head, *next;
} mgrid;

static int cell_of(const mgrid *g, float x)
{
	int c = (int)floorf((x - g->lo) * g->inv);
	if (c < 0) c = 0;
	if (c >= g->ng) c = g->ng - 1;
	return c;
}

static void grid_fill(mgrid *g, int nAct, const int *act, const float *mpos)
{
	int i;
	memset(g->head, 0xff, (size_t)g->ng * g->ng * g->ng * sizeof(int));
	for (i = 0; i < nAct; ++i) {
		int m = act[i];
		int c = (cell_of(g, mpos[3 * m + 2]) * g->ng + cell_of(g, mpos[3 * m + 1])) * g->ng + cell_of(g, mpos[3 * m]);
		g->next[m] = g->head[c];
		g->head[c] = m;
	}
}

/* one smAccDensity pass (smooth1.c:408-518) over alive entities; returns fScatDens */
static float scatter_pass(int nEnt, const float *epos, const float *eball2, const float *emass, const float *erho,
                          unsigned char *alive, const mgrid *g, const float *mpos, float *acc, int bInitial,
                          float *erho_mut)
{
	float fScatDens = 0.0f;
	int e, cx, cy, cz;
	for (e = 0; e < nEnt; ++e) {
		float h, x, y, z;
		int c0[3], c1[3], any = 0;
		if (!alive[e]) continue;
		x = epos[3 * e];
		y = epos[3 * e + 1];
		z = epos[3 * e + 2];
		h = sqrtf(eball2[e]) * 1.0001f;
		c0[0] = cell_of(g, x - h);
		c1[0] = cell_of(g, x + h);
		c0[1] = cell_of(g, y - h);
		c1[1] = cell_of(g, y + h);
		c0[2] = cell_of(g, z - h);
		c1[2] = cell_of(g, z + h);
		for (cz = c0[2]; cz <= c1[2]; ++cz)
			for (cy = c0[1]; cy <= c1[1]; ++cy)
				for (cx = c0[0]; cx <= c1[0]; ++cx) {
					int m = g->head[(cz * g->ng + cy) * g->ng + cx];
					for (; m >= 0; m = g->next[m]) {
						float dx = x - mpos[3 * m], dy = y - mpos[3 * m + 1], dz = z - mpos[3 * m + 2];
						float d2 = dist2f(dx, dy, dz);
						if (d2 < eball2[e]) {
							float w = grad_weight(d2, eball2[e], emass[e]);
							acc[3 * m] += dx * w;
							acc[3 * m + 1] += dy * w;
							acc[3 * m + 2] += dz * w;
							any = 1;
						}
					}
				}
		if (any) {
			if (fScatDens == 0.0f || erho[e] < fScatDens) fScatDens = erho[e];
		} else if (bInitial) {
			erho_mut[e] = 0.0f; /* smooth1.c:463-470 */
		}
	}
	return fScatDens;
}

static void move_particles(int nAct, const int *act, float *mpos, const float *acc, float fStep, float period,
                           const float *center)
{
	int i, j;
	for (i = 0; i < nAct; ++i) { /* kdMoveParticles kd.c:711-729 */
		int m = act[i];
		float ax = acc[3 * m], ay = acc[3 * m + 1], az = acc[3 * m + 2];
		float ai = sqrt(ax * ax + ay * ay + az * az);
		if (ai > 0.0) ai = fStep / sqrt(ax * ax + ay * ay + az * az);
		else ai = 0.0;
		mpos[3 * m] -= ai * ax;
		mpos[3 * m + 1] -= ai * ay;
		mpos[3 * m + 2] -= ai * az;
		if (period > 0.0f)
			for (j = 0; j < 3; ++j) {
				if (mpos[3 * m + j] > center[j] + 0.5 * period) mpos[3 * m + j] -= period;
				if (mpos[3 * m + j] <= center[j] - 0.5 * period) mpos[3 * m + j] += period;
			}
	}
}

int orc_move_loop(int nEnt, const float *epos, const float *eball2, const float *emass, float *erho,
                  int nMove, float *mpos, float period, const float *center, float fCvg, float fStep,
                  int bInitial, int bNoPrune, int maxlog, int *log_nactive, int *log_nscatter,
                  int nMicro, float fMicroStep, float *mpos_at_fof)
{
	unsigned char *alive = (unsigned char *)malloc((size_t)(nEnt ? nEnt : 1));
	float *acc = (float *)calloc((size_t)(nMove ? nMove : 1) * 3, sizeof(float));
	float *rold = (float *)malloc((size_t)(nMove ? nMove : 1) * 3 * sizeof(float));
	int *act = (int *)malloc((size_t)(nMove ? nMove : 1) * sizeof(int));
	mgrid g;
	int nAct = nMove, nIttr = 0, step, i, e, nScat;
	float lo = FLT_MAX, hi = -FLT_MAX;
	if (bNoPrune) bInitial = 0;
	for (e = 0; e < nEnt; ++e) alive[e] = 1;
	for (i = 0; i < nMove; ++i) act[i] = i;
	memcpy(rold, mpos, (size_t)nMove * 3 * sizeof(float));
	/* grid over the region the movers can occupy */
	if (period > 0.0f) {
		lo = center[0] - 0.5f * period;
		hi = center[0] + 0.5f * period;
		for (i = 1; i < 3; ++i) {
			if (center[i] - 0.5f * period < lo) lo = center[i] - 0.5f * period;
			if (center[i] + 0.5f * period > hi) hi = center[i] + 0.5f * period;
		}
	} else {
		for (i = 0; i < 3 * nMove; ++i) {
			if (mpos[i] < lo) lo = mpos[i];
			if (mpos[i] > hi) hi = mpos[i];
		}
		lo -= 0.05f * (hi - lo) + 1e-6f;
		hi += 0.05f * (hi - lo) + 1e-6f;
	}
	g.ng = 64;
	g.lo = lo;
	g.inv = g.ng / (hi - lo);
	g.head = (int *)malloc((size_t)g.ng * g.ng * g.ng * sizeof(int));
	g.next = (int *)malloc((size_t)(nMove ? nMove : 1) * sizeof(int));

#define ORC_ONE_STEP(initial, stepLen)                                                                 \
	do {                                                                                           \
		float fsd_;                                                                            \
		for (i = 0; i < nAct; ++i) acc[3 * act[i]] = acc[3 * act[i] + 1] = acc[3 * act[i] + 2] = 0.0f; \
		grid_fill(&g, nAct, act, mpos);                                                        \
		fsd_ = scatter_pass(nEnt, epos, eball2, emass, erho, alive, &g, mpos, acc, (initial), erho); \
		if (!bNoPrune) /* ScatterCut smooth1.c:509-513 */                                      \
			for (e = 0; e < nEnt; ++e)                                                     \
				if (alive[e] && !(erho[e] >= fsd_)) alive[e] = 0;                      \
		move_particles(nAct, act, mpos, acc, (stepLen), period, center);                       \
	} while (0)

	/* step 0 (main.c:396-404) */
	ORC_ONE_STEP(bInitial, fStep);
	nScat = 0;
	for (e = 0; e < nEnt; ++e) nScat += alive[e];
	if (nIttr < maxlog) {
		log_nactive[nIttr] = nAct;
		log_nscatter[nIttr] = nScat;
	}
	++nIttr;
	while (nAct) { /* main.c:409-419 */
		int keep = 0;
		const float hx = 0.5 * period, fCvg2 = fCvg * fCvg;
		for (step = 0; step < 5; ++step) ORC_ONE_STEP(0, fStep);
		for (i = 0; i < nAct; ++i) { /* kdPruneInactive kd.c:735-793 */
			int m = act[i], j;
			float d[3];
			for (j = 0; j < 3; ++j) {
				d[j] = mpos[3 * m + j] - rold[3 * m + j];
				if (period > 0.0f) {
					if (d[j] > hx) d[j] -= 2 * hx;
					if (d[j] <= -hx) d[j] += 2 * hx;
				}
			}
			if (dist2f(d[0], d[1], d[2]) >= fCvg2) {
				act[keep++] = m;
				for (j = 0; j < 3; ++j) rold[3 * m + j] = mpos[3 * m + j];
			}
		}
		nAct = keep;
		nScat = 0;
		for (e = 0; e < nEnt; ++e) nScat += alive[e];
		if (nIttr < maxlog) {
			log_nactive[nIttr] = nAct;
			log_nscatter[nIttr] = nScat;
		}
		++nIttr;
	}
	if (mpos_at_fof) memcpy(mpos_at_fof, mpos, (size_t)nMove * 3 * sizeof(float));
	/* micro steps after FoF (main.c:431-438): every mover is active again */
	nAct = nMove;
	for (i = 0; i < nMove; ++i) act[i] = i;
	for (step = 0; step < nMicro; ++step) ORC_ONE_STEP(0, fMicroStep);
#undef ORC_ONE_STEP
	free(g.head);
	free(g.next);
	free(alive);
	free(acc);
	free(rold);
	free(act);
	return nIttr;
}

/* ---- friends of friends ---------------------------------------------------------------------- */

static int uf_find(int *p, int x)
{
	while (p[x] != x) {
		p[x] = p[p[x]];
		x = p[x];
	}
	return x;
}

int orc_fof(int n, const float *pos, float tau, float period, int *label)
{
	/* connected components of {min-image d2 < tau^2} (kd.c:869-883): cell list with cells >= tau */
	int *parent = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
	int *head, *next = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
	int *first;
	float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, cs[3];
	int ng[3], i, j, d, G = 0;
	const float tau2 = tau * tau;
	for (i = 0; i < n; ++i)
		for (d = 0; d < 3; ++d) {
			if (pos[3 * i + d] < lo[d]) lo[d] = pos[3 * i + d];
			if (pos[3 * i + d] > hi[d]) hi[d] = pos[3 * i + d];
		}
	for (d = 0; d < 3; ++d) {
		float ext = hi[d] - lo[d];
		if (period > 0.0f) {
			lo[d] = -FLT_MAX; /* periodic: grid spans exactly one period, set below */
			ext = period;
		}
		ng[d] = (int)floor(ext / tau);
		if (ng[d] < 1) ng[d] = 1;
		if (ng[d] > 256) ng[d] = 256;
		cs[d] = ext / ng[d];
	}
	if (period > 0.0f) {
		/* any origin works for a periodic grid; use the minimum coordinate */
		for (d = 0; d < 3; ++d) {
			lo[d] = FLT_MAX;
			for (i = 0; i < n; ++i)
				if (pos[3 * i + d] < lo[d]) lo[d] = pos[3 * i + d];
		}
	}
	head = (int *)malloc((size_t)ng[0] * ng[1] * ng[2] * sizeof(int));
	memset(head, 0xff, (size_t)ng[0] * ng[1] * ng[2] * sizeof(int));
#define ORC_CELL(i, d) ((int)fmin(ng[d] - 1, fmax(0, floor((pos[3 * (i) + (d)] - lo[d]) / cs[d]))))
	for (i = 0; i < n; ++i) {
		int c = (ORC_CELL(i, 2) * ng[1] + ORC_CELL(i, 1)) * ng[0] + ORC_CELL(i, 0);
		parent[i] = i;
		next[i] = head[c];
		head[c] = i;
	}
	for (i = 0; i < n; ++i) {
		int c[3], o[3];
		for (d = 0; d < 3; ++d) c[d] = ORC_CELL(i, d);
		for (o[2] = -1; o[2] <= 1; ++o[2])
			for (o[1] = -1; o[1] <= 1; ++o[1])
				for (o[0] = -1; o[0] <= 1; ++o[0]) {
					int b[3], ok = 1;
					for (d = 0; d < 3; ++d) {
						b[d] = c[d] + o[d];
						if (period > 0.0f) b[d] = (b[d] + ng[d]) % ng[d];
						else if (b[d] < 0 || b[d] >= ng[d]) ok = 0;
					}
					if (!ok) continue;
					for (j = head[(b[2] * ng[1] + b[1]) * ng[0] + b[0]]; j >= 0; j = next[j]) {
						float d2;
						if (j >= i) continue;
						if (uf_find(parent, i) == uf_find(parent, j)) continue;
						d2 = dist2f(minimg(pos[3 * i], pos[3 * j], period), minimg(pos[3 * i + 1], pos[3 * j + 1], period),
						            minimg(pos[3 * i + 2], pos[3 * j + 2], period));
						if (d2 < tau2) parent[uf_find(parent, i)] = uf_find(parent, j);
					}
				}
	}
#undef ORC_CELL
	first = (int *)malloc((size_t)(n ? n : 1) * sizeof(int));
	for (i = 0; i < n; ++i) first[i] = 0;
	for (i = 0; i < n; ++i) {
		int r = uf_find(parent, i);
		if (!first[r]) first[r] = ++G;
		label[i] = first[r];
	}
	free(first);
	free(parent);
	free(head);
	free(next);
	return G;
}

/* ---- unbinding of one group --------------------------------------------------------------------- */

/* SPLINE_POT (grav.h:11-31) / Plummer (grav.c:24-26); returns the float "dir" */
static float soft_dir(float d2, float twoh, int iSoftType)
{
	float dir;
	if (iSoftType == ORC_PLUMMER) {
		dir = 1.0 / sqrt(d2 + 0.25 * twoh * twoh);
	} else {
		double r = sqrt(d2), a;
		if (r < twoh) {
			double dih = 2.0 / twoh, u = r * dih;
			if (u < 1.0) a = dih * (7.0 / 5.0 - 2.0 / 3.0 * u * u + 3.0 / 10.0 * u * u * u * u - 1.0 / 10.0 * u * u * u * u * u);
			else {
				double ir = 1.0 / r;
				a = -1.0 / 15.0 * ir + dih * (8.0 / 5.0 - 4.0 / 3.0 * u * u + u * u * u - 3.0 / 10.0 * u * u * u * u +
				                              1.0 / 30.0 * u * u * u * u * u);
			}
		} else a = 1.0 / r;
		dir = a;
	}
	return dir;
}

int orc_unbind_group(int n, const float *r, const float *v, const float *mass, const float *soft, int nScoop,
                     const float *sr, const float *smass, const float *ssoft, float G, float z, float fCosmo,
                     int iSoftType, int bNoUnbind, int bSubPot, unsigned char *removed, double *boundMass,
                     double *vcmOut)
{
	double *pot = (double *)calloc((size_t)(n ? n : 1), sizeof(double));
	int *idx = (int *)malloc((size_t)(n ? n : 1) * sizeof(int)); /* current arrangement -> member index */
	const float fShift = 1.0 / (1.0 + z);
	double dMass = 0.0, rcm[3] = {0, 0, 0}, vcm[3] = {0, 0, 0};
	int i, j, k, m = n, nRemoved = 0;
	for (i = 0; i < n; ++i) {
		idx[i] = i;
		removed[i] = 0;
	}
	/* kdCellPot (grav.c:8-36) */
	for (i = 0; i < n - 1; ++i)
		for (j = i + 1; j < n; ++j) {
			float dx = r[3 * i] - r[3 * j], dy = r[3 * i + 1] - r[3 * j + 1], dz = r[3 * i + 2] - r[3 * j + 2];
			float dir = soft_dir(dist2f(dx, dy, dz), soft[i] + soft[j], iSoftType);
			pot[i] += G * mass[j] * dir;
			pot[j] += G * mass[i] * dir;
		}
	/* kdAddScoopPot (grav.c:107-131); always applied (kd.c:1379-1381 is always true) */
	for (k = 0; k < nScoop; ++k)
		for (i = 0; i < n; ++i) {
			float dx = sr[3 * k] - r[3 * i], dy = sr[3 * k + 1] - r[3 * i + 1], dz = sr[3 * k + 2] - r[3 * i + 2];
			float dir = soft_dir(dist2f(dx, dy, dz), ssoft[k] + soft[i], iSoftType);
			pot[i] += G * smass[k] * dir;
		}
	for (i = 0; i < n; ++i) { /* kd.c:1360-1375 */
		dMass += mass[i];
		for (j = 0; j < 3; ++j) {
			rcm[j] += mass[i] * r[3 * i + j];
			vcm[j] += mass[i] * v[3 * i + j];
		}
	}
	for (j = 0; j < 3; ++j) {
		rcm[j] /= dMass;
		vcm[j] /= dMass;
	}
	while (1) { /* kd.c:1385-1447 */
		int iBig = 0;
		float fTotBig = -1.0;
		for (i = 0; i < m; ++i) {
			int p = idx[i];
			float dv2 = 0.0, fTot;
			for (j = 0; j < 3; ++j) {
				float dv = fShift * (v[3 * p + j] - vcm[j]) + fCosmo * (r[3 * p + j] - rcm[j]);
				dv2 += dv * dv;
			}
			fTot = 0.5 * dv2 - pot[p] * (1.0 + z);
			if (fTot > fTotBig) {
				fTotBig = fTot;
				iBig = i;
			}
		}
		if (fTotBig < 0 || bNoUnbind) break;
		{
			int p = idx[iBig];
			removed[p] = 1;
			++nRemoved;
			--m;
			if (m == 0) {
				dMass = 0.0;
				for (j = 0; j < 3; ++j) vcm[j] = 0.0;
				break;
			}
			dMass -= mass[p];
			for (j = 0; j < 3; ++j) {
				rcm[j] += mass[p] / dMass * (rcm[j] - r[3 * p + j]);
				vcm[j] += mass[p] / dMass * (vcm[j] - v[3 * p + j]);
			}
			idx[iBig] = idx[m];
			idx[m] = p;
			if (bSubPot) /* kdSubPot grav.c:39-60 */
				for (i = 0; i < m; ++i) {
					int q = idx[i];
					float dx = r[3 * p] - r[3 * q], dy = r[3 * p + 1] - r[3 * q + 1], dz = r[3 * p + 2] - r[3 * q + 2];
					float dir = soft_dir(dist2f(dx, dy, dz), soft[p] + soft[q], iSoftType);
					pot[q] -= G * mass[p] * dir;
				}
		}
	}
	*boundMass = dMass;
	for (j = 0; j < 3; ++j) vcmOut[j] = vcm[j];
	free(pot);
	free(idx);
	return nRemoved;
}

/* ---- kdOutStats (kd.c:1703-1839) ----------------------------------------------------------------
 * Per group: members sorted by squared distance from rCenter (CmpRadius kd.c:1690-1700), then the
 * reference's sequential float32 accumulation, expression by expression (the literals 0.5, 4.0, 3.0
 * are doubles there, so those sub-expressions are evaluated in double). */
typedef struct {
	float rad2;
	float rel[3];
	int idx;
} stat_member;

static int cmp_stat_member(const void *a, const void *b)
{
	float x = ((const stat_member *)a)->rad2, y = ((const stat_member *)b)->rad2;
	return (x > y) - (x < y);
}

void orc_stats(int n, const float *pos, const float *vel, const float *mass, const float *soft, const float *temp,
               const float *rho, int nGas, int nDark, const int *piGroup, int nGroup, const float *rCenter,
               const float *vcm, const float *period, float G, float z, double dExpHub, float fDensMin,
               float fTempMax, orc_stat_row *rows)
{
	int *start = (int *)calloc((size_t)nGroup + 1, sizeof(int));
	int *fill = (int *)calloc((size_t)nGroup + 1, sizeof(int));
	stat_member *all = (stat_member *)malloc((size_t)(n > 0 ? n : 1) * sizeof(stat_member));
	const float fExp = 1.0 / (1.0 + z);
	const float fExpHub = dExpHub;
	float half[3];
	int i, k, ig;
	for (k = 0; k < 3; ++k) half[k] = 0.5 * period[k];
	memset(rows, 0, (size_t)nGroup * sizeof(orc_stat_row));
	for (i = 0; i < n; ++i) start[piGroup[i] + 1]++;
	for (ig = 0; ig < nGroup; ++ig) start[ig + 1] += start[ig];
	for (i = 0; i < n; ++i) {
		ig = piGroup[i];
		all[start[ig] + fill[ig]++].idx = i;
	}
	for (ig = 1; ig < nGroup; ++ig) {
		stat_member *q = all + start[ig];
		const int nm = start[ig + 1] - start[ig];
		float fTotMass = 0.0, fGasMass = 0.0, fStarMass = 0.0, fHalfMass = 0.0;
		float fVcirc = 0.0, fmVcirc = 0.0, flVcirc, fVdisp = 0.0, fRVmax = 0.0, fRhmass = 0.0;
		int j;
		if (nm <= 0) continue;
		for (j = 0; j < nm; ++j) {
			const int p = q[j].idx;
			float r2 = 0.0;
			for (k = 0; k < 3; ++k) {
				float d = pos[3 * p + k] - rCenter[3 * ig + k];
				if (d > half[k]) d -= 2 * half[k];
				if (d <= -half[k]) d += 2 * half[k];
				q[j].rel[k] = d;
			}
			for (k = 0; k < 3; ++k) r2 += q[j].rel[k] * q[j].rel[k];
			q[j].rad2 = r2;
		}
		qsort(q, (size_t)nm, sizeof(stat_member), cmp_stat_member);
		for (j = 0; j < nm; ++j) fHalfMass += 0.5 * mass[q[j].idx];
		for (j = 0; j < nm; ++j) {
			const int p = q[j].idx;
			fTotMass += mass[p];
			if (q[j].rad2 > 4.0 * soft[p] * soft[p] && G * fTotMass / sqrt(q[j].rad2) > fVcirc) {
				fRVmax = sqrt(q[j].rad2);
				fVcirc = G * fTotMass / fRVmax;
			}
			if (p < nGas && rho[p] >= fDensMin && temp[p] <= fTempMax) fGasMass += mass[p];
			if (p >= nGas + nDark) fStarMass += mass[p];
			if (fTotMass > fHalfMass && fmVcirc == 0.0) {
				fRhmass = sqrt(q[j].rad2);
				fmVcirc = G * fTotMass / fRhmass;
			}
			for (k = 0; k < 3; ++k) {
				float dv = fExp * (vel[3 * p + k] - vcm[3 * ig + k]) + fExpHub * q[j].rel[k];
				fVdisp += dv * dv;
			}
		}
		flVcirc = G * fTotMass / sqrt(q[nm - 1].rad2);
		if (fVcirc == 0.0) {
			fVcirc = flVcirc;
			fRVmax = sqrt(q[nm - 1].rad2);
		}
		rows[ig].nMembers = nm;
		rows[ig].fTotMass = fTotMass;
		rows[ig].fGasMass = fGasMass;
		rows[ig].fStarMass = fStarMass;
		rows[ig].fVcirc = fVcirc;
		rows[ig].fmVcirc = fmVcirc;
		rows[ig].flVcirc = flVcirc;
		rows[ig].fRVmax = fRVmax;
		rows[ig].fRhmass = fRhmass;
		rows[ig].fRouter2 = q[nm - 1].rad2;
		rows[ig].fVdispSum = fVdisp;
	}
	free(all);
	free(start);
	free(fill);
}
