/* Output side of the host driver.  Byte formats follow the reference writers:
 * kdOutGroup (kd.c:1502-1523), kdOutDensity (1526-1547), kdOutVector (1550-1608),
 * kdWriteGroup (1611-1687), kdOutStats (1703-1839). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "skid_host.h"

/* The three ASCII arrays are formatted in chunks on all host threads (fastio.c); the bytes are the
 * ones fprintf("%d\n") / ("%.10g\n") / ("%g\n") would produce. */
static char *fmt_group_chunk(void *arg, size_t lo, size_t hi, char *out)
{
	const int *piGroup = (const int *)arg;
	size_t i;
	for (i = lo; i < hi; ++i) {
		out = fmt_int(out, piGroup[i]);
		*out++ = '\n';
	}
	return out;
}

int out_group(const char *path, int n, const int *piGroup)
{
	FILE *fp = fopen(path, "w");
	int rc;
	if (!fp) return -1;
	fprintf(fp, "%d\n", n);
	rc = chunked_write(fp, (size_t)(n > 0 ? n : 0), 12, fmt_group_chunk, (void *)piGroup);
	return fclose(fp) | rc;
}

static char *fmt_density_chunk(void *arg, size_t lo, size_t hi, char *out)
{
	const float *rho = (const float *)arg;
	size_t i;
	for (i = lo; i < hi; ++i) {
		out = fmt_g(out, rho[i], 10);
		*out++ = '\n';
	}
	return out;
}

int out_density(const char *path, int n, const float *rho)
{
	FILE *fp = fopen(path, "w");
	int rc;
	if (!fp) return -1;
	fprintf(fp, "%d\n", n);
	rc = chunked_write(fp, (size_t)(n > 0 ? n : 0), 40, fmt_density_chunk, (void *)rho);
	return fclose(fp) | rc;
}

/* .ray: N, then every x displacement, every y, every z; movers get the min-image of
 * (final - initial), everything else a literal 0. */
typedef struct {
	const snapshot *s;
	const int *slot; /* mover slot of particle i or -1 */
	const float *r3;
	float half;
	int axis;
} ray_state;

static char *fmt_ray_chunk(void *arg, size_t lo, size_t hi, char *out)
{
	const ray_state *st = (const ray_state *)arg;
	size_t i;
	for (i = lo; i < hi; ++i) {
		const int m = st->slot[i];
		if (m >= 0) {
			float d = st->r3[3 * (size_t)m + st->axis] - st->s->p[i].r[st->axis];
			if (d > st->half) d -= 2 * st->half;
			if (d <= -st->half) d += 2 * st->half;
			out = fmt_g(out, d, 6);
		} else {
			*out++ = '0';
		}
		*out++ = '\n';
	}
	return out;
}

int out_vector(const char *path, const snapshot *s, int nMove, const int *iOrder, const float *r3,
               const float fPeriod[3])
{
	FILE *fp = fopen(path, "w");
	ray_state st;
	int *slot;
	int axis, i, m, rc = 0;
	if (!fp) return -1;
	fprintf(fp, "%d\n", s->n);
	/* the reference walks the movers in ascending iOrder next to the particles (kd.c:1577-1590):
	 * a mover whose iOrder is out of sequence would be skipped there, so it is skipped here */
	slot = (int *)malloc((size_t)(s->n ? s->n : 1) * sizeof(int));
	if (!slot) {
		fclose(fp);
		return -1;
	}
	m = 0;
	for (i = 0; i < s->n; ++i) {
		if (m < nMove && iOrder[m] == i) slot[i] = m++;
		else slot[i] = -1;
	}
	st.s = s;
	st.slot = slot;
	st.r3 = r3;
	for (axis = 0; axis < 3 && !rc; ++axis) {
		st.half = 0.5 * fPeriod[axis];
		st.axis = axis;
		rc = chunked_write(fp, (size_t)s->n, 40, fmt_ray_chunk, &st);
	}
	free(slot);
	return fclose(fp) | rc;
}

static void put_be32(FILE *fp, const void *v)
{
	const unsigned char *b = (const unsigned char *)v;
	unsigned char o[4] = {b[3], b[2], b[1], b[0]};
	fwrite(o, 1, 4, fp);
}

int out_gtp(const char *path, int bStandard, double fTime, int nGroup, const skidgpu_pgroup *g)
{
	FILE *fp = fopen(path, "wb");
	int i, j, ng = nGroup - 1;
	if (!fp) return -1;
	if (bStandard) {
		const unsigned char *t = (const unsigned char *)&fTime;
		unsigned char o[8];
		int hdr[6];
		for (i = 0; i < 8; ++i) o[i] = t[7 - i];
		fwrite(o, 1, 8, fp);
		hdr[0] = ng; /* nbodies */
		hdr[1] = 3;  /* ndim */
		hdr[2] = 0;  /* nsph */
		hdr[3] = 0;  /* ndark */
		hdr[4] = ng; /* nstar */
		hdr[5] = 0;  /* pad */
		for (i = 0; i < 6; ++i) put_be32(fp, &hdr[i]);
	} else {
		struct {
			double time;
			int nbodies, ndim, nsph, ndark, nstar, pad;
		} h;
		memset(&h, 0, sizeof h);
		h.time = fTime;
		h.nbodies = ng;
		h.ndim = 3;
		h.nstar = ng;
		fwrite(&h, 32, 1, fp);
	}
	for (i = 1; i < nGroup; ++i) {
		float rec[11];
		rec[0] = g[i].fMass;
		for (j = 0; j < 3; ++j) {
			rec[1 + j] = g[i].rCenter[j];
			rec[4 + j] = g[i].vcm[j];
		}
		rec[7] = 0.0f;         /* metals */
		rec[8] = (float)fTime; /* tform */
		rec[9] = g[i].fRadius; /* eps = group radius */
		rec[10] = 0.0f;        /* phi */
		if (bStandard)
			for (j = 0; j < 11; ++j) put_be32(fp, &rec[j]);
		else
			fwrite(rec, 4, 11, fp);
	}
	return fclose(fp);
}

/* ---- .stat ---------------------------------------------------------------------------- */
typedef struct {
	float rad2; /* squared distance from the group centre */
	float rel[3];
	int idx; /* file index */
} member;

static int cmp_member(const void *a, const void *b)
{
	float x = ((const member *)a)->rad2, y = ((const member *)b)->rad2;
	return (x > y) - (x < y);
}

static int species(const snapshot *s, int i)
{
	if (i < s->nGas) return SKIDGPU_GAS;
	if (i < s->nGas + s->nDark) return SKIDGPU_DARK;
	return SKIDGPU_STAR;
}

int out_stats(const char *path, const snapshot *s, const float *rho, const int *piGroup, int nGroup,
              const skidgpu_pgroup *g, const float fPeriod[3], float G, float z, double dExpHub,
              float fDensMin, float fTempMax)
{
	FILE *fp = fopen(path, "w");
	int *start, *fill;
	member *all;
	int i, k, ig;
	const float fExp = 1.0 / (1.0 + z);
	const float fExpHub = dExpHub;
	float half[3];
	if (!fp) return -1;
	for (k = 0; k < 3; ++k) half[k] = 0.5 * fPeriod[k];
	/* bucket the members of every group (counting sort by group id) */
	start = (int *)calloc((size_t)nGroup + 1, sizeof(int));
	fill = (int *)calloc((size_t)nGroup + 1, sizeof(int));
	for (i = 0; i < s->n; ++i) start[piGroup[i] + 1]++;
	for (ig = 0; ig < nGroup; ++ig) start[ig + 1] += start[ig];
	all = (member *)malloc((size_t)(s->n > 0 ? s->n : 1) * sizeof(member));
	for (i = 0; i < s->n; ++i) {
		ig = piGroup[i];
		all[start[ig] + fill[ig]++].idx = i;
	}
	for (ig = 1; ig < nGroup; ++ig) {
		member *q = all + start[ig];
		const int n = start[ig + 1] - start[ig];
		float fTotMass = 0.0, fGasMass = 0.0, fStarMass = 0.0, fHalfMass = 0.0;
		float fVcirc = 0.0, fmVcirc = 0.0, flVcirc, fVdisp = 0.0, fRVmax = 0.0, fRhmass = 0.0;
		int j;
		if (n <= 0) continue;
		for (j = 0; j < n; ++j) {
			const skidgpu_pinit *p = &s->p[q[j].idx];
			float r2 = 0.0;
			for (k = 0; k < 3; ++k) {
				float d = p->r[k] - g[ig].rCenter[k];
				if (d > half[k]) d -= 2 * half[k];
				if (d <= -half[k]) d += 2 * half[k];
				q[j].rel[k] = d;
			}
			for (k = 0; k < 3; ++k) r2 += q[j].rel[k] * q[j].rel[k];
			q[j].rad2 = r2;
		}
		qsort(q, (size_t)n, sizeof(member), cmp_member);
		for (j = 0; j < n; ++j) fHalfMass += 0.5 * s->p[q[j].idx].fMass;
		for (j = 0; j < n; ++j) {
			const skidgpu_pinit *p = &s->p[q[j].idx];
			const int sp = species(s, q[j].idx);
			fTotMass += p->fMass;
			if (q[j].rad2 > 4.0 * p->fSoft * p->fSoft && G * fTotMass / sqrt(q[j].rad2) > fVcirc) {
				fRVmax = sqrt(q[j].rad2);
				fVcirc = G * fTotMass / fRVmax;
			}
			if (sp == SKIDGPU_GAS && rho[q[j].idx] >= fDensMin && p->fTemp <= fTempMax) fGasMass += p->fMass;
			if (sp == SKIDGPU_STAR) fStarMass += p->fMass;
			if (fTotMass > fHalfMass && fmVcirc == 0.0) {
				fRhmass = sqrt(q[j].rad2);
				fmVcirc = G * fTotMass / fRhmass;
			}
			for (k = 0; k < 3; ++k) {
				float dv = fExp * (p->v[k] - g[ig].vcm[k]) + fExpHub * q[j].rel[k];
				fVdisp += dv * dv;
			}
		}
		flVcirc = G * fTotMass / sqrt(q[n - 1].rad2);
		if (fVcirc == 0.0) {
			fVcirc = flVcirc;
			fRVmax = sqrt(q[n - 1].rad2);
		}
		fVdisp = sqrt(fVdisp / (3.0 * n));
		fprintf(fp, "%d %d %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g\n", ig, n, fTotMass, fGasMass,
		        fStarMass, sqrt(fVcirc), sqrt(fmVcirc), sqrt(flVcirc), fRVmax, fRhmass, sqrt(q[n - 1].rad2), fVdisp,
		        g[ig].rCenter[0], g[ig].rCenter[1], g[ig].rCenter[2], g[ig].vcm[0], g[ig].vcm[1], g[ig].vcm[2],
		        g[ig].rBound[0], g[ig].rBound[1], g[ig].rBound[2]);
	}
	free(all);
	free(start);
	free(fill);
	return fclose(fp);
}
