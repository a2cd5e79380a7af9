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

/* ---- .stat ----------------------------------------------------------------------------
 * kdOutStats' print statement (kd.c:1822-1836) over the accumulator rows computed on the device by
 * skidgpu_stats (the per-group radial sort and sequential sums are GPU work, see csrc/stats.cu). */
int out_stats(const char *path, int nGroup, const skidgpu_pgroup *g, const skidgpu_stat_row *row)
{
	FILE *fp = fopen(path, "w");
	int ig;
	if (!fp) return -1;
	for (ig = 1; ig < nGroup; ++ig) {
		const skidgpu_stat_row *r = &row[ig];
		const int n = r->nMembers;
		float fVdisp;
		if (n <= 0) continue;
		fVdisp = sqrt(r->fVdispSum / (3.0 * n));
		fprintf(fp, "%d %d %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g %g\n", ig, n, r->fTotMass, r->fGasMass,
		        r->fStarMass, sqrt(r->fVcirc), sqrt(r->fmVcirc), sqrt(r->flVcirc), r->fRVmax, r->fRhmass,
		        sqrt(r->fRouter2), fVdisp, g[ig].rCenter[0], g[ig].rCenter[1], g[ig].rCenter[2], g[ig].vcm[0],
		        g[ig].vcm[1], g[ig].vcm[2], g[ig].rBound[0], g[ig].rBound[1], g[ig].rBound[2]);
	}
	return fclose(fp);
}
