/* TIPSY input for the host driver: replaces kdReadTipsy (kd.c:122-222), kdInGroup (kd.c:920-962) and
 * the file branch of kdReadCenter (kd.c:1083-1159).  Formats: reference tipsydefs.h:6-48; the XDR
 * "standard" variant is big-endian words with one explicit pad int after the header (kd.c:16-28). */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "skid_host.h"

#define GAS_F 12
#define DARK_F 9
#define STAR_F 11

static uint32_t bswap32(uint32_t v)
{
	return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
}

static int read_exact(FILE *fp, void *buf, size_t nbytes)
{
	size_t got = 0;
	while (got < nbytes) {
		size_t r = fread((char *)buf + got, 1, nbytes - got, fp);
		if (r == 0) return -1;
		got += r;
	}
	return 0;
}

typedef struct {
	double time;
	int nbodies, ndim, nsph, ndark, nstar, pad;
} tipsy_header;

static int read_header(FILE *fp, int bStandard, tipsy_header *h)
{
	unsigned char raw[32];
	if (read_exact(fp, raw, 32)) return -1;
	if (bStandard) {
		/* big-endian double then six big-endian ints */
		unsigned char t[8];
		int i;
		uint32_t w[6];
		for (i = 0; i < 8; ++i) t[i] = raw[7 - i];
		memcpy(&h->time, t, 8);
		memcpy(w, raw + 8, 24);
		h->nbodies = (int)bswap32(w[0]);
		h->ndim = (int)bswap32(w[1]);
		h->nsph = (int)bswap32(w[2]);
		h->ndark = (int)bswap32(w[3]);
		h->nstar = (int)bswap32(w[4]);
		h->pad = 0;
	} else {
		memcpy(h, raw, 32);
	}
	return 0;
}

/* read n records of nf floats each into a freshly malloc'd host-endian buffer */
static float *read_records(FILE *fp, int bStandard, size_t n, int nf)
{
	size_t cnt = n * (size_t)nf, i;
	float *a = (float *)malloc((cnt ? cnt : 1) * sizeof(float));
	if (!a) return NULL;
	if (read_exact(fp, a, cnt * sizeof(float))) {
		free(a);
		return NULL;
	}
	if (bStandard) {
		uint32_t *w = (uint32_t *)a;
		for (i = 0; i < cnt; ++i) w[i] = bswap32(w[i]);
	}
	return a;
}

int tipsy_read(FILE *fp, int bStandard, snapshot *s)
{
	tipsy_header h;
	float *a;
	int i, j, base;
	if (read_header(fp, bStandard, &h)) return -1;
	if (h.nsph < 0 || h.ndark < 0 || h.nstar < 0) return -1;
	s->time = h.time;
	s->nGas = h.nsph;
	s->nDark = h.ndark;
	s->nStar = h.nstar;
	s->n = h.nsph + h.ndark + h.nstar;
	s->p = (skidgpu_pinit *)calloc((size_t)(s->n ? s->n : 1), sizeof(skidgpu_pinit));
	if (!s->p) return -1;
	/* gas: mass pos3 vel3 rho temp hsmooth metals phi */
	a = read_records(fp, bStandard, (size_t)h.nsph, GAS_F);
	if (!a) return -1;
	for (i = 0; i < h.nsph; ++i) {
		const float *r = a + (size_t)i * GAS_F;
		skidgpu_pinit *q = &s->p[i];
		q->fMass = r[0];
		for (j = 0; j < 3; ++j) {
			q->r[j] = r[1 + j];
			q->v[j] = r[4 + j];
		}
		q->fTemp = r[8];
		q->fSoft = r[9];
	}
	free(a);
	/* dark: mass pos3 vel3 eps phi */
	base = h.nsph;
	a = read_records(fp, bStandard, (size_t)h.ndark, DARK_F);
	if (!a) return -1;
	for (i = 0; i < h.ndark; ++i) {
		const float *r = a + (size_t)i * DARK_F;
		skidgpu_pinit *q = &s->p[base + i];
		q->fMass = r[0];
		for (j = 0; j < 3; ++j) {
			q->r[j] = r[1 + j];
			q->v[j] = r[4 + j];
		}
		q->fSoft = r[7];
	}
	free(a);
	/* star: mass pos3 vel3 metals tform eps phi */
	base += h.ndark;
	a = read_records(fp, bStandard, (size_t)h.nstar, STAR_F);
	if (!a) return -1;
	for (i = 0; i < h.nstar; ++i) {
		const float *r = a + (size_t)i * STAR_F;
		skidgpu_pinit *q = &s->p[base + i];
		q->fMass = r[0];
		for (j = 0; j < 3; ++j) {
			q->r[j] = r[1 + j];
			q->v[j] = r[4 + j];
		}
		q->fSoft = r[9];
	}
	free(a);
	for (i = 0; i < s->n; ++i) s->p[i].iOrder = i;
	return 0;
}

int grp_read(const char *path, int n, int *piGroup)
{
	FILE *fp = fopen(path, "r");
	int nf, i, g, ng = 0;
	if (!fp) {
		fprintf(stderr, "ERROR: Could not open file:%s\n", path);
		return -1;
	}
	if (fscanf(fp, "%d", &nf) != 1 || nf != n) {
		fprintf(stderr, "ERROR: Mismatched number of particles\n");
		fprintf(stderr, "Number in Group file %s: %d\n", path, nf);
		fprintf(stderr, "Number in TIPSY BINARY input file: %d\n", n);
		fclose(fp);
		return -1;
	}
	for (i = 0; i < n; ++i) {
		if (fscanf(fp, "%d", &g) != 1) g = 0;
		piGroup[i] = g;
		if (g > ng) ng = g;
	}
	fclose(fp);
	return ng + 1;
}

int gtp_read(const char *path, int bStandard, int nGroup, skidgpu_pgroup *g)
{
	FILE *fp = fopen(path, "rb");
	tipsy_header h;
	float *a;
	int i, j;
	if (!fp) return 0;
	if (read_header(fp, bStandard, &h)) {
		fclose(fp);
		return -1;
	}
	if (h.nstar != nGroup - 1) {
		fprintf(stderr, "ERROR: grp and gtp files don't match: %d vs. %d.\n", h.nstar, nGroup - 1);
		fclose(fp);
		return -1;
	}
	/* skip any gas / dark records (kd.c:1121-1128) */
	fseek(fp, (long)h.nsph * GAS_F * 4 + (long)h.ndark * DARK_F * 4, SEEK_CUR);
	a = read_records(fp, bStandard, (size_t)h.nstar, STAR_F);
	fclose(fp);
	if (!a) return -1;
	for (i = 0; i < h.nstar; ++i) {
		const float *r = a + (size_t)i * STAR_F;
		for (j = 0; j < 3; ++j) {
			g[i + 1].rCenter[j] = r[1 + j];
			g[i + 1].vcm[j] = r[4 + j];
		}
	}
	free(a);
	return 1;
}
