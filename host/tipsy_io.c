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

/* read n records of nf floats each into a freshly malloc'd buffer, still in FILE byte order (the
 * swap happens in the parallel scatter below, or in fix_endian for small catalogue reads) */
static float *read_records(FILE *fp, size_t n, int nf)
{
	size_t cnt = n * (size_t)nf;
	float *a = (float *)malloc((cnt ? cnt : 1) * sizeof(float));
	if (!a) return NULL;
	if (read_exact(fp, a, cnt * sizeof(float))) {
		free(a);
		return NULL;
	}
	return a;
}

static void fix_endian(float *a, size_t cnt, int bStandard)
{
	uint32_t *w = (uint32_t *)a;
	size_t i;
	if (bStandard)
		for (i = 0; i < cnt; ++i) w[i] = bswap32(w[i]);
}

/* one species block: records of nf floats -> PINIT entries [base, base+n); field positions differ
 * per species (tipsydefs.h:6-37): mass 0, pos 1-3, vel 4-6 always; temp/soft columns given */
typedef struct {
	const float *rec;
	skidgpu_pinit *p;
	int nf, base, iTemp, iSoft, bSwap;
} scatter_job;

static float ld(const float *r, int i, int bSwap)
{
	uint32_t w;
	float f;
	memcpy(&w, r + i, 4);
	if (bSwap) w = bswap32(w);
	memcpy(&f, &w, 4);
	return f;
}

static void scatter_range(void *arg, size_t lo, size_t hi, int tid)
{
	const scatter_job *j = (const scatter_job *)arg;
	size_t i;
	int k;
	(void)tid;
	for (i = lo; i < hi; ++i) {
		const float *r = j->rec + i * (size_t)j->nf;
		skidgpu_pinit *q = &j->p[(size_t)j->base + i];
		memset(q, 0, sizeof *q);
		q->fMass = ld(r, 0, j->bSwap);
		for (k = 0; k < 3; ++k) {
			q->r[k] = ld(r, 1 + k, j->bSwap);
			q->v[k] = ld(r, 4 + k, j->bSwap);
		}
		if (j->iTemp >= 0) q->fTemp = ld(r, j->iTemp, j->bSwap);
		q->fSoft = ld(r, j->iSoft, j->bSwap);
		q->iOrder = j->base + (int)i;
	}
}

static int read_species(FILE *fp, int bStandard, snapshot *s, int base, int n, int nf, int iTemp, int iSoft)
{
	scatter_job j;
	float *a = read_records(fp, (size_t)n, nf);
	if (!a) return -1;
	j.rec = a;
	j.p = s->p;
	j.nf = nf;
	j.base = base;
	j.iTemp = iTemp;
	j.iSoft = iSoft;
	j.bSwap = bStandard;
	par_for((size_t)n, 1u << 16, scatter_range, &j);
	free(a);
	return 0;
}

int tipsy_read(FILE *fp, int bStandard, snapshot *s)
{
	tipsy_header h;
	if (read_header(fp, bStandard, &h)) return -1;
	if (h.nsph < 0 || h.ndark < 0 || h.nstar < 0) return -1;
	if ((long long)h.nsph + h.ndark + h.nstar > 2147483647LL) return -1;
	s->time = h.time;
	s->nGas = h.nsph;
	s->nDark = h.ndark;
	s->nStar = h.nstar;
	s->n = h.nsph + h.ndark + h.nstar;
	s->p = (skidgpu_pinit *)malloc((size_t)(s->n ? s->n : 1) * sizeof(skidgpu_pinit));
	if (!s->p) return -1;
	/* gas: mass pos3 vel3 rho temp hsmooth metals phi; dark: mass pos3 vel3 eps phi;
	 * star: mass pos3 vel3 metals tform eps phi */
	if (read_species(fp, bStandard, s, 0, h.nsph, GAS_F, 8, 9)) return -1;
	if (read_species(fp, bStandard, s, h.nsph, h.ndark, DARK_F, -1, 7)) return -1;
	if (read_species(fp, bStandard, s, h.nsph + h.ndark, h.nstar, STAR_F, -1, 9)) return -1;
	return 0;
}

/* kdInGroup: the whole file is read at once and parsed without stdio; like the reference's loop of
 * fscanf("%d") a token that is not an integer ends the conversion and the rest reads as group 0 */
static char *slurp(FILE *fp, size_t *len)
{
	size_t cap = 1u << 20, n = 0;
	char *buf = (char *)malloc(cap);
	while (buf) {
		size_t r = fread(buf + n, 1, cap - n, fp);
		n += r;
		if (r == 0) break;
		if (n == cap) {
			char *nb = (char *)realloc(buf, cap *= 2);
			if (!nb) free(buf);
			buf = nb;
		}
	}
	*len = n;
	return buf;
}

int grp_read(const char *path, int n, int *piGroup)
{
	FILE *fp = fopen(path, "r");
	char *text;
	size_t len, got;
	int nf = 0, i, ng = 0, *all;
	if (!fp) {
		fprintf(stderr, "ERROR: Could not open file:%s\n", path);
		return -1;
	}
	text = slurp(fp, &len);
	fclose(fp);
	all = (int *)malloc(((size_t)(n > 0 ? n : 0) + 1) * sizeof(int));
	if (!text || !all) {
		free(text);
		free(all);
		return -1;
	}
	got = parse_ints(text, len, all, (size_t)(n > 0 ? n : 0) + 1);
	free(text);
	if (got) nf = all[0];
	if (got < 1 || nf != n) {
		fprintf(stderr, "ERROR: Mismatched number of particles\n");
		fprintf(stderr, "Number in Group file %s: %d\n", path, nf);
		fprintf(stderr, "Number in TIPSY BINARY input file: %d\n", n);
		free(all);
		return -1;
	}
	for (i = 0; i < n; ++i) {
		const int g = (size_t)i + 1 < got ? all[i + 1] : 0;
		piGroup[i] = g;
		if (g > ng) ng = g;
	}
	free(all);
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
	a = read_records(fp, (size_t)h.nstar, STAR_F);
	fclose(fp);
	if (!a) return -1;
	fix_endian(a, (size_t)h.nstar * STAR_F, bStandard);
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
