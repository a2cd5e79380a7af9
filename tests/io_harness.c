/*
 * TEST / BENCH harness for the host-side I/O (host/fastio.c, outputs.c, tipsy_io.c); no GPU, no
 * libskidgpu.  The checker in every mode is the reference's own way of doing it: one
 * fprintf("%d\n" | "%.10g\n" | "%g\n") per value (kd.c:1518-1519, 1542-1545, 1570-1606) and a
 * word-at-a-time XDR decode (kd.c:141-206).
 *
 *   io_harness fmt <count> <seed>     random + adversarial floats/ints: fmt_g/fmt_int vs snprintf
 *   io_harness exh <lo> <hi>          EVERY float bit pattern in [lo, hi) (hex): fmt_g vs snprintf for %g and %.10g
 *   io_harness files <n> <seed> <dir> writes .grp/.den/.ray with the fast writers and with fprintf,
 *                                     compares the bytes, prints both timings as one JSON line
 *   io_harness emit <n> <dir>         <dir>/grp.i32, den.f32, ray.f32 (n x 3 displacements) -> <dir>/out.grp,
 *                                     out.den, out.ray through the product writers (golden md5 test)
 *   io_harness read <n> <seed> <dir>  writes a TIPSY -std file + .grp, reads them with the fast
 *                                     readers and with the word-at-a-time / fscanf way, compares
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "skid_host.h"

static uint64_t g_rng = 88172645463325252ull;
static uint64_t rnd(void)
{
	g_rng ^= g_rng << 13;
	g_rng ^= g_rng >> 7;
	g_rng ^= g_rng << 17;
	return g_rng;
}
static double now(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}
static float bits2f(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}

static long g_bad = 0, g_checked = 0;
static void check_g(float f)
{
	char a[64], b[64];
	int prec;
	for (prec = 6; prec <= 10; prec += 4) {
		char *e = fmt_g(a, f, prec);
		*e = 0;
		snprintf(b, sizeof b, "%.*g", prec, f);
		++g_checked;
		if (strcmp(a, b)) {
			if (g_bad++ < 20) fprintf(stderr, "MISMATCH prec %d bits %a: fast '%s' printf '%s'\n", prec, f, a, b);
		}
	}
}
static void check_d(int v)
{
	char a[32], b[32];
	*fmt_int(a, v) = 0;
	snprintf(b, sizeof b, "%d", v);
	++g_checked;
	if (strcmp(a, b) && g_bad++ < 20) fprintf(stderr, "MISMATCH int %d: '%s' vs '%s'\n", v, a, b);
}

static int mode_fmt(long count, uint64_t seed)
{
	long i;
	int e, d;
	g_rng ^= seed * 0x9E3779B97F4A7C15ull;
	/* every bit pattern class: uniform random bits (all exponents, inf, nan, denormals) */
	for (i = 0; i < count; ++i) check_g(bits2f((uint32_t)rnd()));
	/* values in the ranges the writers see: densities, displacements, coordinates */
	for (i = 0; i < count; ++i) {
		double u = (double)(rnd() >> 11) / 9007199254740992.0;
		check_g((float)(u - 0.5));
		check_g((float)(exp(40.0 * u - 20.0)));
		check_g((float)(1e-4 * (u - 0.5)));
	}
	/* ties and decade boundaries: k/2^j around powers of ten, exact integers and half-integers */
	for (e = -12; e <= 12; ++e)
		for (d = -3000; d <= 3000; ++d) {
			const float base = (float)pow(10.0, e);
			uint32_t ub;
			memcpy(&ub, &base, 4);
			check_g(bits2f((uint32_t)((int32_t)ub + d)));
			check_g(-bits2f((uint32_t)((int32_t)ub + d)));
		}
	for (i = 0; i < 4000000; ++i) {
		check_g((float)i * 0.5f);
		check_g((float)i + 100000.5f);
		check_g((float)i * 0.03125f);
		check_g((float)(999990 + (i & 1023)) * 0.0009765625f * (float)(1 << (i % 20)));
	}
	check_g(0.0f);
	check_g(-0.0f);
	check_g(bits2f(1));
	check_g(bits2f(0x7f7fffff));
	check_g(bits2f(0x7f800000));
	check_g(bits2f(0xff800000));
	check_g(bits2f(0x7fc00000));
	for (i = 0; i < count; ++i) check_d((int)(uint32_t)rnd());
	for (i = -100000; i <= 100000; ++i) check_d((int)i);
	check_d(2147483647);
	check_d(-2147483647 - 1);
	printf("{\"mode\": \"fmt\", \"checked\": %ld, \"mismatches\": %ld}\n", g_checked, g_bad);
	return g_bad != 0;
}

static int same_file(const char *a, const char *b)
{
	FILE *fa = fopen(a, "rb"), *fb = fopen(b, "rb");
	static char ba[1 << 16], bb[1 << 16];
	int same = fa && fb;
	while (same) {
		size_t ra = fread(ba, 1, sizeof ba, fa), rb = fread(bb, 1, sizeof bb, fb);
		if (ra != rb || memcmp(ba, bb, ra)) same = 0;
		if (ra == 0) break;
	}
	if (fa) fclose(fa);
	if (fb) fclose(fb);
	return same;
}

/* a synthetic run result: labels, densities, movers (about half of the particles) */
static void make_run(int n, snapshot *s, int **piGroup, float **rho, int *nMove, int **iOrder, float **r3)
{
	int i, m = 0, k;
	s->n = n;
	s->nGas = 0;
	s->nDark = n;
	s->nStar = 0;
	s->time = 1.0;
	s->p = (skidgpu_pinit *)calloc((size_t)n, sizeof(skidgpu_pinit));
	*piGroup = (int *)malloc((size_t)n * sizeof(int));
	*rho = (float *)malloc((size_t)n * sizeof(float));
	*iOrder = (int *)malloc((size_t)n * sizeof(int));
	*r3 = (float *)malloc((size_t)n * 3 * sizeof(float));
	for (i = 0; i < n; ++i) {
		double u = (double)(rnd() >> 11) / 9007199254740992.0;
		for (k = 0; k < 3; ++k) {
			s->p[i].r[k] = (float)((double)(rnd() >> 11) / 9007199254740992.0 - 0.5);
			s->p[i].v[k] = (float)((double)(rnd() >> 11) / 9007199254740992.0 - 0.5);
		}
		s->p[i].fMass = 1.0f / n;
		s->p[i].fSoft = 1e-4f;
		s->p[i].iOrder = i;
		(*rho)[i] = (float)exp(12.0 * u - 2.0);
		(*piGroup)[i] = (rnd() & 3) ? 0 : (int)(rnd() % (uint64_t)(n / 64 + 2));
		if (rnd() & 1) {
			(*iOrder)[m] = i;
			for (k = 0; k < 3; ++k) {
				float d = (float)(0.02 * ((double)(rnd() >> 11) / 9007199254740992.0 - 0.5));
				float x = s->p[i].r[k] + d;
				if (x > 0.5f) x -= 1.0f;
				if (x <= -0.5f) x += 1.0f;
				(*r3)[3 * (size_t)m + k] = x;
			}
			++m;
		}
	}
	*nMove = m;
}

static int mode_files(int n, uint64_t seed, const char *dir)
{
	snapshot s;
	int *piGroup, *iOrder, nMove, i, axis, m, ok;
	float *rho, *r3;
	const float fPeriod[3] = {1.0f, 1.0f, 1.0f};
	char fa[3][512], fb[3][512];
	const char *ext[3] = {"grp", "den", "ray"};
	double t0, tf[3], tr[3];
	FILE *fp;
	g_rng ^= seed * 0x9E3779B97F4A7C15ull;
	make_run(n, &s, &piGroup, &rho, &nMove, &iOrder, &r3);
	for (i = 0; i < 3; ++i) {
		snprintf(fa[i], sizeof fa[i], "%s/fast.%s", dir, ext[i]);
		snprintf(fb[i], sizeof fb[i], "%s/printf.%s", dir, ext[i]);
	}
	t0 = now();
	out_group(fa[0], n, piGroup);
	tf[0] = now() - t0;
	t0 = now();
	out_density(fa[1], n, rho);
	tf[1] = now() - t0;
	t0 = now();
	out_vector(fa[2], &s, nMove, iOrder, r3, fPeriod);
	tf[2] = now() - t0;
	/* the reference's way */
	t0 = now();
	fp = fopen(fb[0], "w");
	fprintf(fp, "%d\n", n);
	for (i = 0; i < n; ++i) fprintf(fp, "%d\n", piGroup[i]);
	fclose(fp);
	tr[0] = now() - t0;
	t0 = now();
	fp = fopen(fb[1], "w");
	fprintf(fp, "%d\n", n);
	for (i = 0; i < n; ++i) fprintf(fp, "%.10g\n", rho[i]);
	fclose(fp);
	tr[1] = now() - t0;
	t0 = now();
	fp = fopen(fb[2], "w");
	fprintf(fp, "%d\n", n);
	for (axis = 0; axis < 3; ++axis) {
		const float h = 0.5f * fPeriod[axis];
		for (i = 0, m = 0; i < n; ++i) {
			if (m < nMove && iOrder[m] == i) {
				float d = r3[3 * (size_t)m + axis] - s.p[i].r[axis];
				if (d > h) d -= 2 * h;
				if (d <= -h) d += 2 * h;
				fprintf(fp, "%g\n", d);
				++m;
			} else {
				fprintf(fp, "0\n");
			}
		}
	}
	fclose(fp);
	tr[2] = now() - t0;
	ok = 1;
	for (i = 0; i < 3; ++i) ok &= same_file(fa[i], fb[i]);
	printf("{\"mode\": \"files\", \"n\": %d, \"movers\": %d, \"threads\": %d, \"identical\": %s, "
	       "\"fast_s\": {\"grp\": %.4f, \"den\": %.4f, \"ray\": %.4f}, "
	       "\"fprintf_s\": {\"grp\": %.4f, \"den\": %.4f, \"ray\": %.4f}}\n",
	       n, nMove, host_threads(), ok ? "true" : "false", tf[0], tf[1], tf[2], tr[0], tr[1], tr[2]);
	for (i = 0; i < 3; ++i) {
		remove(fa[i]);
		remove(fb[i]);
	}
	return !ok;
}

static uint32_t be32(uint32_t v)
{
	return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
}
static void put_f(FILE *fp, float f)
{
	uint32_t w;
	memcpy(&w, &f, 4);
	w = be32(w);
	fwrite(&w, 4, 1, fp);
}
static float get_f(FILE *fp)
{
	uint32_t w = 0;
	float f;
	if (fread(&w, 4, 1, fp) != 1) return 0;
	w = be32(w);
	memcpy(&f, &w, 4);
	return f;
}

static int mode_read(int n, uint64_t seed, const char *dir)
{
	/* gas + dark + star so that every record layout is covered */
	const int nGas = n / 4, nStar = n / 8, nDark = n - nGas - nStar;
	char ft[512], fg[512];
	FILE *fp;
	snapshot s;
	skidgpu_pinit *ref;
	int *grp, *g1, *g2, i, k, ok = 1, ng1, ng2 = 0, nf;
	double t0, t_fast, t_ref, t_gfast, t_gref;
	unsigned char hdr[32];
	double tm = 0.75;
	uint32_t w[6];
	g_rng ^= seed * 0x9E3779B97F4A7C15ull;
	snprintf(ft, sizeof ft, "%s/in.std", dir);
	snprintf(fg, sizeof fg, "%s/in.grp", dir);
	fp = fopen(ft, "wb");
	for (i = 0; i < 8; ++i) hdr[i] = ((unsigned char *)&tm)[7 - i];
	w[0] = be32((uint32_t)n);
	w[1] = be32(3);
	w[2] = be32((uint32_t)nGas);
	w[3] = be32((uint32_t)nDark);
	w[4] = be32((uint32_t)nStar);
	w[5] = 0;
	memcpy(hdr + 8, w, 24);
	fwrite(hdr, 1, 32, fp);
	for (i = 0; i < nGas * 12 + nDark * 9 + nStar * 11; ++i) put_f(fp, (float)((double)(rnd() >> 11) / 9007199254740992.0 - 0.5));
	fclose(fp);
	grp = (int *)malloc((size_t)n * sizeof(int));
	fp = fopen(fg, "w");
	fprintf(fp, "%d\n", n);
	for (i = 0; i < n; ++i) {
		grp[i] = (rnd() & 1) ? 0 : (int)(rnd() % 100000);
		fprintf(fp, "%d\n", grp[i]);
	}
	fclose(fp);
	/* fast readers */
	fp = fopen(ft, "rb");
	t0 = now();
	if (tipsy_read(fp, 1, &s)) ok = 0;
	t_fast = now() - t0;
	fclose(fp);
	g1 = (int *)malloc((size_t)n * sizeof(int));
	g2 = (int *)malloc((size_t)n * sizeof(int));
	t0 = now();
	ng1 = grp_read(fg, n, g1);
	t_gfast = now() - t0;
	/* the reference's way: one XDR word at a time into PINIT (kd.c:141-206) */
	ref = (skidgpu_pinit *)calloc((size_t)n, sizeof(skidgpu_pinit));
	fp = fopen(ft, "rb");
	t0 = now();
	fseek(fp, 32, SEEK_SET);
	for (i = 0; i < n; ++i) {
		const int nfl = i < nGas ? 12 : i < nGas + nDark ? 9 : 11;
		float rec[12];
		for (k = 0; k < nfl; ++k) rec[k] = get_f(fp);
		ref[i].fMass = rec[0];
		for (k = 0; k < 3; ++k) {
			ref[i].r[k] = rec[1 + k];
			ref[i].v[k] = rec[4 + k];
		}
		if (i < nGas) {
			ref[i].fTemp = rec[8];
			ref[i].fSoft = rec[9];
		} else if (i < nGas + nDark) {
			ref[i].fSoft = rec[7];
		} else {
			ref[i].fSoft = rec[9];
		}
		ref[i].iOrder = i;
	}
	t_ref = now() - t0;
	fclose(fp);
	fp = fopen(fg, "r");
	t0 = now();
	if (fscanf(fp, "%d", &nf) != 1) nf = -1;
	for (i = 0; i < n; ++i) {
		int g = 0;
		if (fscanf(fp, "%d", &g) != 1) g = 0;
		g2[i] = g;
		if (g > ng2) ng2 = g;
	}
	++ng2;
	t_gref = now() - t0;
	fclose(fp);
	ok &= s.n == n && s.nGas == nGas && s.nDark == nDark && s.nStar == nStar && s.time == tm && nf == n;
	ok &= !memcmp(s.p, ref, (size_t)n * sizeof(skidgpu_pinit));
	ok &= ng1 == ng2 && !memcmp(g1, g2, (size_t)n * sizeof(int)) && !memcmp(g1, grp, (size_t)n * sizeof(int));
	/* kdInGroup's sanity check (kd.c:946-953): a .grp written for another particle count is refused;
	 * a token that is not an integer ends the conversion like fscanf("%d") does (rest = group 0) */
	{
		FILE *null_err = freopen("/dev/null", "w", stderr);
		(void)null_err;
		ok &= grp_read(fg, n + 1, g2) == -1;
		fp = fopen(fg, "w");
		fprintf(fp, "%d\n5\n 7 \nx 9\n", n);
		fclose(fp);
		if (n >= 4) {
			ok &= grp_read(fg, n, g2) == 8 && g2[0] == 5 && g2[1] == 7 && g2[2] == 0 && g2[n - 1] == 0;
		}
	}
	printf("{\"mode\": \"read\", \"n\": %d, \"threads\": %d, \"identical\": %s, \"fast_s\": {\"tipsy_std\": %.4f, \"grp\": %.4f}, "
	       "\"wordwise_s\": {\"tipsy_std\": %.4f, \"grp\": %.4f}}\n",
	       n, host_threads(), ok ? "true" : "false", t_fast, t_gfast, t_ref, t_gref);
	remove(ft);
	remove(fg);
	return !ok;
}

static void *load_raw(const char *dir, const char *name, size_t bytes)
{
	char path[512];
	void *buf = malloc(bytes ? bytes : 1);
	FILE *fp;
	snprintf(path, sizeof path, "%s/%s", dir, name);
	fp = fopen(path, "rb");
	if (!fp || fread(buf, 1, bytes, fp) != bytes) {
		fprintf(stderr, "cannot read %s\n", path);
		exit(2);
	}
	fclose(fp);
	return buf;
}

/* displacements d (n x 3, zero rows = not a mover) are fed to out_vector as r = 0, r_moved = d */
static int mode_emit(int n, const char *dir)
{
	int *grp = (int *)load_raw(dir, "grp.i32", (size_t)n * 4);
	float *den = (float *)load_raw(dir, "den.f32", (size_t)n * 4);
	float *ray = (float *)load_raw(dir, "ray.f32", (size_t)n * 12);
	const float fPeriod[3] = {1.0f, 1.0f, 1.0f};
	snapshot s;
	int *iOrder = (int *)malloc((size_t)n * sizeof(int));
	float *r3 = (float *)malloc((size_t)n * 12);
	char path[512];
	int i, m = 0, rc = 0;
	memset(&s, 0, sizeof s);
	s.n = s.nDark = n;
	s.p = (skidgpu_pinit *)calloc((size_t)n, sizeof(skidgpu_pinit));
	for (i = 0; i < n; ++i)
		if (ray[3 * i] != 0 || ray[3 * i + 1] != 0 || ray[3 * i + 2] != 0) {
			iOrder[m] = i;
			memcpy(r3 + 3 * (size_t)m, ray + 3 * (size_t)i, 12);
			++m;
		}
	snprintf(path, sizeof path, "%s/out.grp", dir);
	rc |= out_group(path, n, grp);
	snprintf(path, sizeof path, "%s/out.den", dir);
	rc |= out_density(path, n, den);
	snprintf(path, sizeof path, "%s/out.ray", dir);
	rc |= out_vector(path, &s, m, iOrder, r3, fPeriod);
	return rc != 0;
}

static int mode_exh(const char *slo, const char *shi)
{
	const uint64_t lo = strtoull(slo, NULL, 16), hi = strtoull(shi, NULL, 16);
	uint64_t u;
	for (u = lo; u < hi; ++u) check_g(bits2f((uint32_t)u));
	printf("{\"mode\": \"exh\", \"lo\": \"%llx\", \"hi\": \"%llx\", \"checked\": %ld, \"mismatches\": %ld}\n",
	       (unsigned long long)lo, (unsigned long long)hi, g_checked, g_bad);
	return g_bad != 0;
}

int main(int argc, char **argv)
{
	if (argc >= 4 && !strcmp(argv[1], "exh")) return mode_exh(argv[2], argv[3]);
	if (argc >= 4 && !strcmp(argv[1], "emit")) return mode_emit(atoi(argv[2]), argv[3]);
	if (argc >= 4 && !strcmp(argv[1], "fmt")) return mode_fmt(atol(argv[2]), (uint64_t)atol(argv[3]));
	if (argc >= 5 && !strcmp(argv[1], "files")) return mode_files(atoi(argv[2]), (uint64_t)atol(argv[3]), argv[4]);
	if (argc >= 5 && !strcmp(argv[1], "read")) return mode_read(atoi(argv[2]), (uint64_t)atol(argv[3]), argv[4]);
	fprintf(stderr, "usage: io_harness fmt <count> <seed> | files <n> <seed> <dir> | read <n> <seed> <dir> | emit <n> <dir>\n");
	return 2;
}
