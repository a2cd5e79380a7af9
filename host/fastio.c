/*
 * Fast text/binary I/O helpers of the host driver ("I/O at scale").
 *
 * The reference writes its ASCII arrays with one fprintf per value (kd.c:1518-1519 "%d",
 * kd.c:1542-1545 "%.10g", kd.c:1570-1606 "%g") and decodes XDR one word at a time (kd.c:141-206);
 * at 2^24..2^27 particles that costs many times the GPU pipeline.  Here the formats are kept BYTE
 * IDENTICAL to glibc's printf and produced by:
 *   fmt_int / fmt_g   branch-light formatters; fmt_g takes a float promoted to double (what the
 *                     reference passes through "...") and falls back to snprintf whenever its
 *                     double-precision scaling cannot prove the correctly rounded digit string
 *                     (ties, values within 1e-5 of a rounding boundary, |10^k| beyond 10^22, inf/nan)
 *   par_for           a small pthread fork/join used to format chunks and to byte-swap/scatter
 *                     TIPSY records in parallel
 *   chunked_write     formats [0,n) in chunks on all threads, writes the chunks in order while the
 *                     next batch is being formatted
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "skid_host.h"

/* ---- threads ---------------------------------------------------------------------------- */
int host_threads(void)
{
	static int n = 0;
	if (!n) {
		const char *e = getenv("SKID_HOST_THREADS");
		long v = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
		if (v < 1) v = 1;
		if (v > 64) v = 64;
		n = (int)v;
	}
	return n;
}

typedef struct {
	par_fn fn;
	void *arg;
	size_t n;
	int tid, nt;
} par_task;

static void *par_tramp(void *p)
{
	par_task *t = (par_task *)p;
	size_t lo = t->n * (size_t)t->tid / (size_t)t->nt, hi = t->n * (size_t)(t->tid + 1) / (size_t)t->nt;
	if (hi > lo) t->fn(t->arg, lo, hi, t->tid);
	return NULL;
}

/* run fn over [0,n) split into one contiguous range per thread; small n stays on the caller */
void par_for(size_t n, size_t grain, par_fn fn, void *arg)
{
	int nt = host_threads(), i;
	pthread_t th[64];
	par_task task[64];
	if (grain && n / grain < (size_t)nt) nt = (int)(n / grain);
	if (nt <= 1) {
		if (n) fn(arg, 0, n, 0);
		return;
	}
	for (i = 0; i < nt; ++i) {
		task[i].fn = fn;
		task[i].arg = arg;
		task[i].n = n;
		task[i].tid = i;
		task[i].nt = nt;
		if (i && pthread_create(&th[i], NULL, par_tramp, &task[i])) {
			/* could not start a thread: run its share here */
			par_tramp(&task[i]);
			th[i] = th[0];
			task[i].nt = -1;
		}
	}
	par_tramp(&task[0]);
	for (i = 1; i < nt; ++i)
		if (task[i].nt > 0) pthread_join(th[i], NULL);
}

/* ---- integer formatting ----------------------------------------------------------------- */
static const char g_digits2[201] = "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
                                   "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
                                   "8081828384858687888990919293949596979899";

static char *put_u64(char *out, uint64_t v)
{
	char tmp[24];
	int n = 0;
	while (v >= 100) {
		const unsigned r = (unsigned)(v % 100);
		v /= 100;
		tmp[n++] = g_digits2[2 * r + 1];
		tmp[n++] = g_digits2[2 * r];
	}
	if (v >= 10) {
		tmp[n++] = g_digits2[2 * v + 1];
		tmp[n++] = g_digits2[2 * v];
	} else {
		tmp[n++] = (char)('0' + v);
	}
	while (n) *out++ = tmp[--n];
	return out;
}

/* "%d" */
char *fmt_int(char *out, int v)
{
	uint64_t u;
	if (v < 0) {
		*out++ = '-';
		u = (uint64_t)(-(int64_t)v);
	} else {
		u = (uint64_t)v;
	}
	return put_u64(out, u);
}

/* ---- "%.<prec>g" ------------------------------------------------------------------------ */
static const double g_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

static char *fmt_g_slow(char *out, double v, int prec)
{
	return out + snprintf(out, 40, "%.*g", prec, v);
}

/* "%.<prec>g" of v, 1 <= prec <= 10, byte-identical to glibc (round-half-even on the exact binary
 * value).  Needs <= 40 bytes.  The scaled value |v|*10^k is ONE correctly rounded double operation
 * (10^k is exact for k <= 22), so its absolute error is below 10^10 * 2^-53 = 1.2e-6: inside the
 * 1e-5 band within which the decision is handed to snprintf. */
char *fmt_g(char *out, double v, int prec)
{
	double a, s, f;
	uint64_t digits, lim;
	int e2, E, k, nd, i;
	char buf[20];
	if (v == 0.0) {
		if (signbit(v)) *out++ = '-';
		*out++ = '0';
		return out;
	}
	if (!(fabs(v) <= 1.7976931348623157e308) || prec < 1 || prec > 10) return fmt_g_slow(out, v, prec);
	a = fabs(v);
	(void)frexp(a, &e2); /* a = m * 2^e2, 0.5 <= m < 1 */
	E = (int)floor((e2 - 1) * 0.30102999566398120); /* floor(log10 a) or one less */
	if (E < -22 || E > 22) return fmt_g_slow(out, v, prec);
	if (E >= 0 ? a >= g_pow10[E] * 10.0 : a * g_pow10[-E - 1] >= 1.0) {
		/* the estimate was one short; only used to pick k, the exact decision is made on s below */
		++E;
	}
	k = prec - 1 - E;
	if (k > 22 || k < -22) return fmt_g_slow(out, v, prec);
	s = k >= 0 ? a * g_pow10[k] : a / g_pow10[-k];
	lim = (uint64_t)g_pow10[prec]; /* 10^prec */
	/* s should lie in [10^(prec-1), 10^prec); repair an off-by-one E from rounding at a power of ten */
	if (s >= (double)lim) {
		++E;
		--k;
		if (k < -22) return fmt_g_slow(out, v, prec);
		s = k >= 0 ? a * g_pow10[k] : a / g_pow10[-k];
	} else if (s < g_pow10[prec - 1]) {
		--E;
		++k;
		if (k > 22) return fmt_g_slow(out, v, prec);
		s = k >= 0 ? a * g_pow10[k] : a / g_pow10[-k];
	}
	if (!(s >= g_pow10[prec - 1] * (1.0 - 1e-12) && s < (double)lim)) return fmt_g_slow(out, v, prec);
	digits = (uint64_t)s;
	f = s - (double)digits;
	/* near a tie, or so close to an integer that the boundary of the decade is unclear: be exact */
	if (fabs(f - 0.5) < 1e-5) return fmt_g_slow(out, v, prec);
	if (f < 1e-5 || f > 1.0 - 1e-5) {
		/* rounding direction is clear (down / up to the next integer) unless the integer itself is a
		 * decade boundary, where the exponent depends on which side the exact value lies */
		uint64_t r = digits + (f > 0.5);
		if (r == lim || r == lim / 10) return fmt_g_slow(out, v, prec);
	}
	if (f > 0.5) ++digits;
	if (digits >= lim) {
		digits = lim / 10;
		++E;
	}
	if (digits < lim / 10) return fmt_g_slow(out, v, prec);
	/* digit string, prec digits, then strip trailing zeros */
	for (i = prec - 1; i >= 0; --i) {
		buf[i] = (char)('0' + digits % 10);
		digits /= 10;
	}
	nd = prec;
	while (nd > 1 && buf[nd - 1] == '0') --nd;
	if (v < 0) *out++ = '-';
	if (E < -4 || E >= prec) {
		int ae = E < 0 ? -E : E;
		*out++ = buf[0];
		if (nd > 1) {
			*out++ = '.';
			memcpy(out, buf + 1, (size_t)(nd - 1));
			out += nd - 1;
		}
		*out++ = 'e';
		*out++ = E < 0 ? '-' : '+';
		if (ae >= 100) {
			*out++ = (char)('0' + ae / 100);
			ae %= 100;
		}
		*out++ = (char)('0' + ae / 10);
		*out++ = (char)('0' + ae % 10);
	} else if (E >= 0) {
		/* E+1 integer digits (zero padded if they were stripped), then the rest */
		for (i = 0; i <= E; ++i) *out++ = i < nd ? buf[i] : '0';
		if (nd > E + 1) {
			*out++ = '.';
			memcpy(out, buf + E + 1, (size_t)(nd - E - 1));
			out += nd - E - 1;
		}
	} else {
		*out++ = '0';
		*out++ = '.';
		for (i = 0; i < -E - 1; ++i) *out++ = '0';
		memcpy(out, buf, (size_t)nd);
		out += nd;
	}
	return out;
}

/* ---- chunked parallel writer ------------------------------------------------------------ */
typedef struct {
	chunk_fmt_fn fn;
	void *arg;
	size_t base, n, chunk, max_per_item;
	char **buf;
	size_t *len;
} cw_state;

static void cw_worker(void *p, size_t lo, size_t hi, int tid)
{
	cw_state *st = (cw_state *)p;
	size_t c;
	(void)tid;
	for (c = lo; c < hi; ++c) {
		size_t a = st->base + c * st->chunk, b = a + st->chunk;
		if (b > st->n) b = st->n;
		st->len[c] = (size_t)(st->fn(st->arg, a, b, st->buf[c]) - st->buf[c]);
	}
}

typedef struct {
	FILE *fp;
	char **buf;
	size_t *len;
	size_t count;
	int rc;
} cw_flush;

static void *cw_flush_thread(void *p)
{
	cw_flush *f = (cw_flush *)p;
	size_t c;
	for (c = 0; c < f->count; ++c)
		if (fwrite(f->buf[c], 1, f->len[c], f->fp) != f->len[c]) f->rc = -1;
	return NULL;
}

/* Formats items [0,n) with fn (which appends the text of items [lo,hi) to out and returns the end;
 * at most max_per_item bytes per item) on all host threads and writes the pieces in order.  Two
 * slabs: while one batch of chunks is being written by a helper thread the next one is formatted. */
int chunked_write(FILE *fp, size_t n, size_t max_per_item, chunk_fmt_fn fn, void *arg)
{
	const size_t chunk = 1u << 15;
	const size_t per_batch = (size_t)host_threads() * 4;
	cw_state st[2];
	cw_flush fl;
	pthread_t writer;
	char *slab[2] = {NULL, NULL};
	size_t done = 0, c;
	int cur = 0, writing = 0, rc = 0, i;
	if (!n) return 0;
	memset(st, 0, sizeof st);
	fl.rc = 0;
	for (i = 0; i < 2; ++i) {
		st[i].fn = fn;
		st[i].arg = arg;
		st[i].n = n;
		st[i].chunk = chunk;
		st[i].max_per_item = max_per_item;
		st[i].buf = (char **)malloc(per_batch * sizeof(char *));
		st[i].len = (size_t *)malloc(per_batch * sizeof(size_t));
		slab[i] = (char *)malloc(per_batch * chunk * max_per_item);
		if (!st[i].buf || !st[i].len || !slab[i]) rc = -1;
		else
			for (c = 0; c < per_batch; ++c) st[i].buf[c] = slab[i] + c * chunk * max_per_item;
	}
	while (done < n && !rc) {
		const size_t left = (n - done + chunk - 1) / chunk, batch = left < per_batch ? left : per_batch;
		st[cur].base = done;
		par_for(batch, 1, cw_worker, &st[cur]);
		if (writing) {
			pthread_join(writer, NULL);
			writing = 0;
			if (fl.rc) rc = -1;
		}
		fl.fp = fp;
		fl.buf = st[cur].buf;
		fl.len = st[cur].len;
		fl.count = batch;
		if (pthread_create(&writer, NULL, cw_flush_thread, &fl) == 0) writing = 1;
		else {
			cw_flush_thread(&fl);
			if (fl.rc) rc = -1;
		}
		done += batch * chunk;
		cur ^= 1;
	}
	if (writing) {
		pthread_join(writer, NULL);
		if (fl.rc) rc = -1;
	}
	for (i = 0; i < 2; ++i) {
		free(st[i].buf);
		free(st[i].len);
		free(slab[i]);
	}
	return rc;
}

/* ---- ASCII integer array reader (kdInGroup's fscanf("%d") loop, kd.c:943-958) -------------- */
/* Parses up to n whitespace-separated decimal integers from text[0..len) into out; like a loop of
 * fscanf("%d") a malformed token stops the conversion for good.  Returns how many were parsed. */
size_t parse_ints(const char *text, size_t len, int *out, size_t n)
{
	size_t i = 0, k = 0;
	while (k < n) {
		int neg = 0, any = 0;
		int64_t v = 0;
		while (i < len && (text[i] == ' ' || (text[i] >= '\t' && text[i] <= '\r'))) ++i;
		if (i >= len) break;
		if (text[i] == '+' || text[i] == '-') {
			neg = text[i] == '-';
			++i;
		}
		while (i < len && text[i] >= '0' && text[i] <= '9') {
			if (v < ((int64_t)1 << 40)) v = v * 10 + (text[i] - '0');
			++i;
			any = 1;
		}
		if (!any) break;
		out[k++] = (int)(neg ? -v : v);
	}
	return k;
}
