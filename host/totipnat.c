/*
 * totipnat (B200 host tools): TIPSY "standard" (XDR, big-endian) -> native binary, stdin -> stdout.
 * Drop-in for the reference's converter (totipnat.c:59-135), which its demo pipes into skid
 * (`./totipnat < dark.std | ./skid ...`, demo:2): every snapshot in the stream is converted until the
 * next header cannot be read, and "read time <t>" goes to stderr per snapshot.  Layout: the native
 * header is struct dump {double time; int nbodies, ndim, nsph, ndark, nstar;} = 32 bytes with 4 bytes of
 * tail padding (tipsydefs.h:41-48); the standard header is the same six values big-endian plus one
 * explicit pad word (totipnat.c:8-27); records are plain float32 words (12 gas / 9 dark / 11 star).
 * The reference decodes one word per xdr_float call; here the records stream through a 64 MiB buffer
 * and are byte-swapped on all host threads (fastio.c par_for).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "skid_host.h"

static void swap_range(void *arg, size_t lo, size_t hi, int tid)
{
	uint32_t *w = (uint32_t *)arg;
	size_t i;
	(void)tid;
	for (i = lo; i < hi; ++i) w[i] = __builtin_bswap32(w[i]);
}

static size_t read_full(void *buf, size_t nbytes)
{
	size_t got = 0;
	while (got < nbytes) {
		size_t r = fread((char *)buf + got, 1, nbytes - got, stdin);
		if (r == 0) break;
		got += r;
	}
	return got;
}

int main(void)
{
	const size_t cap = (size_t)64 << 20;
	uint32_t *buf = (uint32_t *)malloc(cap);
	if (!buf) return 1;
	for (;;) {
		unsigned char raw[32];
		struct {
			double time;
			int nbodies, ndim, nsph, ndark, nstar, pad;
		} h;
		uint64_t t;
		uint32_t w[5];
		unsigned long long left;
		if (read_full(raw, 32) != 32) break;
		memcpy(&t, raw, 8);
		t = __builtin_bswap64(t);
		memcpy(&h.time, &t, 8);
		memcpy(w, raw + 8, 20);
		h.nbodies = (int)__builtin_bswap32(w[0]);
		h.ndim = (int)__builtin_bswap32(w[1]);
		h.nsph = (int)__builtin_bswap32(w[2]);
		h.ndark = (int)__builtin_bswap32(w[3]);
		h.nstar = (int)__builtin_bswap32(w[4]);
		h.pad = 0;
		if (h.nsph < 0 || h.ndark < 0 || h.nstar < 0) {
			fprintf(stderr, "totipnat: bad header (nsph %d ndark %d nstar %d)\n", h.nsph, h.ndark, h.nstar);
			return 1;
		}
		fwrite(&h, 32, 1, stdout);
		left = 4ull * (12ull * (unsigned)h.nsph + 9ull * (unsigned)h.ndark + 11ull * (unsigned)h.nstar);
		while (left) {
			const size_t want = left < cap ? (size_t)left : cap;
			const size_t got = read_full(buf, want);
			par_for(got / 4, (size_t)1 << 18, swap_range, buf);
			if (fwrite(buf, 1, got, stdout) != got) return 1;
			if (got < want) break; /* truncated input: what was there has been converted, like the reference */
			left -= got;
		}
		fprintf(stderr, "read time %lf\n", h.time);
	}
	free(buf);
	return fflush(stdout) ? 1 : 0;
}
