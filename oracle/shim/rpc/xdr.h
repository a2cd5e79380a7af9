/*
 * TEST INFRASTRUCTURE ONLY (oracle build).  Header-only stand-in for the subset
 * of Sun-RPC XDR the reference calls (kd.c:16-28,137,170,186,202,218,1659-1683;
 * totipnat.c).  XDR for int/float/double is just big-endian 4/8-byte words over
 * a stdio stream, so that is all this implements.
 */
#ifndef SKID_SHIM_RPC_XDR_H
#define SKID_SHIM_RPC_XDR_H

#include "types.h"

enum xdr_op { XDR_ENCODE = 0, XDR_DECODE = 1, XDR_FREE = 2 };

typedef struct {
	enum xdr_op x_op;
	FILE *x_fp;
} XDR;

typedef bool_t (*xdrproc_t)();

static inline void xdrstdio_create(XDR *x, FILE *fp, enum xdr_op op)
{
	x->x_op = op;
	x->x_fp = fp;
}

static inline void xdr_destroy_(XDR *x) { (void)x; }
#define xdr_destroy(x) xdr_destroy_(x)

static inline bool_t xdr_shim_word(XDR *x, void *p, int nbytes)
{
	unsigned char b[8], *q = (unsigned char *)p;
	int i;
	if (x->x_op == XDR_DECODE) {
		if (fread(b, 1, nbytes, x->x_fp) != (size_t)nbytes) return FALSE;
		for (i = 0; i < nbytes; ++i) q[i] = b[nbytes - 1 - i];
	} else if (x->x_op == XDR_ENCODE) {
		for (i = 0; i < nbytes; ++i) b[i] = q[nbytes - 1 - i];
		if (fwrite(b, 1, nbytes, x->x_fp) != (size_t)nbytes) return FALSE;
	}
	return TRUE;
}

static inline bool_t xdr_int(XDR *x, int *p) { return xdr_shim_word(x, p, 4); }
static inline bool_t xdr_float(XDR *x, float *p) { return xdr_shim_word(x, p, 4); }
static inline bool_t xdr_double(XDR *x, double *p) { return xdr_shim_word(x, p, 8); }

static inline bool_t xdr_vector(XDR *x, char *base, u_int n, u_int sz, xdrproc_t proc)
{
	u_int i;
	(void)proc; /* every reference call site passes xdr_float with sz == 4 */
	for (i = 0; i < n; ++i)
		if (!xdr_shim_word(x, base + (size_t)i * sz, (int)sz)) return FALSE;
	return TRUE;
}

#endif
