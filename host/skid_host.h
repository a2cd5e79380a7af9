/*
 * Host side of the drop-in driver: SKID's command-line, TIPSY input and .grp/.gtp/.den/.ray/.stat
 * output surface (reference main.c, kd.c I/O functions), in C, calling the GPU hot path through
 * include/skidgpu.h.  Nothing here computes on the hot path.
 */
#ifndef SKID_HOST_H
#define SKID_HOST_H

#include <stdio.h>
#include "skidgpu.h"

typedef struct {
	double time;
	int n, nGas, nDark, nStar;
	skidgpu_pinit *p; /* file order, iOrder == index */
} snapshot;

/* kdReadTipsy (kd.c:122-222): native or XDR "standard" TIPSY binary from fp. 0 on success. */
int tipsy_read(FILE *fp, int bStandard, snapshot *s);

/* kdInGroup (kd.c:920-962): ASCII .grp.  Returns nGroup (max id + 1) or -1. */
int grp_read(const char *path, int n, int *piGroup);
/* kdReadCenter (kd.c:1083-1159): centres/velocities from a .gtp; 1 = read, 0 = no file, -1 = mismatch */
int gtp_read(const char *path, int bStandard, int nGroup, skidgpu_pgroup *g);

/* kdOutGroup (kd.c:1502-1523), kdOutDensity (1526-1547), kdOutVector (1550-1608), kdWriteGroup (1611-1687) */
int out_group(const char *path, int n, const int *piGroup);
int out_density(const char *path, int n, const float *rho);
int out_vector(const char *path, const snapshot *s, int nMove, const int *iOrder, const float *r3,
               const float fPeriod[3]);
int out_gtp(const char *path, int bStandard, double fTime, int nGroup, const skidgpu_pgroup *g);
/* print statement of kdOutStats (kd.c:1822-1836) over the rows of skidgpu_stats */
int out_stats(const char *path, int nGroup, const skidgpu_pgroup *g, const skidgpu_stat_row *row);

/* fastio.c: formatters byte-identical to printf's "%d" / "%.<prec>g", a pthread fork/join and an
 * ordered chunked writer */
#include <stddef.h>
typedef void (*par_fn)(void *arg, size_t lo, size_t hi, int tid);
typedef char *(*chunk_fmt_fn)(void *arg, size_t lo, size_t hi, char *out);
int host_threads(void);
void par_for(size_t n, size_t grain, par_fn fn, void *arg);
char *fmt_int(char *out, int v);
char *fmt_g(char *out, double v, int prec);
int chunked_write(FILE *fp, size_t n, size_t max_per_item, chunk_fmt_fn fn, void *arg);
size_t parse_ints(const char *text, size_t len, int *out, size_t n);

/* csmExp2Hub (cosmo.c:46-58) */
double cosmo_exp2hub(double dExp, double H0, double Omega0, double Lambda, double OmegaRad, double Quintess);

#endif
