#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference (SKID v1.4.1) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored):
#
#   oracle/_ref/skid_ref        unmodified main.c kd.c smooth1.c grav.c cosmo.c romberg.c runge.c
#   oracle/_ref/totipnat_ref    unmodified totipnat.c
#   oracle/_ref/skid_ref_dump   same sources + test-only instrumentation inserted into
#                               temporary copies (deleted after the compile):
#                                 SKID_DUMP=<prefix>  -> <prefix>.knn/.step0/.fof/.ub0/.ub1 binary dumps
#                                 SKID_NOPRUNE=1      -> never prune scatterers (the "-nsp" semantics)
#
# The reference's own Makefile needs libtirpc (absent here); oracle/shim/rpc/*.h is a
# header-only XDR stand-in.  No -march/-mfma/-ffast-math: the golden outputs are plain
# x86-64 SSE2 arithmetic (no FMA contraction), which is what the GPU path reproduces.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SKID_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
	echo "build_ref: $REF not present (GPU box) - using prebuilt $OUT" >&2
	exit 0
fi
mkdir -p "$OUT"
CFLAGS="-O3 -D_GNU_SOURCE -w -I$HERE/shim -I$REF"
SRCS="main.c kd.c smooth1.c grav.c cosmo.c romberg.c runge.c"

objs=""
for s in $SRCS; do
	gcc $CFLAGS -c "$REF/$s" -o "$OUT/${s%.c}.o"
	objs="$objs $OUT/${s%.c}.o"
done
gcc -O3 -o "$OUT/skid_ref" $objs -lm
gcc $CFLAGS -o "$OUT/totipnat_ref" "$REF/totipnat.c" -lm

# ---- instrumented variant: insert our own dump statements into temporary copies ----
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp "$REF"/main.c "$REF"/kd.c "$REF"/smooth1.c "$TMP"/

# smooth1.c: (a) globals after the includes, (b) kNN dump after fBall2 is stored (line 243),
# (c) SKID_NOPRUNE: bInitial=0 on entry to smAccDensity (after line 418) and skip ScatterCut (509-513)
sed -i \
 -e '5a\
#include <stdlib.h>\
static FILE *orc_knn; static int orc_knn_init;' \
 -e '243a\
	if (!orc_knn_init) { char orc_f[512]; orc_knn_init = 1; if (getenv("SKID_DUMP")) { snprintf(orc_f,sizeof orc_f,"%s.knn",getenv("SKID_DUMP")); orc_knn = fopen(orc_f,"wb"); fwrite(&smx->kd->nInitActive,4,1,orc_knn); fwrite(&nSmooth,4,1,orc_knn);} }\
	if (orc_knn) { PQ *orc_q; fwrite(&p[pi].iOrder,4,1,orc_knn); fwrite(&p[pi].fBall2,4,1,orc_knn); for (orc_q=smx->pq;orc_q<=pqLast;++orc_q) { fwrite(&p[orc_q->p].iOrder,4,1,orc_knn); fwrite(&orc_q->fKey,4,1,orc_knn);} }' \
 -e '278a\
	if (orc_knn) { fclose(orc_knn); orc_knn = NULL; }' \
 -e '418a\
	if (getenv("SKID_NOPRUNE")) bInitial = 0;' \
 -e '508a\
	if (!getenv("SKID_NOPRUNE")) {' \
 -e '513a\
	}' \
 "$TMP/smooth1.c"

# main.c: dump step-0 accelerations + surviving scatterers before the first Ittr line (line 402)
sed -i \
 -e '401a\
	if (getenv("SKID_DUMP")) { char orc_f[512]; FILE *orc; int orc_i; snprintf(orc_f,sizeof orc_f,"%s.step0",getenv("SKID_DUMP")); orc = fopen(orc_f,"wb");\
		fwrite(&kd->nMove,4,1,orc); for (orc_i=0;orc_i<kd->nMove;++orc_i) { fwrite(&kd->pMove[orc_i].iOrder,4,1,orc); fwrite(kd->pMove[orc_i].a,4,3,orc); }\
		fwrite(&kd->nParticles,4,1,orc); fwrite(&kd->nInitActive,4,1,orc); for (orc_i=0;orc_i<kd->nParticles;++orc_i) { fwrite(&kd->pInit[orc_i].iOrder,4,1,orc); fwrite(&kd->pInit[orc_i].fBall2,4,1,orc); fwrite(&kd->pInit[orc_i].fDensity,4,1,orc); }\
		fwrite(&smx->nExtraScat,4,1,orc); for (orc_i=0;orc_i<smx->nExtraScat;++orc_i) { fwrite(&smx->pp[orc_i].iOrder,4,1,orc); fwrite(smx->pp[orc_i].r,4,3,orc); }\
		fclose(orc); }' \
 "$TMP/main.c"

# kd.c: (a) FoF dump after the labelling loop (line 904), (b) catalogue before/after unbinding (1326, 1461)
sed -i \
 -e '13a\
#include <string.h>\
static void orc_dump_groups(KD kd,const char *ext) { char orc_f[512]; FILE *orc; if (!getenv("SKID_DUMP")) return; snprintf(orc_f,sizeof orc_f,"%s.%s",getenv("SKID_DUMP"),ext); orc = fopen(orc_f,"wb"); fwrite(&kd->nParticles,4,1,orc); fwrite(&kd->nGroup,4,1,orc); fwrite(kd->piGroup,4,kd->nParticles,orc); fwrite(kd->pGroup,sizeof(PGROUP),kd->nGroup,orc); fclose(orc); }' \
 -e '904a\
	if (getenv("SKID_DUMP")) { char orc_f[512]; FILE *orc; snprintf(orc_f,sizeof orc_f,"%s.fof",getenv("SKID_DUMP")); orc = fopen(orc_f,"wb"); fwrite(&kd->nActive,4,1,orc); for (pn=0;pn<kd->nActive;++pn) { fwrite(&p[pn].iOrder,4,1,orc); fwrite(p[pn].r,4,3,orc); fwrite(&Group[pn],4,1,orc);} fclose(orc); }' \
 -e '1326a\
	orc_dump_groups(kd,"ub0");' \
 -e '1461a\
	orc_dump_groups(kd,"ub1");' \
 "$TMP/kd.c"

objs=""
for s in $SRCS; do
	src="$REF/$s"; [ -f "$TMP/$s" ] && src="$TMP/$s"
	gcc $CFLAGS -c "$src" -o "$TMP/${s%.c}.o"
	objs="$objs $TMP/${s%.c}.o"
done
gcc -O3 -o "$OUT/skid_ref_dump" $objs -lm
rm -f "$OUT"/*.o
echo "build_ref: built $OUT/skid_ref $OUT/skid_ref_dump $OUT/totipnat_ref"
