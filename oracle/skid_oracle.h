/*
 * TEST INFRASTRUCTURE ONLY - never linked or called by the product path.
 *
 * CPU restatement, in plain C, of the algorithm of the reference's group-finding hot path
 * (SKID v1.4.1: smooth1.c, kd.c, grav.c).  It restates WHAT the reference computes with
 * deliberately simple data structures (brute force / uniform grids instead of kd-trees), so
 * that every stage of the CUDA path can be checked against something small enough to read.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Pinning: every function here is checked in tests/test_oracle_cpu.py against
 * tests/golden/demo_golden.npz, which was produced by the UNMODIFIED reference compiled from
 * /root/reference (oracle/build_ref.sh; tests/golden/make_golden.py) - kNN radii bitwise,
 * densities, step-0 gradients and survivors, FoF partition, unbinding counts, the .stat file text -
 * and, as a whole stage script (oracle/pipeline.py), against goldens of the reference on synthetic dark,
 * gas+dark, gas-only, gas+dark+star, massive-halo and pruning-disabled boxes (tests/golden/synth_golden.npz,
 * species_golden.npz: same iteration counts, group counts, unbound counts, same-group fraction 1.0).
 */
#ifndef SKID_ORACLE_H
#define SKID_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Exact periodic k-nearest neighbours + spline density (smBallSearch smooth1.c:41-129,
 * smDensityInit smooth1.c:150-277).  pos: n*3 floats.  period: <= 0 for non-periodic.
 * Outputs: ball2[n] (k-th squared distance incl. self), rho[n], and optionally nbr[n*k]
 * (indices, ascending (d2,index)) / nbrd2[n*k]. */
void orc_knn_density(int n, const float *pos, const float *mass, int k, float period, float *ball2,
                     float *rho, int *nbr, float *nbrd2);

/* Periodic replica scatterers (smooth1.c:278-332).  Returns the count; if rep_src/rep_pos are
 * non-NULL fills source index and shifted position of each replica (capacity cap). */
int orc_replicas(int n, const float *pos, const float *ball2, float period, const float *center, int cap,
                 int *rep_src, float *rep_pos);

/* One gradient evaluation in gather form (smBallGather smooth1.c:338-384 + smAccDensity
 * smooth1.c:408-518): for every mover m and every ACTIVE scatterer entity e with
 * |x_e - x_m|^2 < ball2_e (float32, non-periodic): a_m += (x_e-x_m)*g(q)*fNorm_e.
 * ent_alive[e] != 0 marks active entities.  touched[e] is set for entities with >= 1 hit.
 * Returns fScatDens = min rho over touched entities (0 if none). */
void orc_gradient_abs(int nEnt, const float *epos, const float *eball2, const float *emass, int nMove,
                      const float *mpos, double *sabs);
float orc_gradient(int nEnt, const float *epos, const float *eball2, const float *emass, const float *erho,
                   const unsigned char *ent_alive, int nMove, const float *mpos, float *acc,
                   unsigned char *touched);

/* The whole flow loop (main.c:394-419) + kdMoveParticles (kd.c:702-732) + kdPruneInactive
 * (kd.c:735-793) + ScatterCut (smooth1.c:387-405), scatter form over a uniform mover grid.
 * Entities = n originals (+ replicas when period > 0).  mpos (nMove*3) is updated in place to the
 * converged positions.  log_nactive/log_nscatter (capacity maxlog) receive the "Ittr" lines.
 * Returns the number of Ittr lines. */
int orc_move_loop(int nEnt, const float *epos, const float *eball2, const float *emass, float *erho,
                  int nMove, float *mpos, float period, const float *center, float fCvg, float fStep,
                  int bInitial, int bNoPrune, int maxlog, int *log_nactive, int *log_nscatter,
                  int nMicro, float fMicroStep, float *mpos_at_fof);

/* Friends-of-friends (kdFoF kd.c:802-917): labels 1..G by ascending first member index;
 * returns G. */
int orc_fof(int n, const float *pos, float tau, float period, int *label);

/* kdUnbind (kd.c:1299-1466) for ONE group given in coordinates relative to its reference
 * point: r/v n*3, mass, soft, plus nScoop scoop sources (positions relative to the same point).
 * bSubPot: update potentials after each removal (pure dark / pure star inputs, kd.c:1441).
 * removed[i] = 1 for unbound members.  Returns the number removed; *boundMass, vcm[3] out. */
int orc_unbind_group(int n, const float *r, const float *v, const float *mass, const float *soft, int nScoop,
                     const float *sr, const float *smass, const float *ssoft, float G, float z, float fCosmo,
                     int iSoftType, int bNoUnbind, int bSubPot, unsigned char *removed, double *boundMass,
                     double *vcm);

/* kdOutStats (kd.c:1703-1839): the accumulators behind one .stat line.  The writer prints
 * ig, nMembers, fTotMass, fGasMass, fStarMass, sqrt(fVcirc), sqrt(fmVcirc), sqrt(flVcirc), fRVmax, fRhmass,
 * sqrt(fRouter2), (float)sqrt(fVdispSum/(3.0*nMembers)), rCenter, vcm, rBound with "%g".  pos/vel: n*3;
 * rCenter/vcm: nGroup*3 (row 0 unused); species by index range (gas, dark, star). */
typedef struct {
	int nMembers;
	float fTotMass, fGasMass, fStarMass;
	float fVcirc, fmVcirc, flVcirc;
	float fRVmax, fRhmass, fRouter2;
	float fVdispSum;
} orc_stat_row;
void orc_stats(int n, const float *pos, const float *vel, const float *mass, const float *soft, const float *temp,
               const float *rho, int nGas, int nDark, const int *piGroup, int nGroup, const float *rCenter,
               const float *vcm, const float *period, float G, float z, double dExpHub, float fDensMin,
               float fTempMax, orc_stat_row *rows);

#ifdef __cplusplus
}
#endif
#endif
