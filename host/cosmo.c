/* The one cosmology scalar the hot path needs: H(a).  Restates csmExp2Hub (reference
 * cosmo.c:46-58) with the same operation order; everything else in cosmo.c/romberg.c/runge.c is
 * unused by SKID. */
#include <math.h>
#include "skid_host.h"

double cosmo_exp2hub(double a, double H0, double Omega0, double Lambda, double OmegaRad, double Quintess)
{
	const double curve = 1.0 - Omega0 - Lambda - OmegaRad - Quintess;
	const double e2 = Omega0 * a + curve * a * a + OmegaRad + Quintess * a * a * sqrt(a) + Lambda * a * a * a * a;
	return H0 * sqrt(e2) / (a * a);
}
