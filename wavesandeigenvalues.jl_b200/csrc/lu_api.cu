// placeholder until the LU lands
#include "lu.h"
struct LuSolver { int dummy; };
extern "C" {
int32_t wae_lu_analyze(wae_ctx* h, int32_t, int32_t*, int64_t*, double*) { if (h) h->err = "LU not built yet"; return WAE_E_INVALID; }
int32_t wae_lu_factor(wae_ctx* h, int32_t, int32_t) { if (h) h->err = "LU not built yet"; return WAE_E_INVALID; }
int32_t wae_lu_solve(wae_ctx* h, int32_t, int32_t, int32_t, double*) { if (h) h->err = "LU not built yet"; return WAE_E_INVALID; }
int32_t wae_eigs_si(wae_ctx* h, int32_t, int32_t, int32_t, int32_t, int32_t, const double*, double*, double*, int32_t*) { if (h) h->err = "LU not built yet"; return WAE_E_INVALID; }
int32_t wae_beyn_moments(wae_ctx* h, int32_t, int32_t, int32_t, const double*, const double*, const double*, int32_t, int32_t, void*) { if (h) h->err = "LU not built yet"; return WAE_E_INVALID; }
}
