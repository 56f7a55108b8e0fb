/* abi_smoke.c -- the header consumed from plain C (gcc -std=c11), outside ctypes.
 *
 * Built by tests/test_abi.py (CPU suite: compiles, links against libode_b200.so, checks the struct layouts a
 * binding relies on, and -- without a GPU -- that a solve fails with ODE_B200_E_NO_DEVICE instead of falling back)
 * and run on the GPU box by tests/test_gpu_ensemble.py, where it solves the reference's README problem
 * (README.md:17-27: u' = 1.01 u, u(0) = 0.5, t in [0, 1], reltol = abstol = 1e-8 through solve(), i.e. with
 * dtmin = span/1e-9 and dtmax = span/2.5, src/common.jl:13-14) with ode45 and prints the result.
 *
 *   gcc -std=c11 -Wall -Wextra -Iinclude tests/abi_smoke.c -o abi_smoke -Lode.jl_b200 -lode_b200 -Wl,-rpath,...
 */
#include <assert.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ode_b200.h"

/* what julia/ODEB200.jl, ode.jl_b200/_lib.py and oracle/oracle.py marshal */
static_assert(sizeof(ode_b200_opts) == 112, "ode_b200_opts is 112 bytes");
static_assert(offsetof(ode_b200_opts, max_steps) == 32, "max_steps");
static_assert(offsetof(ode_b200_opts, reltol) == 40, "reltol");
static_assert(offsetof(ode_b200_opts, initstep) == 72, "initstep");
static_assert(offsetof(ode_b200_opts, mathmode) == 80, "mathmode");
static_assert(offsetof(ode_b200_opts, n_gpus) == 88, "n_gpus");
static_assert(offsetof(ode_b200_opts, fp_mode) == 92, "fp_mode");
static_assert(sizeof(ode_b200_ensemble_io) == 152, "ode_b200_ensemble_io is 152 bytes");
static_assert(offsetof(ode_b200_ensemble_io, t_final) == 40, "t_final");
static_assert(offsetof(ode_b200_ensemble_io, pts_cap) == 112, "pts_cap");
static_assert(sizeof(ode_b200_large_stats) == 48, "ode_b200_large_stats is 48 bytes");

int main(void) {
  if (ode_b200_version() != ODE_B200_ABI_VERSION) {
    fprintf(stderr, "ABI version mismatch\n");
    return 2;
  }
  ode_b200_opts o;
  memset(&o, 0, sizeof(o));
  o.method = ode_b200_method_id("ode45");
  o.rhs = ode_b200_rhs_id("linear");
  o.dof = ode_b200_rhs_dof(o.rhs);
  o.n_params = ode_b200_rhs_nparams(o.rhs);
  o.points = ODE_B200_POINTS_ALL;
  o.reltol = 1e-8;
  o.abstol = 1e-8;
  o.minstep = 1.0 / 1e-9; /* src/common.jl:13 (sic) */
  o.maxstep = 1.0 / 2.5;  /* :14 */
  const double u0 = 0.5, a = 1.01, tspan[2] = {0.0, 1.0};
  if (o.method != ODE_B200_M_DOPRI5 || o.rhs != ODE_B200_RHS_LINEAR || o.dof != 1 || o.n_params != 1) {
    fprintf(stderr, "id tables\n");
    return 2;
  }
  ode_b200_result* r = NULL;
  const int rc = ode_b200_solve_single(&o, &u0, &a, tspan, 2, &r);
  if (ode_b200_device_count() < 1) { /* no GPU: the call must refuse, not fall back */
    if (rc != ODE_B200_E_NO_DEVICE || r != NULL) {
      fprintf(stderr, "expected ODE_B200_E_NO_DEVICE without a GPU, got %d\n", rc);
      return 3;
    }
    printf("abi_smoke: no CUDA device -> ODE_B200_E_NO_DEVICE (\"%s\")\n", ode_b200_last_error());
    return 0;
  }
  if (rc != ODE_B200_OK) {
    fprintf(stderr, "solve failed: %d %s\n", rc, ode_b200_last_error());
    return 4;
  }
  const long long n = (long long)ode_b200_result_len(r);
  double* t = (double*)malloc((size_t)n * sizeof(double));
  double* y = (double*)malloc((size_t)n * sizeof(double));
  ode_b200_result_copy(r, t, y);
  const long long nacc = (long long)ode_b200_result_counts(r, 0), nrej = (long long)ode_b200_result_counts(r, 1);
  const int status = ode_b200_result_status(r);
  ode_b200_result_free(r);
  /* 11 accepted, 0 rejected, y(1) = 1.3728005108017367 (exact 0.5 exp(1.01) = 1.37280050751) */
  printf("abi_smoke: %lld points, %lld accepted, %lld rejected, status %d, t_end %.17g, y_end %.17g\n", n, nacc, nrej,
         status, t[n - 1], y[n - 1]);
  const int ok = n == 12 && nacc == 11 && nrej == 0 && status == ODE_B200_ST_OK && t[n - 1] == 1.0 &&
                 y[n - 1] == 1.3728005108017367 && fabs(y[n - 1] - 0.5 * exp(1.01)) < 1e-8;
  free(t);
  free(y);
  return ok ? 0 : 5;
}
