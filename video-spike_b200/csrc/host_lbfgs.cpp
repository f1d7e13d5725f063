// Host half of FusedLBFGS (R5): torch.optim.LBFGS's two-loop recursion in COEFFICIENT space.
// The search direction d = -H g is a linear combination of the basis {g, s_0..s_{m-1}, y_0..y_{m-1}}; the recursion
// only needs inner products of basis vectors (gathered by vs_lbfgs_dots): sg[i] = s_i.g, yg[i] = y_i.g,
// SY[i][j] = s_i.y_j, YY[i][j] = y_i.y_j, gg = g.g.  O(m^2) flops on an m x m matrix (m <= 100): it runs on the host
// between the two device passes, inside the one host<->device round trip an L-BFGS iteration needs anyway.
// Same arithmetic as optim.py::lbfgs_two_loop (the readable Python statement, unit-tested against the explicit
// vector recursion); this compiled copy only removes ~60 us of interpreter time per iteration.
#include <stdint.h>

#include "../../include/vs_b200.h"

extern "C" int vs_host_lbfgs_two_loop(int m, double gg, const double* sg, const double* yg, const double* SY,
                                      const double* YY, int ld, double H_diag, double* coef_out, double* gtd_out) {
  if (m < 0 || m > VS_LBFGS_MAX_HIST || !coef_out || !gtd_out || (m > 0 && (!sg || !yg || !SY || !YY || ld < m)))
    return VS_ERR_INVALID;
  double al[VS_LBFGS_MAX_HIST], ro[VS_LBFGS_MAX_HIST];
  double* cs = coef_out + 1;
  double* cy = coef_out + 1 + m;
  for (int i = 0; i < m; ++i) ro[i] = 1.0 / SY[(long long)i * ld + i];
  // q = -g - sum_j al_j y_j
  for (int i = m - 1; i >= 0; --i) {
    double sq = -sg[i];
    for (int j = i + 1; j < m; ++j) sq -= al[j] * SY[(long long)i * ld + j];
    al[i] = sq * ro[i];
  }
  // r = H q + sum_j (al_j - be_j) s_j
  for (int i = 0; i < m; ++i) {
    double yr = -yg[i];
    for (int j = 0; j < m; ++j) yr -= al[j] * YY[(long long)i * ld + j];
    yr *= H_diag;
    for (int j = 0; j < i; ++j) yr += cs[j] * SY[(long long)j * ld + i];
    cs[i] = al[i] - yr * ro[i];
  }
  const double cg = -H_diag;
  coef_out[0] = cg;
  double dot_s = 0.0, dot_y = 0.0;
  for (int i = 0; i < m; ++i) {
    cy[i] = -H_diag * al[i];
  }
  for (int i = 0; i < m; ++i) dot_s += cs[i] * sg[i];
  for (int i = 0; i < m; ++i) dot_y += cy[i] * yg[i];
  *gtd_out = m ? cg * gg + dot_s + dot_y : cg * gg;
  return VS_OK;
}
