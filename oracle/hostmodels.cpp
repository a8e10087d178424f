// TEST INFRASTRUCTURE (oracle side) -- not part of the product path.
//
// Host build of the analytic model templates in drake_ddp_b200/csrc/models.h so the
// CPU oracle's dynamics shim evaluates exactly the arithmetic the CUDA kernels
// compile for the device.  Plays the role of Drake's CalcForcedDiscreteVariableUpdate
// on the double system (/root/reference/ilqr.py:223-229) and on the AutoDiffXd
// clone with n+m seeds (/root/reference/ilqr.py:253-270).
//
// Build: g++ -O2 -shared -fPIC oracle/hostmodels.cpp -o oracle/_build/libhostmodels.so
#include "../drake_ddp_b200/csrc/models.h"

namespace {
template <class Model>
int step_impl(const double* x, const double* u, const double* p, double* xn) {
  Model::template step<double>(x, u, xn, p);
  return 0;
}
template <class Model>
int jac_impl(const double* x, const double* u, const double* p, double* xn, double* fx,
             double* fu) {
  constexpr int n = Model::n, m = Model::m, K = n + m;
  typedef ddp::Dual<K> D;
  D* xs = new D[n];
  D* us = new D[m];
  D* out = new D[n];
  for (int i = 0; i < n; ++i) {
    xs[i] = D(x[i]);
    xs[i].d[i] = 1.0;
  }
  for (int j = 0; j < m; ++j) {
    us[j] = D(u[j]);
    us[j].d[n + j] = 1.0;
  }
  Model::template step<D>(xs, us, out, p);
  for (int i = 0; i < n; ++i) {
    if (xn) xn[i] = out[i].v;
    for (int j = 0; j < n; ++j) fx[i * n + j] = out[i].d[j];
    for (int j = 0; j < m; ++j) fu[i * m + j] = out[i].d[n + j];
  }
  delete[] xs;
  delete[] us;
  delete[] out;
  return 0;
}
}  // namespace

extern "C" {
int hostmodel_dims(int model_id, int* n, int* m, int* np) {
  DDP_MODEL_SWITCH(model_id, { *n = Model::n; *m = Model::m; *np = Model::np; });
  return 0;
}
// x+ = f(x, u)
int hostmodel_step(int model_id, const double* x, const double* u, const double* p, double* xn) {
  DDP_MODEL_SWITCH(model_id, return step_impl<Model>(x, u, p, xn));
  return 0;
}
// fx (n x n row-major), fu (n x m row-major) of the discrete map, exact forward-mode AD
int hostmodel_jac(int model_id, const double* x, const double* u, const double* p, double* xn,
                  double* fx, double* fu) {
  DDP_MODEL_SWITCH(model_id, return jac_impl<Model>(x, u, p, xn, fx, fu));
  return 0;
}
}
