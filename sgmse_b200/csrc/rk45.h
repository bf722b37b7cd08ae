// Host-side controller of the probability-flow ODE sampler.
//
// The reference integrates the reverse ODE with scipy.integrate.solve_ivp(method='RK45') and moves the whole state
// through host numpy for every right-hand-side evaluation (/root/reference/sgmse/sampling/__init__.py:117-141).  The
// integrator is third-party (scipy, pinned 1.10.1 in requirements_version.txt:14; 1.18.1 in this image): its published
// algorithm -- Dormand & Prince 5(4) with the step-size control of Hairer, Norsett & Wanner, "Solving Ordinary
// Differential Equations I", Sec. II.4, in the form of scipy/integrate/_ivp/rk.py (RungeKutta._step_impl, rk_step),
// common.py (select_initial_step, norm) and base.py (OdeSolver.step) -- is restated here as a scalar state machine.
// All vector work goes through `Ops`, so the same controller drives the CUDA kernels of the sampler (engine.cu:
// OdeDeviceOps -- only ONE double per norm crosses the PCIe bus) and a plain host callback
// (sgmse_b200_rk45_host, CPU-tested against scipy itself in tests/test_cabi_host.py).
//
// Ops contract (y: current state, K[0..6]: stage derivatives, "stage": the argument of the next evaluation):
//   size_t size()                                            number of (complex) unknowns
//   void   combine(int s, const double* a, double h, bool to_new)
//                                                            stage = y + (sum_{j<s} a[j] K[j]) * h ; to_new: y_new = stage
//   void   eval(double t, int slot)                          K[slot] = f(t, stage)
//   double norm_init(int which, double rtol, double atol)    rms(v / (atol + |y| rtol)), v = y | K[0] | K[1] - K[0]
//   double norm_err(const double* E, double h, double rtol, double atol)
//                                                            rms((sum_{j<7} E[j] K[j]) * h / (atol + max(|y|, |y_new|) rtol))
//   void   accept()                                          y = y_new, K[0] = K[6]   (first-same-as-last)
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>

namespace sgmse {
namespace rk45 {

constexpr int kStages = 6;
constexpr double C[6] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
constexpr double A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
constexpr double B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
constexpr double E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
constexpr double kSafety = 0.9, kMinFactor = 0.2, kMaxFactor = 10.0;
constexpr double kErrorExponent = -1.0 / 5.0;   // -1 / (error_estimator_order + 1), error_estimator_order = 4

struct Result {
  double t = 0;       // time reached
  int nfev = 0;       // right-hand-side evaluations (what the reference returns as `nfe`)
  int steps = 0;      // accepted steps
  int rejected = 0;   // rejected step attempts
  int status = 0;     // 0: reached t_bound; -1: required step size below the spacing of doubles (scipy's "failed");
                      // -2: max_attempts exhausted; -3: non-finite error norm or step size.  -2 / -3 are not scipy states:
                      // with a NaN right-hand side scipy's loop never terminates (h_abs becomes NaN and every comparison
                      // is false); here the solve stops and reports it
};

template <class Ops>
Result solve(Ops& ops, double t0, double t_bound, double rtol, double atol, int max_attempts) {
  Result r;
  if (rtol < 100 * DBL_EPSILON) rtol = 100 * DBL_EPSILON;                    // validate_tol
  const double direction = t_bound != t0 ? (t_bound > t0 ? 1.0 : -1.0) : 1.0;
  double t = t0;
  ops.combine(0, nullptr, 0.0, false);
  ops.eval(t, 0);                                                            // self.f = fun(t0, y0)
  ++r.nfev;
  // ---- select_initial_step ----
  double h_abs;
  const double interval = std::fabs(t_bound - t0);
  if (ops.size() == 0) h_abs = INFINITY;
  else if (interval == 0.0) h_abs = 0.0;
  else {
    const double d0 = ops.norm_init(0, rtol, atol), d1 = ops.norm_init(1, rtol, atol);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    h0 = std::min(h0, interval);
    const double one = 1.0;
    ops.combine(1, &one, h0 * direction, false);                             // y1 = y0 + h0 * direction * f0
    ops.eval(t0 + h0 * direction, 1);
    ++r.nfev;
    const double d2 = ops.norm_init(2, rtol, atol) / h0;
    const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? std::max(1e-6, h0 * 1e-3) : std::pow(0.01 / std::max(d1, d2), 1.0 / 5.0);
    h_abs = std::min(std::min(100 * h0, h1), interval);
  }
  // ---- OdeSolver.step until finished ----
  int attempts = 0;
  for (;;) {
    if (ops.size() == 0 || t == t_bound) { t = t_bound; r.status = 0; break; }
    const double min_step = 10 * std::fabs(std::nextafter(t, direction * INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;                                  // (max_step = inf)
    bool accepted = false, rejected = false, failed = false, exhausted = false, nonfinite = !std::isfinite(h_abs);
    double t_new = t;
    while (!accepted && !nonfinite) {
      if (h_abs < min_step) { failed = true; break; }
      if (attempts >= max_attempts) { exhausted = true; break; }
      ++attempts;
      double h = h_abs * direction;
      t_new = t + h;
      if (direction * (t_new - t_bound) > 0) t_new = t_bound;
      h = t_new - t;
      h_abs = std::fabs(h);
      for (int s = 1; s < kStages; ++s) {                                    // rk_step
        ops.combine(s, A[s], h, false);
        ops.eval(t + C[s] * h, s);
        ++r.nfev;
      }
      ops.combine(kStages, B, h, true);
      ops.eval(t + h, kStages);
      ++r.nfev;
      const double err = ops.norm_err(E, h, rtol, atol);
      if (!std::isfinite(err) && !(err > 0)) { nonfinite = true; break; }      // NaN (inf just rejects the step)
      if (err < 1) {
        double factor = err == 0 ? kMaxFactor : std::min(kMaxFactor, kSafety * std::pow(err, kErrorExponent));
        if (rejected) factor = std::min(1.0, factor);
        h_abs *= factor;
        accepted = true;
      } else {
        // err = +inf: pow() gives 0 and the step shrinks by MIN_FACTOR, like Python's max(0.2, 0.0)
        const double f = kSafety * std::pow(err, kErrorExponent);
        h_abs *= (kMinFactor < f) ? f : kMinFactor;
        rejected = true;
        ++r.rejected;
      }
    }
    if (nonfinite) { r.status = -3; break; }
    if (failed) { r.status = -1; break; }
    if (exhausted) { r.status = -2; break; }
    ops.accept();
    ++r.steps;
    t = t_new;
    if (direction * (t - t_bound) >= 0) { r.status = 0; break; }
  }
  r.t = t;
  return r;
}

}  // namespace rk45
}  // namespace sgmse
