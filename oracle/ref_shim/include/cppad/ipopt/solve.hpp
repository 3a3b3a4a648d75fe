// cppad/ipopt/solve.hpp -- STAND-IN for CppAD's Ipopt glue (test infrastructure, NOT product code).
//
// Same call as the reference makes at src/control/MPC.cpp:290-292.  Like CppAD::ipopt::solve without
// "Retape true", FG_eval is recorded ONCE, at the start point xi; f, g, the Jacobian and the Hessian
// of the Lagrangian then come from that tape.  Ipopt itself is not installed: the NLP is handed to
// the oracle's interior-point core (oracle/mpc_oracle.c, orc_ipm_solve -- a restatement of Ipopt 3.12
// with its defaults).  So here the reference's OWN text defines the problem; only the solver is restated.
#ifndef MPC_REF_SHIM_IPOPT_SOLVE_HPP
#define MPC_REF_SHIM_IPOPT_SOLVE_HPP

#include <cppad/cppad.hpp>
#include <sstream>
#include <string>

extern "C" {
#include "mpc_oracle.h"
}

namespace CppAD {
namespace ipopt {

template <class Dvector> class solve_result {
 public:
  enum status_type {
    not_defined, success, maxiter_exceeded, stop_at_tiny_step, stop_at_acceptable_point,
    local_infeasibility, user_requested_stop, feasible_point_found, diverging_iterates,
    restoration_failure, error_in_step_computation, invalid_number_detected,
    too_few_degrees_of_freedom, internal_error, unknown
  };
  status_type status;
  Dvector x, zl, zu, g, lambda;
  double obj_value;
  solve_result() : status(not_defined), obj_value(0.0) {}
};

// statistics of the last solve on this thread (the reference's MPC::solve returns no status)
struct shim_stats { int status, iters, tape_size; };
inline shim_stats &last_stats() { static thread_local shim_stats s = {0, 0, 0}; return s; }

namespace detail {
struct user { TapeFun *f; int n, m; std::vector<double> w; };
inline double cb_f(void *u, const double *x) { user *q = (user *)u; q->f->forward0(x); return q->f->value(0); }
inline void cb_grad(void *u, const double *x, double *g) {
  user *q = (user *)u;
  q->f->forward0(x);
  q->w.assign(q->m + 1, 0.0); q->w[0] = 1.0;
  q->f->reverse(q->w.data(), g, 0, false);
}
inline void cb_g(void *u, const double *x, double *c) {
  user *q = (user *)u;
  q->f->forward0(x);
  for (int j = 0; j < q->m; j++) c[j] = q->f->value(j + 1);
}
inline void cb_jac(void *u, const double *x, double *J) {
  user *q = (user *)u;
  q->f->forward0(x);
  for (int j = 0; j < q->m; j++) {
    q->w.assign(q->m + 1, 0.0); q->w[j + 1] = 1.0;
    q->f->reverse(q->w.data(), J + (size_t)j * q->n, 0, false);
  }
}
inline void cb_hess(void *u, const double *x, double sigma, const double *lam, double *H) {
  user *q = (user *)u;
  q->f->forward0(x);
  q->w.assign(q->m + 1, 0.0); q->w[0] = sigma;
  for (int j = 0; j < q->m; j++) q->w[j + 1] = lam[j];
  std::vector<double> g(q->n);
  for (int i = 0; i < q->n; i++) {
    q->f->forward1(i);
    q->f->reverse(q->w.data(), g.data(), H + (size_t)i * q->n, true);   // column i == row i (symmetric)
  }
}
}  // namespace detail

template <class Dvector, class FG_eval>
void solve(const std::string &options, const Dvector &xi, const Dvector &xl, const Dvector &xu,
           const Dvector &gl, const Dvector &gu, FG_eval &fg_eval, solve_result<Dvector> &solution) {
  typedef typename FG_eval::ADvector ADvector;
  const int n = (int)xi.size(), m = (int)gl.size();
  orc_config opt;
  orc_config_defaults(&opt);
  bool retape = false;
  {  // option lines: "Retape b", "Sparse b mode", "Integer|Numeric|String name value"
    std::istringstream in(options);
    std::string kind, name, val;
    while (in >> kind) {
      if (kind == "Retape") { in >> val; retape = (val == "true"); }
      else if (kind == "Sparse") { in >> val >> name; }
      else {
        in >> name >> val;
        if (name == "tol") opt.tol = atof(val.c_str());
        else if (name == "max_iter") opt.max_iter = atoi(val.c_str());
        // print_level, linear_solver, max_cpu_time: no counterpart (iteration cap instead of a clock)
      }
    }
  }
  solution.status = solve_result<Dvector>::unknown;
  if (retape) return;   // not supported by this stand-in (the reference never sets it)

  // ---- record FG_eval once, at xi
  Tape &T = tape();
  T.clear();
  T.recording = true;
  ADvector ax(n), afg(m + 1);
  TapeFun F;
  for (int i = 0; i < n; i++) {
    ax[i] = AD<double>(xi[i], T.push(OP_IND, i, -1, xi[i]));
    F.ind.push_back(ax[i].id_);
  }
  T.n_ind = n;
  fg_eval(afg, ax);
  T.recording = false;
  for (int j = 0; j <= m; j++) { F.dep.push_back(afg[j].id_); F.dep_const.push_back(afg[j].v_); }
  F.t = T;

  detail::user U;
  U.f = &F; U.n = n; U.m = m;
  orc_nlp nlp = {n, m, &U, detail::cb_f, detail::cb_grad, detail::cb_g, detail::cb_jac, detail::cb_hess};
  std::vector<double> x0(n), l(n), u(n), cl(m), cu(m), x(n), lam(m), zl(n), zu(n);
  for (int i = 0; i < n; i++) { x0[i] = xi[i]; l[i] = xl[i]; u[i] = xu[i]; }
  for (int j = 0; j < m; j++) { cl[j] = gl[j]; cu[j] = gu[j]; }
  orc_ipm_stats st;
  if (orc_ipm_solve(&nlp, &opt, x0.data(), l.data(), u.data(), cl.data(), cu.data(), x.data(), lam.data(),
                    zl.data(), zu.data(), &st) != 0)
    return;
  solution.status = (typename solve_result<Dvector>::status_type)st.status;
  solution.obj_value = st.obj;
  solution.x = Dvector(n); solution.zl = Dvector(n); solution.zu = Dvector(n);
  solution.g = Dvector(m); solution.lambda = Dvector(m);
  F.forward0(x.data());
  for (int i = 0; i < n; i++) { solution.x[i] = x[i]; solution.zl[i] = zl[i]; solution.zu[i] = zu[i]; }
  for (int j = 0; j < m; j++) { solution.g[j] = F.value(j + 1); solution.lambda[j] = lam[j]; }
  last_stats().status = st.status;
  last_stats().iters = st.iters;
  last_stats().tape_size = (int)F.t.size();
}

}  // namespace ipopt
}  // namespace CppAD

#endif
