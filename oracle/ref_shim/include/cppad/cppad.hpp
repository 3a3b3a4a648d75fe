// cppad/cppad.hpp -- STAND-IN for CppAD (test infrastructure, NOT product code, NOT a copy of CppAD).
//
// The reference's hot path is written on CppAD::AD<double> (src/control/MPC.cpp:50-154,
// src/model/RoadGeometry.cpp:14-16,49-55, src/model/Vehicle.cpp:50-64, src/utils/utils.h:28-47) and
// CppAD is not installed in this image.  This header provides just enough of that interface for the
// reference's UNMODIFIED sources to compile: a recording scalar type whose operations are written to
// a tape (one tape per thread), comparison operators that return a plain bool from the values seen
// while recording -- which is exactly what freezes FG_eval's `if` branches at the start point when
// the tape is recorded once -- and tape replay with first- and second-order derivatives.
#ifndef MPC_REF_SHIM_CPPAD_HPP
#define MPC_REF_SHIM_CPPAD_HPP

#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <vector>

namespace CppAD {

enum TapeOp { OP_CONST, OP_IND, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SIN, OP_COS, OP_ATAN, OP_FABS };

struct Tape {
  std::vector<unsigned char> op;
  std::vector<int> a, b;
  std::vector<double> val;     // value while recording; constants keep it on replay
  bool recording = false;
  int n_ind = 0;
  void clear() { op.clear(); a.clear(); b.clear(); val.clear(); n_ind = 0; }
  int push(TapeOp o, int x, int y, double v) {
    op.push_back((unsigned char)o); a.push_back(x); b.push_back(y); val.push_back(v);
    return (int)op.size() - 1;
  }
  size_t size() const { return op.size(); }
};
inline Tape &tape() { static thread_local Tape t; return t; }

template <class Base> class AD;

template <> class AD<double> {
 public:
  double v_;
  int id_;   // node on the tape, -1 = constant (parameter)
  AD() : v_(0.0), id_(-1) {}
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
  AD(T c) : v_((double)c), id_(-1) {}
  AD(double v, int id) : v_(v), id_(id) {}
  int node() const { return id_ >= 0 ? id_ : tape().push(OP_CONST, -1, -1, v_); }
  static AD bin(TapeOp o, const AD &x, const AD &y, double v) {
    if ((x.id_ < 0 && y.id_ < 0) || !tape().recording) return AD(v, -1);
    int xn = x.node(), yn = y.node();
    return AD(v, tape().push(o, xn, yn, v));
  }
  static AD un(TapeOp o, const AD &x, double v) {
    if (x.id_ < 0 || !tape().recording) return AD(v, -1);
    return AD(v, tape().push(o, x.id_, -1, v));
  }
  AD operator-() const { return un(OP_NEG, *this, -v_); }
  AD &operator+=(const AD &y) { *this = bin(OP_ADD, *this, y, v_ + y.v_); return *this; }
  AD &operator-=(const AD &y) { *this = bin(OP_SUB, *this, y, v_ - y.v_); return *this; }
  AD &operator*=(const AD &y) { *this = bin(OP_MUL, *this, y, v_ * y.v_); return *this; }
  AD &operator/=(const AD &y) { *this = bin(OP_DIV, *this, y, v_ / y.v_); return *this; }
};

typedef AD<double> ADd;
inline ADd operator+(const ADd &x, const ADd &y) { return ADd::bin(OP_ADD, x, y, x.v_ + y.v_); }
inline ADd operator-(const ADd &x, const ADd &y) { return ADd::bin(OP_SUB, x, y, x.v_ - y.v_); }
inline ADd operator*(const ADd &x, const ADd &y) { return ADd::bin(OP_MUL, x, y, x.v_ * y.v_); }
inline ADd operator/(const ADd &x, const ADd &y) { return ADd::bin(OP_DIV, x, y, x.v_ / y.v_); }
#define MPC_SHIM_MIXED(OPSYM)                                                                           \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>              \
  inline ADd operator OPSYM(const ADd &x, T y) { return x OPSYM ADd(y); }                               \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>              \
  inline ADd operator OPSYM(T x, const ADd &y) { return ADd(x) OPSYM y; }
MPC_SHIM_MIXED(+) MPC_SHIM_MIXED(-) MPC_SHIM_MIXED(*) MPC_SHIM_MIXED(/)
#undef MPC_SHIM_MIXED
// comparisons: decided by the values at recording time (CppAD semantics without Retape / CondExp)
#define MPC_SHIM_CMP(OPSYM)                                                                             \
  inline bool operator OPSYM(const ADd &x, const ADd &y) { return x.v_ OPSYM y.v_; }                    \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>              \
  inline bool operator OPSYM(const ADd &x, T y) { return x.v_ OPSYM (double)y; }                        \
  template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>              \
  inline bool operator OPSYM(T x, const ADd &y) { return (double)x OPSYM y.v_; }
MPC_SHIM_CMP(<) MPC_SHIM_CMP(<=) MPC_SHIM_CMP(>) MPC_SHIM_CMP(>=) MPC_SHIM_CMP(==) MPC_SHIM_CMP(!=)
#undef MPC_SHIM_CMP

inline ADd sin(const ADd &x) { return ADd::un(OP_SIN, x, std::sin(x.v_)); }
inline ADd cos(const ADd &x) { return ADd::un(OP_COS, x, std::cos(x.v_)); }
inline ADd atan(const ADd &x) { return ADd::un(OP_ATAN, x, std::atan(x.v_)); }
inline ADd fabs(const ADd &x) { return ADd::un(OP_FABS, x, std::fabs(x.v_)); }
inline ADd abs(const ADd &x) { return fabs(x); }
inline double Value(const ADd &x) { return x.v_; }
inline std::ostream &operator<<(std::ostream &os, const ADd &x) { return os << x.v_; }

// CPPAD_TESTVECTOR
template <class T> class vector : public std::vector<T> {
 public:
  vector() {}
  explicit vector(size_t n) : std::vector<T>(n) {}
};
template <class T> inline std::ostream &operator<<(std::ostream &os, const vector<T> &v) {
  os << "{ ";
  for (size_t i = 0; i < v.size(); i++) os << (i ? ", " : "") << v[i];
  return os << " }";
}

// ---- a recorded function y = F(x): replay with derivatives --------------------------------------------
class TapeFun {
 public:
  Tape t;
  std::vector<int> ind;                 // nodes of the independents
  std::vector<int> dep;                 // node of every dependent, -1 = constant
  std::vector<double> dep_const;
  std::vector<double> v, d, bar, bard;  // work: values, tangents, adjoints, tangents of adjoints

  void forward0(const double *x) {
    const size_t K = t.size();
    v.resize(K);
    for (size_t k = 0; k < K; k++) {
      const int a = t.a[k], b = t.b[k];
      switch (t.op[k]) {
        case OP_CONST: v[k] = t.val[k]; break;
        case OP_IND: v[k] = x[a]; break;
        case OP_ADD: v[k] = v[a] + v[b]; break;
        case OP_SUB: v[k] = v[a] - v[b]; break;
        case OP_MUL: v[k] = v[a] * v[b]; break;
        case OP_DIV: v[k] = v[a] / v[b]; break;
        case OP_NEG: v[k] = -v[a]; break;
        case OP_SIN: v[k] = std::sin(v[a]); break;
        case OP_COS: v[k] = std::cos(v[a]); break;
        case OP_ATAN: v[k] = std::atan(v[a]); break;
        case OP_FABS: v[k] = std::fabs(v[a]); break;
      }
    }
  }
  double value(size_t j) const { return dep[j] >= 0 ? v[dep[j]] : dep_const[j]; }
  // tangents along direction e_i (after forward0)
  void forward1(int i) {
    const size_t K = t.size();
    d.assign(K, 0.0);
    for (size_t k = 0; k < K; k++) {
      const int a = t.a[k], b = t.b[k];
      switch (t.op[k]) {
        case OP_CONST: break;
        case OP_IND: d[k] = (a == i) ? 1.0 : 0.0; break;
        case OP_ADD: d[k] = d[a] + d[b]; break;
        case OP_SUB: d[k] = d[a] - d[b]; break;
        case OP_MUL: d[k] = d[a] * v[b] + v[a] * d[b]; break;
        case OP_DIV: d[k] = (d[a] - v[k] * d[b]) / v[b]; break;
        case OP_NEG: d[k] = -d[a]; break;
        case OP_SIN: d[k] = std::cos(v[a]) * d[a]; break;
        case OP_COS: d[k] = -std::sin(v[a]) * d[a]; break;
        case OP_ATAN: d[k] = d[a] / (1.0 + v[a] * v[a]); break;
        case OP_FABS: d[k] = (v[a] > 0 ? 1.0 : (v[a] < 0 ? -1.0 : 0.0)) * d[a]; break;
      }
    }
  }
  // reverse sweep for w^T F; second = true also propagates the tangents of the adjoints (needs
  // forward1), giving one column of the Hessian of w^T F in gd.
  void reverse(const double *w, double *g, double *gd, bool second) {
    const size_t K = t.size();
    bar.assign(K, 0.0);
    if (second) bard.assign(K, 0.0);
    for (size_t j = 0; j < dep.size(); j++)
      if (dep[j] >= 0) bar[dep[j]] += w[j];
    for (size_t kk = K; kk-- > 0;) {
      const int k = (int)kk, a = t.a[k], b = t.b[k];
      const double bk = bar[k], bdk = second ? bard[k] : 0.0;
      if (bk == 0.0 && bdk == 0.0) continue;
      switch (t.op[k]) {
        case OP_CONST: case OP_IND: break;
        case OP_ADD: bar[a] += bk; bar[b] += bk; if (second) { bard[a] += bdk; bard[b] += bdk; } break;
        case OP_SUB: bar[a] += bk; bar[b] -= bk; if (second) { bard[a] += bdk; bard[b] -= bdk; } break;
        case OP_MUL:
          bar[a] += bk * v[b]; bar[b] += bk * v[a];
          if (second) { bard[a] += bdk * v[b] + bk * d[b]; bard[b] += bdk * v[a] + bk * d[a]; }
          break;
        case OP_DIV: {
          const double iy = 1.0 / v[b];
          bar[a] += bk * iy; bar[b] -= bk * v[k] * iy;
          if (second) {
            bard[a] += bdk * iy - bk * d[b] * iy * iy;
            bard[b] += -(bdk * v[k] + bk * d[k]) * iy + bk * v[k] * d[b] * iy * iy;
          }
          break;
        }
        case OP_NEG: bar[a] -= bk; if (second) bard[a] -= bdk; break;
        case OP_SIN: {
          const double c = std::cos(v[a]);
          bar[a] += bk * c;
          if (second) bard[a] += bdk * c - bk * std::sin(v[a]) * d[a];
          break;
        }
        case OP_COS: {
          const double s = std::sin(v[a]);
          bar[a] -= bk * s;
          if (second) bard[a] -= bdk * s + bk * std::cos(v[a]) * d[a];
          break;
        }
        case OP_ATAN: {
          const double q = 1.0 / (1.0 + v[a] * v[a]);
          bar[a] += bk * q;
          if (second) bard[a] += bdk * q - bk * 2.0 * v[a] * d[a] * q * q;
          break;
        }
        case OP_FABS: {
          const double s = v[a] > 0 ? 1.0 : (v[a] < 0 ? -1.0 : 0.0);
          bar[a] += bk * s;
          if (second) bard[a] += bdk * s;
          break;
        }
      }
    }
    for (size_t i = 0; i < ind.size(); i++) {
      if (g) g[i] = bar[ind[i]];
      if (second && gd) gd[i] = bard[ind[i]];
    }
  }
};

}  // namespace CppAD

#define CPPAD_TESTVECTOR(T) CppAD::vector<T>

#endif
