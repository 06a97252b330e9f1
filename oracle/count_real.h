/* count_real.h -- operation-counting stand-in for `double` (C++ only, -DSAMSIM_COUNT_OPS).
 *
 * TEST INFRASTRUCTURE ONLY.  Compiling oracle/samsim_oracle.c as C++ with `real` = this type counts the
 * source-level FP64 operations the reference algorithm performs: every +, -, *, / counts 1 (unary minus and
 * comparisons are not flops), real-exponent pow / exp / sin calls are counted separately.  Used by
 * tools/count_flops.py to obtain F_alg (flop per column-timestep) for bench.py's roofline numerator. */
#ifndef SAMSIM_COUNT_REAL_H
#define SAMSIM_COUNT_REAL_H
#include <cmath>

struct sam_op_counts { long long add, mul, div, cmp, pw, ex, sn; };
extern sam_op_counts g_sam_ops;

struct real {
  double v;
  real() = default;
  real(double x) : v(x) {}
  real(int x) : v((double)x) {}
  real(long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator long() const { return (long)v; }
  explicit operator int() const { return (int)v; }
};
#define SAM_BIN(op, field)                                                                       \
  inline real operator op(real a, real b) { g_sam_ops.field++; return real(a.v op b.v); }         \
  inline real operator op(real a, double b) { g_sam_ops.field++; return real(a.v op b); }         \
  inline real operator op(double a, real b) { g_sam_ops.field++; return real(a op b.v); }         \
  inline real operator op(real a, int b) { g_sam_ops.field++; return real(a.v op (double)b); }    \
  inline real operator op(int a, real b) { g_sam_ops.field++; return real((double)a op b.v); }
SAM_BIN(+, add) SAM_BIN(-, add) SAM_BIN(*, mul) SAM_BIN(/, div)
#undef SAM_BIN
inline real operator-(real a) { return real(-a.v); }
#define SAM_CMP(op)                                                                  \
  inline bool operator op(real a, real b) { g_sam_ops.cmp++; return a.v op b.v; }     \
  inline bool operator op(real a, double b) { g_sam_ops.cmp++; return a.v op b; }     \
  inline bool operator op(double a, real b) { g_sam_ops.cmp++; return a op b.v; }     \
  inline bool operator op(real a, int b) { g_sam_ops.cmp++; return a.v op (double)b; }
SAM_CMP(<) SAM_CMP(>) SAM_CMP(<=) SAM_CMP(>=) SAM_CMP(==) SAM_CMP(!=)
#undef SAM_CMP
inline real fabs(real a) { return real(std::fabs(a.v)); }
inline bool signbit(real a) { return std::signbit(a.v); }
inline real floor(real a) { return real(std::floor(a.v)); }
inline real pow(real a, double b) { g_sam_ops.pw++; return real(std::pow(a.v, b)); }
inline real exp(real a) { g_sam_ops.ex++; return real(std::exp(a.v)); }
inline real sin(real a) { g_sam_ops.sn++; return real(std::sin(a.v)); }
#endif
