/* samsim_oracle.c -- CPU oracle: scalar restatement of the SAMSIM column timestep.
 *
 * TEST INFRASTRUCTURE ONLY (see samsim_oracle.h).  Every routine cites the reference
 * lines it follows (paths relative to /root/reference).  Operation ORDER follows the
 * Fortran source left-to-right so that, compiled with -ffp-contract=off, +,-,*,/ give
 * the values an unoptimised gfortran build gives; the only possible last-bit
 * differences against the real reference are in pow/exp/sin (libm vs libgfortran's use
 * of libm: identical when both link glibc) and in x**2._wp, x**3._wp, x**4._wp, which
 * this file evaluates by pow() in the libm build and by multiplication in the det build.
 *
 * PARITY PINNED: checked against the reference's own golden output (tests/test_oracle_golden.py, fixtures from
 * reference_output/ via tools/make_fixtures.py): testcase 1 -- all 72 records x 90 layers of every per-layer file
 * at print precision, N_active exact, and the four passive-tracer files to their 8 printed decimals; SHEBA --
 * T2m in all digits, T_top (17 digits) to 1e-13..1e-10, melt / flushing / permeability files to 8 digits and
 * N_active exact through record 347 (one full year) with the snow_precip revision the golden run was made with
 * (-DSAM_VARIANT_SNOW_T2M), the later divergence being the model's own 1-ulp sensitivity.
 * The Fortran reference itself cannot be compiled in this image (no Fortran compiler), so there is no oracle/_ref.
 *
 * Two math back-ends, selected at compile time:
 *   default            : glibc libm (what the reference binary calls)
 *   -DSAM_DETMATH      : samsim_b200/csrc/detmath.h (bit-reproducible on the GPU); this
 *                        build is what the CUDA path is compared against bit-for-bit.
 */
#include "samsim_oracle.h"

#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef SAM_DETMATH
#include "../samsim_b200/csrc/detmath.h"
#define M_POW(x, y) det_pow((x), (y))
#define M_EXP(x) det_exp(x)
#define M_SIN(x) det_sin(x)
#define P2(x) ((x) * (x))
static inline real P3(real x) { return (x * x) * x; }
static inline real P4(real x) { real x2 = x * x; return x2 * x2; }
static const char* k_backend = "det";
#elif defined(SAMSIM_COUNT_OPS)
#define M_POW(x, y) pow((x), (y))
#define M_EXP(x) exp(x)
#define M_SIN(x) sin(x)
#define P2(x) ((x) * (x))
static inline real P3(real x) { return (x * x) * x; }
static inline real P4(real x) { real x2 = x * x; return x2 * x2; }
static const char* k_backend = "libm-count";
sam_op_counts g_sam_ops = {0, 0, 0, 0, 0, 0, 0};
extern "C" void sam_get_op_counts(long long* out) {
  out[0] = g_sam_ops.add; out[1] = g_sam_ops.mul; out[2] = g_sam_ops.div; out[3] = g_sam_ops.cmp;
  out[4] = g_sam_ops.pw; out[5] = g_sam_ops.ex; out[6] = g_sam_ops.sn;
}
extern "C" void sam_reset_op_counts(void) { g_sam_ops = sam_op_counts{0, 0, 0, 0, 0, 0, 0}; }
#else
#define M_POW(x, y) pow((x), (y))
#define M_EXP(x) exp(x)
#define M_SIN(x) sin(x)
#define P2(x) pow((x), 2.0)
#define P3(x) pow((x), 3.0)
#define P4(x) pow((x), 4.0)
static const char* k_backend = "libm";
#endif

const char* sam_math_backend(void) { return k_backend; }
double sam_math_pow(double x, double y) { return M_POW(x, y); }
double sam_math_exp(double x) { return M_EXP(x); }
double sam_math_sin(double x) { return M_SIN(x); }

/* ------------------------------------------------------------------------------------------
 * mo_parameters.f90:33-112.  Values tainted by default-REAL literals keep their single
 * precision bit patterns.
 * ---------------------------------------------------------------------------------------- */
#define F32(x) ((double)(float)(x))
static const double pi_sp = (double)3.1415f;  /* :38 REAL, PARAMETER:: pi = 3.1415_wp */
static const double grav = (double)9.8061f;   /* :39 */
static const double k_s = 2.2, k_l = 0.523;   /* :46-47 */
static const double c_s = 2020.0, c_s_beta = 7.6973, c_l = 3400.0;        /* :49-51 */
static const double rho_s = 920.0, rho_l = 1028.0, latent_heat = 333500.0; /* :52-54 */
static const double zeroK = 273.15;                                        /* :55 */
#define bbeta (0.8 * (double)1e-3f)   /* :56 */
#define mu (2.55 * (double)1e-3f)     /* :57 */
#define kappa_l (k_l / rho_l / c_l)   /* :58 */
#define sigma (5.6704 * (double)1e-8f) /* :59 */
static const double psi_s_min = 0.05, neg_free = -0.05;        /* :69-70 */
static const double x_grav = 0.000584, ray_crit = 4.89;        /* :74-75 */
static const double para_flush_horiz = 1.0;                    /* :79 */
static const double para_flush_gamma = 0.9;                    /* :81 */
static const double psi_s_top_min = 0.40;                      /* :83 */
static const double ratio_flood = 1.50;                        /* :85 */
static const double ref_salinity = 34.0;                       /* :87 */
static const double rho_snow = 330.0;                          /* :91 */
static const double gas_snow_ice2 = 0.20;                      /* :93 */
static const double emissivity_ice = 0.95, emissivity_snow = 1.00, penetr = 0.30, extinc = 2.00; /* :96-99 */
#define Turb_A (0.1 * 0.05 * rho_l / 86400.0) /* :102 */
static const double Turb_B = 0.05;              /* :103 */

static inline real r_max(real a, real b) { return (a > b) ? a : b; } /* Fortran MAX */
static inline real r_min(real a, real b) { return (a < b) ? a : b; } /* Fortran MIN */
static inline real r_abs(real a) { return fabs(a); }
/* Fortran SIGN(a,b): |a| with the sign of b */
static inline real r_sign(real a, real b) { return (b >= 0.0 && !signbit(b)) ? fabs(a) : -fabs(a); }

#define EV(c, id) ((c)->ev[SAM_EV_##id]++) /* branch counter, see samsim_oracle.h */

#define SAM_STOP(c, code)     \
  do {                        \
    (c)->status = (code);     \
    longjmp((c)->jb, 1);      \
  } while (0)

/* ==========================================================================================
 * mo_thermo_functions.f90
 * ======================================================================================== */

/* func_S_br without the optional S_bu clamp, mo_thermo_functions.f90:308-351 */
real sam_func_S_br(const sam_col* c, real T) {
  real c1, c2, c3, c4;
  if (c->salt_flag == 1) { /* :321-326 POLY3 seawater */
    c1 = 0.0; c2 = -18.7; c3 = -0.519; c4 = -0.00535;
  } else { /* :331-336 NaCl */
    c1 = 0.0; c2 = -17.6; c3 = -0.389; c4 = -0.00362;
  }
  return c1 + c2 * T + c3 * P2(T) + c4 * P3(T); /* :340 */
}

/* func_S_br with S_bu present, :353-357 */
real sam_func_S_br2(const sam_col* c, real T, real S_bu) {
  real S_br = sam_func_S_br(c, T);
  if (S_br < S_bu) S_br = S_bu;
  return S_br;
}

/* func_ddT_S_br, mo_thermo_functions.f90:380-414 (old seawater coefficients, clamp below -20) */
real sam_func_ddT_S_br(const sam_col* c, real T) {
  real c2, c3, c4, T_crit = -20.0, d;
  if (c->salt_flag == 1) { /* :393-397 */
    c2 = -21.4; c3 = -0.886; c4 = -0.0170;
  } else { /* :398-402 */
    c2 = -17.6; c3 = -0.389; c4 = -0.00362;
  }
  d = c2 + 2.0 * c3 * T + 3.0 * c4 * P2(T); /* :406 */
  if (T < T_crit) d = c2 + 2.0 * c3 * T_crit + 3.0 * c4 * P2(T_crit); /* :408-412 */
  return d;
}

/* getT, mo_thermo_functions.f90:62-143.  T_in is passed BY VALUE here; the aliasing call
 * sites in snow_coupling pass H/c_l explicitly (see snow_coupling below). */
void sam_getT(sam_col* c, real H, real S_bu, real T_in, real* T_out, real* phi_out, int k) {
  real T, phi = *phi_out, T_0, f, ddT_f, T_fr;
  int i;
  (void)k;
  c->stat_getT_calls++;
  T = H / c_l; /* :80 */
  if (sam_func_S_br2(c, T, S_bu) > S_bu && S_bu > 0.001) { /* :82 */
    T_fr = -1.0; /* :85 */
    while (r_abs(sam_func_S_br(c, T_fr) / S_bu - 1.0) > F32(0.0001)) { /* :87 */
      T_0 = T_fr;
      f = sam_func_S_br(c, T_0) - S_bu;
      ddT_f = sam_func_ddT_S_br(c, T_0);
      T_fr = T_0 - f / ddT_f;
      c->stat_newton_fr++;
    }
    T_0 = T_in; /* :94 */
    f = -latent_heat - H + latent_heat * S_bu / r_max(sam_func_S_br(c, T_0), 0.000000001) + c_s * T_0 +
        c_s_beta * T_0 * T_0 / 2.0; /* :95 */
    ddT_f = c_s + c_s_beta * T_0 -
            latent_heat * S_bu * sam_func_ddT_S_br(c, T_0) / r_max(P2(sam_func_S_br(c, T_0)), 0.0000000001); /* :96 */
    T = T_0 - f / ddT_f; /* :97 */
    i = 0;
    while (r_abs(f) > 1.0) { /* :99 ABS(f)>1_wp */
      T_0 = T;
      if (T_0 > 0.0 || T_0 < -200.0) { T_0 = T_fr; EV(c, GETT_TFR_FALLBACK); } /* :101-103 */
      f = -latent_heat - H + latent_heat * S_bu / r_max(sam_func_S_br(c, T_0), 0.0000000001) + c_s * T_0 +
          c_s_beta * T_0 * T_0 / 2.0; /* :104 */
      {
        real sb = sam_func_S_br(c, T_0);
        ddT_f = c_s + c_s_beta * T_0 - latent_heat * S_bu * sam_func_ddT_S_br(c, T_0) / r_max(sb * sb, 0.0000000001); /* :105 (**2 integer power) */
      }
      T = T_0 - f / ddT_f; /* :106 */
      i = i + 1;
      c->stat_newton_T++;
      if (i == 260) SAM_STOP(c, 99); /* :114-123 */
    }
    phi = 1.0 - S_bu / sam_func_S_br2(c, T, S_bu); /* :125 */
  } else if (S_bu < 0.001) { /* :127 */
    EV(c, GETT_SALTFREE);
    if (H > 0.0) {
      phi = 0.0;
      T = H / c_l;
    } else if (H <= -latent_heat) {
      phi = 1.0;
      T = (H + latent_heat) / c_s;
    } else if (H <= 0.0 && -latent_heat < H) {
      T = 0.0;
      phi = -H / latent_heat;
    }
  } else {
    EV(c, GETT_LIQUID);
    phi = 0.0; /* :139 */
  }
  *T_out = T;
  *phi_out = phi;
}

/* Expulsion, mo_thermo_functions.f90:157-187 */
static void Expulsion(real phi, real thick, real m, real* psi_s, real* psi_l, real* psi_g, real* V_ex) {
  real V_s = m * phi / rho_s;
  real V_l = m * (1.0 - phi) / rho_l;
  if (V_s + V_l > thick) {
    *V_ex = V_l + V_s - thick;
  } else {
    *V_ex = 0.0;
  }
  *psi_s = V_s / thick;
  *psi_l = (V_l - *V_ex) / thick;
  *psi_g = (thick - V_l - V_s + *V_ex) / thick;
  if (*psi_l < 0.0) *psi_l = 0.0;
  if (*psi_g < 0) *psi_g = 0.0;
}

/* sub_fl_Q, mo_thermo_functions.f90:201-224 */
static real sub_fl_Q(real psi_s_1, real psi_l_1, real psi_g_1, real thick_1, real T_1, real psi_s_2, real psi_l_2,
                     real psi_g_2, real thick_2, real T_2) {
  real k_1 = psi_s_1 * k_s + psi_l_1 * k_l + psi_g_1 * 0.0;
  real k_2 = psi_s_2 * k_s + psi_l_2 * k_l + psi_g_2 * 0.0;
  real R = thick_1 / (2.0 * k_1) + thick_2 / (2.0 * k_2);
  return (T_2 - T_1) / R;
}

/* sub_fl_Q_0, mo_thermo_functions.f90:238-265 */
static real sub_fl_Q_0(sam_col* c, real psi_s, real psi_l, real psi_g, real thick, real T, real T_bound,
                       int direct_flag) {
  real k = psi_s * k_s + psi_l * k_l + psi_g * 0.0;
  real R = thick / (2.0 * k);
  if (direct_flag == 1) return (T_bound - T) / R;
  if (direct_flag == -1) return (T - T_bound) / R;
  SAM_STOP(c, 98);
  return 0.0;
}

/* ==========================================================================================
 * mo_functions.f90 (physics part)
 * ======================================================================================== */

/* func_density, mo_functions.f90:51-62 */
real sam_func_density(real T, real S) {
  real density_0 = 999.842594 + 6.8 / 100.0 * T;
  real A = 0.825;
  real B = -5.7 / 1000.0;
  return density_0 + A * S + B * M_POW(r_max(S, 0.0), 1.5);
}

/* func_freeboard, mo_functions.f90:79-130.  SUM(a(i:j)*b(i:j)) is a forward loop from i. */
static real sum_prod(const real* a, const real* b, int i, int j) {
  real s = 0.0;
  int q;
  for (q = i; q <= j; q++) s = s + a[q] * b[q];
  return s;
}
static real sum_arr(const real* a, int i, int j) {
  real s = 0.0;
  int q;
  for (q = i; q <= j; q++) s = s + a[q];
  return s;
}

real sam_func_freeboard(const sam_col* c) {
  const int N_active = c->N_active;
  const real *psi_s = c->psi_s, *psi_g = c->psi_g, *m = c->m, *thick = c->thick;
  real snowmass, freeboard, test1, test2;
  int k;
  if (c->freeboard_snow_flag == 0) snowmass = c->m_snow; else snowmass = 0.0; /* :92-96 */
  if (snowmass > sum_prod(psi_s, thick, 1, N_active) * (rho_l - rho_s) + sum_prod(psi_g, thick, 1, N_active) * rho_l) { /* :99 */
    test2 = sum_prod(psi_s, thick, 1, N_active) * (rho_l - rho_s) + sum_prod(psi_g, thick, 1, N_active) * rho_l;
    freeboard = test2 - snowmass;
    freeboard = freeboard / rho_l;
  } else {
    test1 = 0.0;
    test2 = 1.0;
    k = 0;
    while (test1 < test2) { /* :114-118 */
      k = k + 1;
      test2 = sum_prod(psi_s, thick, k + 1, N_active) * (rho_l - rho_s) + sum_prod(psi_g, thick, k + 1, N_active) * rho_l;
      test1 = sum_arr(m, 1, k) + snowmass;
    }
    test1 = sum_arr(m, 1, k - 1) + snowmass; /* :121 */
    freeboard = test2 - test1 + (rho_l - m[k] / thick[k]) * thick[k]; /* :124 */
    freeboard = freeboard / rho_l;
    freeboard = freeboard + sum_arr(thick, 1, k - 1);
  }
  return freeboard;
}

/* func_albedo, mo_functions.f90:157-208.  Locals are REAL(wp) assigned from default-REAL literals. */
real sam_func_albedo(real thick_snow, real T_snow, real psi_l, real thick_min, int albedo_flag) {
  real albedo;
  const real ice_dry = F32(0.75), ice_wet = F32(0.6), snow_dry = F32(0.85), snow_wet = F32(0.75), water = F32(0.2);
  if (thick_snow > thick_min) {
    if (T_snow < F32(-0.01)) albedo = snow_dry; else albedo = snow_wet;
    albedo = ice_dry + (albedo - ice_dry) * r_min(1.0, thick_snow / 0.3); /* :177 */
  } else {
    if (psi_l > 0.9) {
      albedo = water;
    } else if (psi_l > 0.6) {
      albedo = ice_wet + (water - ice_wet) * ((psi_l - 0.6) / 0.3);
    } else if (psi_l > 0.2) {
      albedo = ice_wet;
    } else {
      albedo = ice_dry;
    }
  }
  if (albedo_flag == 1) { /* :191-205 */
    if (thick_snow > thick_min) {
      if (T_snow < F32(-0.01)) albedo = snow_dry; else albedo = snow_wet;
    } else {
      if (psi_l < F32(0.8)) albedo = ice_dry; else albedo = water;
    }
  }
  return albedo;
}

/* func_T_freeze, mo_functions.f90:239-250.  Default-REAL literals and single*single products. */
real sam_func_T_freeze(real S_bu, int salt_flag) {
  real T_freeze = 0.0;
  if (salt_flag == 2) {
    /* -0.0592_wp*S_bu -9.37*S_bu**2.0 -5.33*10.0**(-7.0)*S_bu**3.0 */
    const double c2 = (double)9.37f;
    const double c3 = (double)(5.33f * 1e-7f); /* single*single: 10.0**(-7.0) folds to the single 1e-7 */
    T_freeze = -0.0592 * S_bu - c2 * P2(S_bu) - c3 * P3(S_bu);
  } else if (salt_flag == 1) {
    /* -0.0575_wp*S_bu +1.710523*1e-3*S_bu**1.5 -2.154996*1e-4*S_bu**2.0 */
    const double a = (double)(1.710523f * 1e-3f);
    const double b = (double)(2.154996f * 1e-4f);
    T_freeze = -0.0575 * S_bu + a * M_POW(S_bu, 1.5) - b * P2(S_bu);
  }
  return T_freeze;
}

/* sub_notzflux, mo_functions.f90:270-289 */
static void sub_notzflux(real time, real* fl_sw, real* fl_rest) {
  real day = time / 86400.0;
  while (day > 360) day = day - 360;
  *fl_sw = 314.0 * M_EXP(-0.5 * P2((day - 164.0) / F32(47.9)));
  *fl_rest = 118.0 * M_EXP(-0.5 * P2((day - 206.0) / F32(53.1))) + 179.0;
  if (day < 60. || day > 300.) *fl_sw = 0.0;
}

/* sub_turb_flux, mo_functions.f90:347-363 */
static void sub_turb_flux(real T_bottom, real S_bu_bottom, real T, real* S_abs, real m, real dt) {
  real turb = Turb_A * M_EXP(Turb_B * (-sam_func_density(T_bottom, S_bu_bottom) + sam_func_density(T, *S_abs / m))) * dt;
  *S_abs = *S_abs - turb * (*S_abs / m - S_bu_bottom);
}

/* sub_melt_thick, mo_functions.f90:386-428 */
/* returns 1 when the gas-fraction correction (:418-426) ran (branch counter only) */
static int sub_melt_thick(real psi_l, real psi_s, real psi_g, real T, real T_freeze, real T_top, real fl_Q,
                          real thick_snow, real dt, real* melt_thick, real* thick, real thick_min) {
  int gas = 0;
  *melt_thick = 0.0;
  if (thick_snow < thick_min && T_top >= T_freeze) { /* :396 */
    *melt_thick = -fl_Q - 2.0 * (psi_l * k_l + psi_s * k_s) / *thick * (T_freeze - T);
    *melt_thick = *melt_thick * dt / r_max((latent_heat * rho_s * psi_s), 0.000000000000001);
    *melt_thick = r_min(psi_l * *thick, *melt_thick);
  }
  if (psi_s < psi_s_top_min) { /* :412 */
    *melt_thick = *thick * (1.0 - psi_s / psi_s_top_min);
  }
  if (*melt_thick > 0.0 && psi_g > gas_snow_ice2) { /* :418 */
    gas = 1;
    if (*melt_thick > (psi_g - gas_snow_ice2) * *thick) {
      *melt_thick = *melt_thick - (psi_g - gas_snow_ice2) * *thick;
      *thick = *thick * (1.0 - (psi_g - gas_snow_ice2));
    } else {
      *thick = *thick - *melt_thick;
      *melt_thick = 0.0;
    }
  }
  return gas;
}

/* sub_melt_snow, mo_functions.f90:443-474; returns 1 when all the snow went into the ice (:453), 0 for the partial branch */
static int sub_melt_snow(real* melt_thick, real* thick, real* thick_snow, real* H_abs, real* H_abs_snow, real* m,
                          real* m_snow, real* psi_g_snow) {
  real shift = 1.0 / r_max(*psi_g_snow, 0.01) * *melt_thick;
  if (shift >= *thick_snow) {
    *melt_thick = *melt_thick - *thick_snow * *psi_g_snow;
    *H_abs = *H_abs + *H_abs_snow;
    *m = *m + *m_snow;
    *thick = *thick + (1.0 - *psi_g_snow) * *thick_snow;
    *thick_snow = 0.0;
    *m_snow = 0.0;
    *H_abs_snow = 0.0;
    return 1;
  } else {
    *H_abs = *H_abs + shift / *thick_snow * *H_abs_snow;
    *H_abs_snow = *H_abs_snow - shift / *thick_snow * *H_abs_snow;
    *m = *m + shift / *thick_snow * *m_snow;
    *m_snow = *m_snow - shift / *thick_snow * *m_snow;
    *thick = *thick + shift - *melt_thick;
    *thick_snow = *thick_snow - shift;
    *melt_thick = 0.0;
  }
  return 0;
}

/* ==========================================================================================
 * mo_mass.f90
 * ======================================================================================== */

/* mass_transfer, mo_mass.f90:53-96.  `0.` comparisons are single zeros (exact). */
static void mass_transfer(sam_col* c, const real* T, real* H_abs, real* S_abs, const real* S_bu, const real* fl_m) {
  const int N_active = c->N_active;
  real* TT = c->scr[0];
  real* SS_bu = c->scr[1];
  real* SS_abs = c->scr[2];
  int k;
  for (k = 1; k <= N_active; k++) { /* :66-68 */
    TT[k] = T[k];
    SS_bu[k] = S_bu[k];
    SS_abs[k] = S_abs[k];
  }
  TT[N_active + 1] = c->T_bottom; /* :70-72 */
  SS_bu[N_active + 1] = c->S_bu_bottom;
  SS_abs[N_active + 1] = c->S_bu_bottom * 2000.0;
  TT[0] = 0.0; SS_bu[0] = 0.0; /* never read: fl_m(1) is 0 at every call site */
  for (k = 1; k <= N_active; k++) { /* :76-95 */
    if (fl_m[k + 1] > 0.) {
      H_abs[k] = H_abs[k] + fl_m[k + 1] * TT[k + 1] * c_l;
      S_abs[k] = S_abs[k] + r_min(fl_m[k + 1] * sam_func_S_br2(c, TT[k + 1], SS_bu[k + 1]), SS_abs[k + 1]);
    } else if (fl_m[k + 1] < 0.) {
      H_abs[k] = H_abs[k] + fl_m[k + 1] * TT[k] * c_l;
      S_abs[k] = S_abs[k] + r_max(fl_m[k + 1] * sam_func_S_br2(c, TT[k], SS_bu[k]), -S_abs[k]);
    }
    if (fl_m[k] > 0.) {
      H_abs[k] = H_abs[k] - fl_m[k] * TT[k] * c_l;
      S_abs[k] = S_abs[k] - r_min(fl_m[k] * sam_func_S_br2(c, TT[k], SS_bu[k]), S_abs[k]);
    } else if (fl_m[k] < 0) {
      H_abs[k] = H_abs[k] - fl_m[k] * TT[k - 1] * c_l;
      S_abs[k] = S_abs[k] - r_max(fl_m[k] * sam_func_S_br2(c, TT[k - 1], SS_bu[k - 1]), -S_abs[k - 1]);
    }
  }
}

/* expulsion_flux, mo_mass.f90:112-136.  `0.001` is a single literal. */
static void expulsion_flux(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *fl_m = c->fl_m, *V_ex = c->V_ex, *psi_g = c->psi_g, *thick = c->thick, *m = c->m;
  int k;
  for (k = 1; k <= Nlayer + 1; k++) fl_m[k] = 0.0;
  fl_m[2] = -V_ex[1] * rho_l;
  for (k = 2; k <= N_active; k++) {
    if (psi_g[k] < F32(0.001)) {
      fl_m[k + 1] = -V_ex[k] * rho_l + fl_m[k];
    } else {
      fl_m[k + 1] = -r_max((V_ex[k] - psi_g[k] * thick[k]) * rho_l, 0.0);
      psi_g[k] = r_max((psi_g[k] * thick[k] - V_ex[k]) / thick[k], 0.0);
    }
  }
  for (k = 1; k <= N_active; k++) m[k] = m[k] + fl_m[k + 1] - fl_m[k];
}

/* ==========================================================================================
 * mo_grav_drain.f90
 * ======================================================================================== */

/* fl_grav_drain, mo_grav_drain.f90:74-201 (bgc bookkeeping omitted: no feedback on H,S,m) */
/* fl_brine_bgc(i,j): brine moved from layer i to layer j this step (i or j = N_active+1: the ocean) */
#define FB(c, i, j) ((c)->fl_brine_bgc[(size_t)(i) * ((c)->Nlayer + 2) + (j)])

/* bgc_advection, mo_mass.f90:150-209: upwind advection of the tracers with the recorded brine fluxes; each flux
 * is limited to a third of the layer's content */
static void bgc_advection(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer, N_bgc = c->N_bgc;
  real *psi_l = c->psi_l, *thick = c->thick;
  int i, j, k;
  for (k = 1; k <= N_bgc; k++) {
    real* bgc_abs = c->bgc_abs[k];
    real* bgc_temp = c->scr[3];
    real* bgc_br = c->scr[4];
    real flux;
    for (i = 1; i <= Nlayer; i++) bgc_temp[i] = bgc_abs[i];
    for (i = 1; i <= N_active; i++) bgc_br[i] = bgc_abs[i] / (r_max(psi_l[i] * thick[i] * rho_l, 0.000000000000001)); /* :171 */
    /* tracers are independent: the reference's (i,j,k) nest visits, for each tracer, the cells in the same (i,j) order */
    for (i = 1; i <= N_active; i++) { /* :179-190 internal flows */
      for (j = 1; j <= N_active; j++) {
        /* an empty cell moves MIN(0*br, abs/3) = 0 unless the layer's content is negative: x -/+ 0 leaves x as it is */
        if (FB(c, i, j) == 0.0 && bgc_abs[i] >= 0.0 && bgc_br[i] == bgc_br[i] && bgc_br[i] - bgc_br[i] == 0.0) continue;
        flux = r_min(FB(c, i, j) * bgc_br[i], bgc_abs[i] / 3.0);
        bgc_temp[i] = bgc_temp[i] - flux;
        bgc_temp[j] = bgc_temp[j] + flux;
      }
    }
    for (i = 1; i <= N_active; i++) { /* :193-199 flows which leave the domain */
      flux = r_min(FB(c, i, N_active + 1) * bgc_br[i], bgc_abs[i] / 3.0);
      bgc_temp[i] = bgc_temp[i] - flux;
    }
    for (j = 1; j <= N_active; j++) { /* :202-208 flows which enter the domain */
      flux = FB(c, N_active + 1, j) * c->bgc_bottom[k];
      bgc_temp[j] = bgc_temp[j] + flux;
    }
    for (i = 1; i <= Nlayer; i++) bgc_abs[i] = bgc_temp[i];
  }
}

static void fl_grav_drain(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *S_br = c->S_br, *S_bu = c->S_bu, *psi_l = c->psi_l, *psi_s = c->psi_s, *thick = c->thick, *S_abs = c->S_abs,
       *H_abs = c->H_abs, *T = c->T, *m = c->m, *ray = c->ray;
  const real dt = c->dt;
  real* fl_up = c->scr[3];
  real* fl_down = c->scr[4];
  real* perm = c->scr[5];
  real* harmonic_perm = c->scr[6];
  real* fl_m = c->scr[7];
  real flux, test1, d_S_br, height, ray_mini, heat_loss;
  int k, kk;

  ray_mini = ray_crit; /* :94 */
  heat_loss = 0.0;
  for (k = 0; k <= Nlayer + 2; k++) { /* :96-101 */
    perm[k] = 0.0; fl_up[k] = 0.0; fl_down[k] = 0.0; harmonic_perm[k] = 0.0; fl_m[k] = 0.0;
  }
  perm[N_active] = 9999999.0; /* :97, overwritten below */
  for (k = 1; k <= Nlayer - 1; k++) ray[k] = 0.0;

  for (k = 1; k <= N_active; k++) { /* :104-106 */
    perm[k] = 1e-17 * M_POW(1000.0 * r_abs(psi_l[k]), 3.10);
  }

  if (c->harmonic_flag == 2) { /* :109-123 */
    for (k = 1; k <= N_active - 1; k++) {
      test1 = perm[k];
      for (kk = k; kk <= N_active - 1; kk++) test1 = r_min(test1, perm[kk]); /* minval(perm(k:N_active-1)) */
      if (test1 < 1e-14) {
        harmonic_perm[k] = 0.0;
      } else {
        for (kk = k; kk <= N_active - 1; kk++) harmonic_perm[k] = harmonic_perm[k] + thick[kk] / perm[kk];
        harmonic_perm[k] = harmonic_perm[k] + (thick[N_active] * psi_s[N_active] / psi_s_min) / perm[N_active];
        harmonic_perm[k] = (sum_arr(thick, k, N_active - 1) + thick[N_active] * psi_s[N_active] / psi_s_min) / harmonic_perm[k];
      }
    }
  }

  for (k = 1; k <= N_active - 1; k++) { /* :126-136 */
    d_S_br = S_br[k] - S_br[N_active];
    height = sum_arr(thick, k + 1, N_active - 1) + thick[N_active] * psi_s[N_active] / psi_s_min;
    if (c->harmonic_flag == 1) {
      real mn = perm[k];
      for (kk = k; kk <= N_active; kk++) mn = r_min(mn, perm[kk]);
      ray[k] = grav * rho_l * bbeta * d_S_br * height * mn;
    } else if (c->harmonic_flag == 2) {
      ray[k] = grav * rho_l * bbeta * d_S_br * height * harmonic_perm[k];
    }
    ray[k] = ray[k] / (kappa_l * mu);
    ray[k] = r_max(ray[k], 0.0);
  }

  c->grav_salt = c->grav_salt + sum_arr(S_abs, 1, Nlayer); /* :141 SUM(S_abs(:)) */

  for (k = 1; k <= N_active - 1; k++) { /* :144-171 */
    if (ray[k] > ray_mini && psi_s[k] > 0.001 && S_abs[k] / m[k] > 0.1 && S_br[k] > S_br[k + 1]) {
      EV(c, GRAV_DRAINED);
      flux = x_grav * (ray[k] - ray_mini) * dt * thick[k];
      flux = r_min(flux, psi_l[k] * rho_l * thick[k]);
      S_abs[k] = S_abs[k] - flux * S_br[k];
      if (S_abs[k] < 0.0) SAM_STOP(c, 21234); /* :149-153 */
      c->grav_temp = c->grav_temp + flux * T[k];
      H_abs[k] = H_abs[k] - flux * c_l * T[k];
      heat_loss = heat_loss + flux * c_l * T[k];
      fl_down[k] = flux;
      for (kk = k; kk <= N_active; kk++) fl_up[kk] = fl_up[kk] + flux;
      fl_up[k] = r_min(fl_up[k], psi_l[k] * rho_l * thick[k]);
    }
  }

  c->grav_salt = c->grav_salt - sum_arr(S_abs, 1, Nlayer); /* :173 */

  fl_m[1] = 0.0; /* :176-177 */
  for (k = 1; k <= N_active; k++) fl_m[k + 1] = fl_up[k];

  if (c->bgc_flag == 2) { /* :178-185 (sic: column N_active+1 is assigned from column N_active) */
    for (k = 1; k <= N_active - 1; k++) FB(c, k, N_active + 1) = FB(c, k, N_active) + fl_down[k];
    for (k = 1; k <= N_active; k++) FB(c, k + 1, k) = FB(c, k + 1, k) + fl_up[k];
  }

  mass_transfer(c, T, H_abs, S_abs, S_bu, fl_m); /* :188 */

  c->grav_drain = c->grav_drain + fl_m[N_active + 1]; /* :190 */

  if (c->grav_heat_flag == 2) { /* :193-195 */
    H_abs[N_active] = H_abs[N_active] + heat_loss - fl_up[N_active] * c_l * c->T_bottom;
  }
  {
    real mn = S_abs[1]; /* :198 MINVAL(S_abs) over all Nlayer */
    for (k = 1; k <= Nlayer; k++) mn = r_min(mn, S_abs[k]);
    if (mn < 0.0) SAM_STOP(c, 1337);
  }
}

/* fl_grav_drain_simple, mo_grav_drain.f90:218-279 (grav_flag 3; not used by the configs) */
static void fl_grav_drain_simple(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *S_br = c->S_br, *psi_l = c->psi_l, *psi_s = c->psi_s, *thick = c->thick, *S_abs = c->S_abs, *ray = c->ray;
  real* perm = c->scr[5];
  real* harmonic_perm = c->scr[6];
  real d_S_br, height, temp;
  int k, kk;
  for (k = 0; k <= Nlayer + 2; k++) { perm[k] = 0.0; harmonic_perm[k] = 0.0; }
  perm[N_active] = 9999999.0;
  for (k = 1; k <= Nlayer - 1; k++) ray[k] = 0.0;
  for (k = 1; k <= N_active; k++) perm[k] = 1e-17 * M_POW(1000.0 * r_abs(psi_l[k]), 3.10);
  if (c->harmonic_flag == 2) {
    for (k = 1; k <= N_active - 1; k++) {
      temp = perm[k];
      for (kk = k; kk <= N_active - 1; kk++) temp = r_min(temp, perm[kk]);
      if (temp < 1e-14) {
        harmonic_perm[k] = 0.0;
      } else {
        for (kk = k; kk <= N_active - 1; kk++) harmonic_perm[k] = harmonic_perm[k] + thick[kk] / perm[kk];
        harmonic_perm[k] = harmonic_perm[k] + (thick[N_active] * psi_s[N_active] / psi_s_min) / perm[N_active];
        harmonic_perm[k] = (sum_arr(thick, k, N_active - 1) + thick[N_active] * psi_s[N_active] / psi_s_min) / harmonic_perm[k];
      }
    }
  }
  for (k = 1; k <= N_active - 1; k++) {
    d_S_br = S_br[k] - S_br[N_active];
    height = sum_arr(thick, k + 1, N_active - 1) + thick[N_active] * psi_s[N_active] / psi_s_min;
    if (c->harmonic_flag == 1) {
      real mn = perm[k];
      for (kk = k; kk <= N_active; kk++) mn = r_min(mn, perm[kk]);
      ray[k] = grav * rho_l * bbeta * d_S_br * height * mn;
    } else if (c->harmonic_flag == 2) {
      ray[k] = grav * rho_l * bbeta * d_S_br * height * harmonic_perm[k];
    }
    ray[k] = ray[k] / (kappa_l * mu);
    ray[k] = r_max(ray[k], 0.0);
  }
  for (k = N_active - 1; k >= 1; k--) {
    if (ray[k] > ray_crit) S_abs[k] = S_abs[k] * F32(0.99); /* :274 */
  }
  c->grav_drain = 0.0;
}

/* ==========================================================================================
 * mo_snow.f90
 * ======================================================================================== */

/* func_k_snow, mo_snow.f90:560-573 */
real sam_func_k_snow(real m_snow, real thick_snow) {
  const real c0 = 0.138, c1 = -1.01 / 1000.0, c2 = 3.233 / 1000000.0;
  real k_snow = c0 + c1 * m_snow / thick_snow + c2 * P2(m_snow / thick_snow);
  k_snow = k_snow + F32(0.15);
  return k_snow;
}

/* snow_coupling, mo_snow.f90:61-104.  The reference passes the same variable as T_in and T
 * (`CALL getT(H, S_bu, T, T, phi, 5702)`); getT assigns T = H/c_l (:80) before it reads T_in
 * (:94), so with by-reference scalars the first guess is H/c_l.  Reproduced explicitly. */
static void getT_aliased(sam_col* c, real H, real S_bu, real* T, real* phi, int k) {
#ifdef SAM_VARIANT_COUPLING_NOALIAS /* hypothesis test only: T_in keeps the caller's old T (copy-in semantics) */
  sam_getT(c, H, S_bu, *T, T, phi, k);
#else
  sam_getT(c, H, S_bu, H / c_l, T, phi, k);
#endif
}

static void snow_coupling(sam_col* c, real* H_abs_snow, real* phi_s, real* T_snow, real* H_abs, real* H, real* phi,
                          real* T, real m_snow, real S_abs_snow, real m, real S_bu) {
  int jj;
  *H_abs = *H_abs + m_snow * latent_heat + *H_abs_snow; /* :69 */
  *H_abs_snow = -m_snow * latent_heat;
  *H = *H_abs / m;
  getT_aliased(c, *H_abs_snow / m_snow, S_abs_snow / m_snow, T_snow, phi_s, 5701); /* :73 */
  getT_aliased(c, *H, S_bu, T, phi, 5702);                                         /* :74 */
  if (*T > 0 && *H_abs <= -*H_abs_snow) { /* :76 */
    EV(c, SNOW_COUPLING_WARM1);
    *H_abs_snow = *H_abs_snow + *H_abs;
    *H_abs = 0.0;
    getT_aliased(c, *H_abs_snow / m_snow, S_abs_snow / m_snow, T_snow, phi_s, 5701);
    getT_aliased(c, *H, S_bu, T, phi, 5702);
  } else if (*T > 0. && *H_abs > -*H_abs_snow) { /* :81 */
    EV(c, SNOW_COUPLING_WARM2);
    *H_abs = (*H_abs + *H_abs_snow) * m / m_snow / (1.0 + m / m_snow);
    *H_abs_snow = *H_abs * m_snow / m;
    getT_aliased(c, *H_abs_snow / m_snow, S_abs_snow / m_snow, T_snow, phi_s, 5701);
    getT_aliased(c, *H, S_bu, T, phi, 5702);
  } else {
    jj = 0;
    while (r_abs(*T - *T_snow) > F32(0.1) && jj < 201) { /* :88 */
      real d = *T_snow - (*T_snow + *T) / 2.0;
      *H_abs_snow = *H_abs_snow - r_sign(r_max(r_abs(d), 0.1), d) * c_s * m_snow; /* :89 */
      d = *T_snow - (*T_snow + *T) / 2.0;
      *H_abs = *H_abs + r_sign(r_max(r_abs(d), 0.1), d) * c_s * m_snow; /* :90 */
      jj = jj + 1;
      *H = *H_abs / m;
      c->stat_coupling_iters++;
      EV(c, SNOW_COUPLING_ITER);
      getT_aliased(c, *H_abs_snow / m_snow, S_abs_snow / m_snow, T_snow, phi_s, 5701);
      getT_aliased(c, *H, S_bu, T, phi, 5702);
    }
    if (jj > 200 && r_abs(*T - *T_snow) > 1.0) SAM_STOP(c, 16); /* :99-102 */
  }
}

/* snow_precip, mo_snow.f90:123-150 */
static void snow_precip(real* m_snow, real* H_abs_snow, real* thick_snow, real dt, real liquid_precip_in, real T2m,
                        int have_solid, real solid_precip_in) {
  real solid_precip, liquid_precip, d_thick;
  if (have_solid) {
    solid_precip = solid_precip_in;
    liquid_precip = liquid_precip_in;
  } else {
    if (T2m > 0.0) {
      solid_precip = 0.0;
      liquid_precip = liquid_precip_in;
    } else {
      solid_precip = liquid_precip_in;
      liquid_precip = 0.0;
    }
  }
  d_thick = dt * solid_precip * rho_l / rho_snow;
  *m_snow = *m_snow + dt * rho_l * (liquid_precip + solid_precip);
  *thick_snow = *thick_snow + d_thick;
  *H_abs_snow = *H_abs_snow + dt * T2m * liquid_precip * rho_l * c_l;
#ifdef SAM_VARIANT_SNOW_T2M /* hypothesis test only: "replaced with T2m" without the min(.,-1) */
  *H_abs_snow = *H_abs_snow + dt * T2m * solid_precip * rho_l * c_s;
#else
  *H_abs_snow = *H_abs_snow + dt * r_min(T2m, -1.0) * solid_precip * rho_l * c_s;
#endif
  *H_abs_snow = *H_abs_snow - dt * solid_precip * rho_l * latent_heat;
}

/* snow_precip_0, mo_snow.f90:167-192 */
static void snow_precip_0(real* H_abs, real* S_abs, real m, real T, real dt, real liquid_precip_in, real T2m,
                          int have_solid, real solid_precip_in) {
  real solid_precip, liquid_precip;
  if (have_solid) {
    solid_precip = solid_precip_in;
    liquid_precip = liquid_precip_in;
  } else {
    if (T2m > 0.0) {
      solid_precip = 0.0;
      liquid_precip = liquid_precip_in;
    } else {
      solid_precip = liquid_precip_in;
      liquid_precip = 0.0;
    }
  }
  *H_abs = *H_abs + (liquid_precip + solid_precip) * (T2m - T) * dt;
  *H_abs = *H_abs - solid_precip * latent_heat * dt;
  *S_abs = *S_abs - (liquid_precip + solid_precip) * *S_abs / m * dt;
}

/* snow_thermo (meltwater==0), mo_snow.f90:212-319, and snow_thermo_meltwater (meltwater==1), :331-458 */
static void snow_thermo_any(sam_col* c, int meltwater, real* m, real* thick, real* H_abs, real* melt_thick_snow) {
  real sat_snow, psi_s_old, phi_snow = 0.0, T_in, H_snow, S_bu_snow, max_lwc, max_lwc_v;
  real psi_l_snow_slush, psi_l_snow_flush;
  real *psi_l_snow = &c->psi_l_snow, *psi_s_snow = &c->psi_s_snow, *psi_g_snow = &c->psi_g_snow,
       *thick_snow = &c->thick_snow, *S_abs_snow = &c->S_abs_snow, *H_abs_snow = &c->H_abs_snow,
       *m_snow = &c->m_snow, *T_snow = &c->T_snow;

  H_snow = *H_abs_snow / *m_snow;
  S_bu_snow = *S_abs_snow / *m_snow;
  psi_s_old = *psi_s_snow;

  if (meltwater) EV(c, SNOW_THERMO_MELTWATER); else EV(c, SNOW_THERMO);
  T_in = *T_snow;
  sam_getT(c, H_snow, S_bu_snow, T_in, T_snow, &phi_snow, 5700);

  *psi_s_snow = *m_snow * phi_snow / rho_s / *thick_snow;
  *psi_l_snow = *m_snow * (1.0 - phi_snow) / rho_l / *thick_snow;
  if (*psi_s_snow + *psi_l_snow > 1.0) {
    *thick_snow = *m_snow * (phi_snow / rho_s + (1.0 - phi_snow) / rho_l);
    *psi_s_snow = *m_snow * phi_snow / rho_s / *thick_snow;
    *psi_l_snow = *m_snow * (1.0 - phi_snow) / rho_l / *thick_snow;
    if (r_abs(*psi_s_snow + *psi_l_snow - 1.0) > 0.0000001) SAM_STOP(c, 345);
  }

  *psi_g_snow = 1.0 - *psi_s_snow - *psi_l_snow;
  if (*psi_s_snow > 0.0) {
    max_lwc = 0.057 * (1.0 - *psi_s_snow) / (*psi_s_snow) + 0.017;
  } else {
    max_lwc = 0.0;
  }

  if (psi_s_old > *psi_s_snow && *psi_s_snow > 0.0) {
    EV(c, SNOW_COMPACTION);
    if ((1.0 - phi_snow) > max_lwc) {
      *thick_snow = *thick_snow * (1.0 - (psi_s_old - *psi_s_snow) / psi_s_old);
    }
    if (*thick_snow < (phi_snow * *m_snow / rho_s + (1.0 - phi_snow) * *m_snow / rho_l)) {
      *thick_snow = (phi_snow * *m_snow / rho_s + (1.0 - phi_snow) * *m_snow / rho_l);
    }
    *psi_s_snow = *m_snow * phi_snow / rho_s / *thick_snow;
    *psi_l_snow = *m_snow * (1.0 - phi_snow) / rho_l / *thick_snow;
    *psi_g_snow = 1.0 - *psi_s_snow - *psi_l_snow;
    *psi_g_snow = r_abs(*psi_g_snow);
  } else if (*psi_s_snow < 0.000001) {
    *thick_snow = *m_snow / rho_l;
    *psi_s_snow = 0.0;
    *psi_g_snow = 0.0;
    *psi_l_snow = 1.0;
  }

  if (!meltwater) {
    /* mo_snow.f90:271-309 */
    if ((1.0 - phi_snow) > max_lwc && *psi_g_snow > 0.0) {
      EV(c, SNOW_WET);
      max_lwc_v = max_lwc * *m_snow / (rho_l * *thick_snow);
      sat_snow = *thick_snow * (*psi_l_snow - max_lwc_v);
      sat_snow = sat_snow / (1.0 - *psi_s_snow - max_lwc_v - r_min(gas_snow_ice2, *psi_g_snow));
      *thick_snow = *thick_snow - sat_snow;
      *thick = *thick + sat_snow;
      *m_snow = *m_snow - sat_snow * (*psi_s_snow * rho_s + (1.0 - *psi_s_snow - gas_snow_ice2) * rho_l);
      *m = *m + sat_snow * (*psi_s_snow * rho_s + (1.0 - *psi_s_snow - gas_snow_ice2) * rho_l);
      *H_abs_snow = *H_abs_snow - sat_snow * *psi_s_snow * rho_s * c_s * *T_snow;
      *H_abs = *H_abs + sat_snow * *psi_s_snow * rho_s * c_s * *T_snow;
      *H_abs_snow = *H_abs_snow + sat_snow * *psi_s_snow * rho_s * latent_heat;
      *H_abs = *H_abs - sat_snow * *psi_s_snow * rho_s * latent_heat;
      *H_abs_snow = *H_abs_snow - sat_snow * (1.0 - *psi_s_snow) * rho_l * c_l * *T_snow;
      *H_abs = *H_abs + sat_snow * (1.0 - *psi_s_snow) * rho_l * c_l * *T_snow;
    } else if (*psi_g_snow <= 0.0) {
      EV(c, SNOW_MERGE);
      sat_snow = *thick_snow;
      *H_abs = *H_abs + *H_abs_snow;
      *m = *m + *m_snow;
      *thick = *thick + *thick_snow;
      *H_abs_snow = 0.0; *m_snow = 0.0; *thick_snow = 0.0;
      *psi_g_snow = 0.0; *psi_s_snow = 0.0; *psi_l_snow = 0.0;
    } else {
      sat_snow = 0.0;
    }
  } else {
    /* mo_snow.f90:398-449 */
    if ((1.0 - phi_snow) > max_lwc && *psi_l_snow > 0.0 && *psi_g_snow > 0.0) {
      real g;
      EV(c, SNOW_WET);
      max_lwc_v = max_lwc * *m_snow / (rho_l * *thick_snow);
      psi_l_snow_slush = (*psi_l_snow - max_lwc_v) * (1.0 - c->k_snow_flush);
      psi_l_snow_flush = (*psi_l_snow - max_lwc_v) * c->k_snow_flush;
      *melt_thick_snow = *thick_snow * psi_l_snow_flush;
      sat_snow = *thick_snow * (psi_l_snow_slush);
      sat_snow = sat_snow / (1.0 - *psi_s_snow - max_lwc_v - r_min(gas_snow_ice2, *psi_g_snow));
      g = r_min(gas_snow_ice2, *psi_g_snow);
      *thick_snow = *thick_snow - sat_snow - *melt_thick_snow;
      *thick = *thick + sat_snow;
      *m_snow = *m_snow - sat_snow * (*psi_s_snow * rho_s + (1.0 - *psi_s_snow - g) * rho_l) - *melt_thick_snow * rho_l;
      *m = *m + sat_snow * (*psi_s_snow * rho_s + (1.0 - *psi_s_snow - g) * rho_l);
      *H_abs_snow = *H_abs_snow - sat_snow * *psi_s_snow * rho_s * c_s * *T_snow;
      *H_abs = *H_abs + sat_snow * *psi_s_snow * rho_s * c_s * *T_snow;
      *H_abs_snow = *H_abs_snow + sat_snow * *psi_s_snow * rho_s * latent_heat;
      *H_abs = *H_abs - sat_snow * *psi_s_snow * rho_s * latent_heat;
      *H_abs_snow = *H_abs_snow - sat_snow * (1.0 - *psi_s_snow - g) * rho_l * c_l * *T_snow -
                    *melt_thick_snow * rho_l * c_l * *T_snow;
      *H_abs = *H_abs + sat_snow * (1.0 - *psi_s_snow - g) * rho_l * c_l * *T_snow;
    } else if (*psi_g_snow <= 0.0) {
      EV(c, SNOW_MERGE);
      sat_snow = *thick_snow;
      *H_abs = *H_abs + *H_abs_snow;
      *m = *m + *m_snow;
      *thick = *thick + *thick_snow;
      *H_abs_snow = 0.0; *m_snow = 0.0; *thick_snow = 0.0;
      *psi_g_snow = 0.0; *psi_s_snow = 0.0; *psi_l_snow = 0.0;
    } else {
      sat_snow = 0.0;
    }
  }
  (void)sat_snow;
  if (*psi_g_snow < 0.0) SAM_STOP(c, 9876);
}

/* the snow block of the driver, mo_grotz.f90:273-292 and :601-621 */
static void snow_thermo_block(sam_col* c) {
  if (c->thick_snow > 0.0) {
    if (c->snow_flush_flag == 0) {
      snow_thermo_any(c, 0, &c->m[1], &c->thick[1], &c->H_abs[1], &c->melt_thick_snow);
      c->melt_thick_snow = 0.0;
    } else if (c->snow_flush_flag == 1) {
      c->melt_thick_snow = 0.0;
      snow_thermo_any(c, 1, &c->m[1], &c->thick[1], &c->H_abs[1], &c->melt_thick_snow);
    }
  } else {
    c->thick_snow = 0.0; c->m_snow = 0.0; c->psi_s_snow = 0.0; c->psi_l_snow = 0.0; c->psi_g_snow = 0.0;
    c->H_abs_snow = 0.0; c->S_abs_snow = 0.0; c->melt_thick_snow = 0.0;
  }
}

/* sub_fl_Q_0_snow_thin, mo_snow.f90:466-487 */
static real sub_fl_Q_0_snow_thin(real m_snow, real thick_snow, real T_snow, real psi_s, real psi_l, real psi_g,
                                 real thick, real T_bound) {
  real k_snow = sam_func_k_snow(m_snow, thick_snow);
  real k = psi_s * k_s + psi_l * k_l + psi_g * 0.0;
  real R;
  k = thick_snow / (thick_snow + thick) * k_snow + thick / (thick_snow + thick) * k;
  R = (thick_snow + thick) / (2.0 * k);
  return (T_snow - T_bound) / R;
}

/* sub_fl_Q_snow, mo_snow.f90:498-518 */
static real sub_fl_Q_snow(real m_snow, real thick_snow, real T_snow, real psi_s_2, real psi_l_2, real thick_2, real T_2) {
  real k_snow = sam_func_k_snow(m_snow, thick_snow);
  real k_2 = psi_s_2 * k_s + psi_l_2 * k_l;
  real R = thick_snow / (2.0 * k_snow) + thick_2 / (2.0 * k_2);
  return (T_2 - T_snow) / R;
}

/* sub_fl_Q_0_snow, mo_snow.f90:528-545 */
static real sub_fl_Q_0_snow(real m_snow, real thick_snow, real T_snow, real T_bound) {
  real k = sam_func_k_snow(m_snow, thick_snow);
  real R = thick_snow / (2.0 * k);
  return (T_snow - T_bound) / R;
}

/* ==========================================================================================
 * mo_flood.f90
 * ======================================================================================== */

/* flood, mo_flood.f90:55-153 */
static void flood(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *psi_s = c->psi_s, *psi_l = c->psi_l, *S_abs = c->S_abs, *H_abs = c->H_abs, *m = c->m, *T = c->T,
       *thick = c->thick;
  const real dt = c->dt, freeboard = c->freeboard, psi_g_snow = c->psi_g_snow;
  real* perm = c->scr[3];
  real* S_bu = c->scr[4];
  real flood_brine, shift_ice, shift_snow, shift, harmonic_perm;
  int k;
  c->stat_flood_calls++;
  EV(c, FLOOD);
  for (k = 1; k <= N_active; k++) perm[k] = 1e-17 * M_POW(1000.0 * psi_l[k], 3.10); /* :73 */
  harmonic_perm = 0.0;
  for (k = 1; k <= N_active - 1; k++) harmonic_perm = harmonic_perm + thick[k] / perm[k]; /* :77-79 */
  harmonic_perm = harmonic_perm + (thick[N_active] * psi_s[N_active] / psi_s_min) / perm[N_active];
  harmonic_perm = (sum_arr(thick, 1, N_active - 1) + thick[N_active] * psi_s[N_active] / psi_s_min) / harmonic_perm;

  flood_brine = -dt * grav * rho_l * rho_l * harmonic_perm * (freeboard) / (mu * sum_arr(thick, 1, N_active)); /* :85 */

  shift_ice = flood_brine / (rho_l * psi_g_snow / ratio_flood); /* :89 */
  shift_snow = shift_ice * (1 + psi_g_snow / (1.0 - psi_g_snow) * (1.0 - 1.0 / ratio_flood));

  for (k = 1; k <= N_active; k++) S_bu[k] = S_abs[k] / m[k];

  S_abs[1] = S_abs[1] + flood_brine * S_bu[N_active]; /* :102-104 */
  H_abs[1] = H_abs[1] + flood_brine * H_abs[N_active] / m[N_active];
  m[1] = m[1] + flood_brine;

  thick[1] = thick[1] + shift_ice; /* :107-112 */
  H_abs[1] = H_abs[1] + shift_snow / c->thick_snow * c->H_abs_snow;
  c->H_abs_snow = c->H_abs_snow - shift_snow / c->thick_snow * c->H_abs_snow;
  m[1] = m[1] + shift_snow / c->thick_snow * c->m_snow;
  c->m_snow = c->m_snow - shift_snow / c->thick_snow * c->m_snow;
  c->thick_snow = c->thick_snow - shift_snow;

  if (freeboard + shift_ice < neg_free) { /* :117-138 */
    EV(c, FLOOD_NEG_FREE);
    shift = neg_free - (freeboard + shift_ice);
    flood_brine = shift * (psi_g_snow)*rho_l;
    S_abs[N_active] = S_abs[N_active] + (c->S_bu_bottom - S_bu[N_active]) * flood_brine;
    H_abs[N_active] = H_abs[N_active] + (c->T_bottom - T[N_active]) * c_l * flood_brine;
    S_abs[1] = S_abs[1] + S_bu[N_active] * flood_brine;
    H_abs[1] = H_abs[1] + T[N_active] * c_l * flood_brine;
    m[1] = m[1] + flood_brine;
    thick[1] = thick[1] + shift;
    H_abs[1] = H_abs[1] + shift / c->thick_snow * c->H_abs_snow;
    c->H_abs_snow = c->H_abs_snow - shift / c->thick_snow * c->H_abs_snow;
    m[1] = m[1] + shift / c->thick_snow * c->m_snow;
    c->m_snow = c->m_snow - shift / c->thick_snow * c->m_snow;
    c->thick_snow = c->thick_snow - shift;
  }
  if (c->bgc_flag == 2) { /* :140-144 */
    FB(c, N_active, 1) = FB(c, N_active, 1) + flood_brine;
    FB(c, N_active + 1, N_active) = FB(c, N_active + 1, N_active) + flood_brine;
  }
  (void)Nlayer;
}

/* flood_simple, mo_flood.f90:167-210 */
static void flood_simple(sam_col* c) {
  real *S_abs = c->S_abs, *H_abs = c->H_abs, *m = c->m, *thick = c->thick;
  real shift = c->freeboard - neg_free;
  real flood_brine = -shift * c->psi_g_snow * rho_l;
  c->stat_flood_calls++;
  EV(c, FLOOD_SIMPLE);
  thick[1] = thick[1] - shift;
  S_abs[1] = S_abs[1] + c->S_bu_bottom * flood_brine;
  H_abs[1] = H_abs[1] - shift / c->thick_snow * c->H_abs_snow;
  H_abs[1] = H_abs[1] + c->T_bottom * c_l * flood_brine;
  m[1] = m[1] - shift / c->thick_snow * c->m_snow;
  m[1] = m[1] + flood_brine;
  c->H_abs_snow = c->H_abs_snow + shift / c->thick_snow * c->H_abs_snow;
  c->m_snow = c->m_snow + shift / c->thick_snow * c->m_snow;
  c->thick_snow = c->thick_snow + shift;
}

/* ==========================================================================================
 * mo_flush.f90
 * ======================================================================================== */

/* flush3, mo_flush.f90:70-237 */
static void flush3(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *psi_l = c->psi_l, *psi_g = c->psi_g, *thick = c->thick, *S_abs = c->S_abs, *H_abs = c->H_abs, *m = c->m,
       *T = c->T, *perm = c->perm, *flush_v = c->flush_v, *flush_h = c->flush_h;
  const real dt = c->dt, freeboard = c->freeboard;
  real* R_h = c->scr[3];
  real* R_v = c->scr[4];
  real* R = c->scr[5];
  real* S_bu = c->scr[6];
  real* fl_m = c->scr[7];
  real konst, flush_total, loss_S_abs, loss_H_abs, sfh;
  int k;
  c->stat_flush_calls++;
  EV(c, FLUSH3);

  for (k = 1; k <= N_active; k++) { /* :101-102 dummy arrays are DIMENSION(N_active) */
    flush_v[k] = 0.0;
    flush_h[k] = 0.0;
  }
  for (k = 0; k <= Nlayer + 2; k++) { S_bu[k] = 0.0; fl_m[k] = 0.0; R[k] = 0.0; R_v[k] = 0.0; R_h[k] = 0.0; } /* :99-100 */
  for (k = 1; k <= N_active; k++) S_bu[k] = S_abs[k] / m[k]; /* :103 */

  konst = sum_arr(thick, 1, N_active) * para_flush_horiz; /* :106 */
  c->melt_thick = r_min(c->melt_thick, psi_l[1] * thick[1]); /* :110 */
  c->melt_thick = r_min(c->melt_thick, c->thick_0 / 3.0);    /* :112 */

  if (c->snow_flush_flag == 1) { /* :114-125 */
    for (k = 1; k <= Nlayer; k++) perm[k] = 0.0;
    for (k = 1; k <= N_active; k++) perm[k] = 1e-17 * M_POW(1000.0 * r_abs(psi_l[k] + 2. * psi_g[k]), 3.10);
    for (k = 1; k <= N_active; k++) {
      if (perm[k] == 0.0) perm[k] = 1.0;
    }
  } else if (c->snow_flush_flag == 0) { /* :126-130 */
    for (k = 1; k <= Nlayer; k++) perm[k] = 1.0;
    for (k = 1; k <= N_active; k++) perm[k] = 1e-17 * M_POW(1000.0 * r_abs(psi_l[k]), 3.10);
  }

  for (k = 1; k <= N_active; k++) { /* :133-137 */
    R_v[k] = mu * thick[k] / r_max(perm[k], 0.00000000000000000000001);
    R_h[k] = mu * konst / (thick[k] * r_max(perm[k], 0.00000000000000000000001));
  }
  R[N_active] = 0.0;
  R[N_active - 1] = R_v[N_active - 1];
  if (N_active > 2) { /* :141-146 */
    for (k = N_active - 2; k >= 1; k--) {
      R[k] = R[k + 1] + R_v[k];
      R[k] = ((R[k]) * R_h[k]) / (R[k] + R_h[k]);
    }
  }

  flush_total = (freeboard + c->melt_thick) / R[1] * grav * dt * sam_func_density(T[1], sam_func_S_br(c, T[1])) * rho_l; /* :152 */
  flush_total = r_min(flush_total, c->melt_thick * rho_l); /* :155 */
  c->melt_err = c->melt_err + c->melt_thick - r_min(flush_total / rho_l, c->melt_thick); /* :156 */

  flush_h[1] = flush_total * (R[2] + R_v[1]) / (R[2] + R_v[1] + R_h[1]); /* :159-160 */
  flush_v[1] = flush_total * R_h[1] / (R[2] + R_v[1] + R_h[1]);
  for (k = 2; k <= N_active - 1; k++) { /* :161-164 */
    flush_h[k] = flush_v[k - 1] * (R[k + 1] + R_v[k]) / (R[k + 1] + R_v[k] + R_h[k]);
    flush_v[k] = flush_v[k - 1] * R_h[k] / (R[k + 1] + R_v[k] + R_h[k]);
  }
  flush_v[N_active] = flush_v[N_active - 1];
  flush_h[N_active] = 0.0;

  if (c->bgc_flag == 2) { /* :168-175; SUM(flush_h(:)) runs over the N_active-long dummy */
    for (k = 1; k <= N_active - 1; k++) FB(c, k, N_active) = FB(c, k, N_active) + flush_h[k];
    FB(c, N_active, N_active + 1) = FB(c, N_active, N_active + 1) + sum_arr(flush_h, 1, N_active);
    for (k = 1; k <= N_active; k++) FB(c, k, k + 1) = FB(c, k, k + 1) + flush_v[k];
  }

  fl_m[1] = 0.0; /* :179-180 */
  for (k = 1; k <= N_active; k++) fl_m[k + 1] = -flush_v[k];

  mass_transfer(c, T, H_abs, S_abs, S_bu, fl_m); /* :182 */

  if (c->flush_heat_flag == 2) { /* :185-187 */
    H_abs[N_active] = H_abs[N_active] - fl_m[N_active + 1] * T[N_active] * c_l;
  }

  m[1] = m[1] - flush_total; /* :190-191 */
  thick[1] = thick[1] - flush_total / rho_l;

  for (k = 1; k <= N_active - 1; k++) { /* :196-206 */
    loss_S_abs = flush_h[k] * sam_func_S_br2(c, T[k], S_abs[k] / m[k]);
    loss_H_abs = flush_h[k] * T[k] * c_l;
    S_abs[k] = S_abs[k] - loss_S_abs;
    H_abs[k] = H_abs[k] - loss_H_abs;
    H_abs[N_active] = H_abs[N_active] + loss_H_abs;
    S_abs[N_active] = S_abs[N_active] + loss_S_abs;
  }
  sfh = sum_arr(flush_h, 1, N_active); /* SUM(flush_h), dummy is DIMENSION(N_active) */
  loss_S_abs = sfh * S_bu[N_active];   /* :207 */
  loss_H_abs = sfh * T[N_active] * c_l;

  if (c->flush_heat_flag == 2) H_abs[N_active] = H_abs[N_active] - loss_H_abs; /* :211-213 */
  S_abs[N_active] = S_abs[N_active] - loss_S_abs;

  {
    real mn = S_abs[1]; /* :218 MINVAL(S_abs) all layers */
    for (k = 1; k <= Nlayer; k++) mn = r_min(mn, S_abs[k]);
    if (mn < -0.00000000000000000000000001) {
      EV(c, FLUSH3_CLAMP);
      for (k = 1; k <= N_active; k++) S_abs[k] = r_max(S_abs[k], 0.0);
    }
  }
  if (r_abs(m[1]) < 0.000001) SAM_STOP(c, 9876); /* :230-233 */
}

/* flush4, mo_flush.f90:253-296 (flush_flag 6; not used by the configs) */
static void flush4(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *psi_l = c->psi_l, *thick = c->thick, *T = c->T, *S_abs = c->S_abs, *H_abs = c->H_abs, *m = c->m;
  real S_bu1 = S_abs[1] / m[1];
  int k;
  EV(c, FLUSH4);
  H_abs[1] = H_abs[1] - c->melt_thick * rho_l * c_l * T[1];
  S_abs[1] = S_abs[1] - c->melt_thick * rho_l * sam_func_S_br2(c, T[1], S_bu1);
  thick[1] = thick[1] - c->melt_thick;
  m[1] = m[1] - c->melt_thick * rho_l;
  c->melt_thick = 0.0;
  k = 2;
  while (k <= Nlayer && psi_l[k] > psi_l[k - 1]) {
    S_abs[k] = para_flush_gamma * S_abs[k];
    k = k + 1;
  }
  S_abs[1] = r_max(S_abs[1], 0.00);
  {
    real mn = S_abs[1];
    for (k = 1; k <= Nlayer; k++) mn = r_min(mn, S_abs[k]);
    if (mn < 0.0) SAM_STOP(c, 9876);
  }
  (void)N_active;
}

/* ==========================================================================================
 * mo_layer_dynamics.f90
 * ======================================================================================== */

/* top_melt, mo_layer_dynamics.f90:191-326 */
/* tracers in the layer-dynamics routines: bgc_temp / bgc_bulk of the reference (the routines work on bgc_abs in place:
 * bgc_temp is copied back unconditionally at their end) */
#define BGC_DECL(c) real* bgc[3] = {NULL, (c)->bgc_abs[1], (c)->bgc_abs[2]}; real* bulk[3] = {NULL, (c)->scr[6], (c)->scr[7]}; \
  const int n_bgc_ = ((c)->bgc_flag == 2) ? (c)->N_bgc : 0
#define BGC_EACH(q) for ((q) = 1; (q) <= n_bgc_; (q)++)

static void top_melt(sam_col* c, real* rho, real* H, real* S_bu) {
  const int Nlayer = c->Nlayer, N_middle = c->N_middle, N_top = c->N_top;
  const real thick_0 = c->thick_0;
  real *m = c->m, *S_abs = c->S_abs, *H_abs = c->H_abs, *thick = c->thick;
  real loss_m, loss_S_abs, loss_H_abs, shift;
  real loss_bgc[3] = {0.0, 0.0, 0.0};
  int k, kmax, q;
  BGC_DECL(c);
  for (k = 1; k <= c->N_active; k++) { /* :218-223 */
    rho[k] = m[k] / thick[k];
    S_bu[k] = S_abs[k] / m[k];
    H[k] = H_abs[k] / m[k];
    BGC_EACH(q) bulk[q][k] = bgc[q][k] / m[k];
  }
  m[1] = m[1] + m[2]; /* :231-235 */
  S_abs[1] = S_abs[1] + S_abs[2];
  H_abs[1] = H_abs[1] + H_abs[2];
  BGC_EACH(q) bgc[q][1] = bgc[q][1] + bgc[q][2];
  thick[1] = thick[1] + thick[2];

  kmax = (N_top - 1 < c->N_active - 1) ? N_top - 1 : c->N_active - 1;
  for (k = 2; k <= kmax; k++) { /* :238-243 */
    m[k] = rho[k + 1] * thick_0;
    S_abs[k] = S_bu[k + 1] * rho[k + 1] * thick_0;
    H_abs[k] = H[k + 1] * rho[k + 1] * thick_0;
    BGC_EACH(q) bgc[q][k] = bulk[q][k + 1] * rho[k + 1] * thick_0;
  }

  if (c->N_active <= N_top) { /* :247-254 */
    EV(c, TOP_MELT_A);
    m[c->N_active] = 0.0; S_abs[c->N_active] = 0.0; H_abs[c->N_active] = 0.0; thick[c->N_active] = 0.0;
    BGC_EACH(q) bgc[q][c->N_active] = 0.0;
    c->N_active = c->N_active - 1;
  } else if (c->N_active > N_top && c->N_active <= Nlayer && thick[N_top + 1] / thick_0 < 1.00001) { /* :256-273 */
    EV(c, TOP_MELT_B);
    for (k = N_top; k <= c->N_active - 1; k++) {
      m[k] = rho[k + 1] * thick_0;
      S_abs[k] = S_bu[k + 1] * rho[k + 1] * thick_0;
      H_abs[k] = H[k + 1] * rho[k + 1] * thick_0;
      BGC_EACH(q) bgc[q][k] = bulk[q][k + 1] * rho[k + 1] * thick_0;
    }
    m[c->N_active] = 0.0; S_abs[c->N_active] = 0.0; H_abs[c->N_active] = 0.0; thick[c->N_active] = 0.0;
    BGC_EACH(q) bgc[q][c->N_active] = 0.0;
    c->N_active = c->N_active - 1;
  }

  if (c->N_active == Nlayer && thick[N_top + 1] - thick_0 >= 0.000001) { /* :275-314 */
    EV(c, TOP_MELT_C);
    loss_m = thick_0 * rho[N_top + 1];
    loss_S_abs = loss_m * S_bu[N_top + 1];
    loss_H_abs = loss_m * H[N_top + 1];
    BGC_EACH(q) loss_bgc[q] = loss_m * bulk[q][N_top + 1];
    m[N_top] = loss_m;
    S_abs[N_top] = loss_S_abs;
    H_abs[N_top] = loss_H_abs;
    BGC_EACH(q) bgc[q][N_top] = loss_bgc[q];
    for (k = N_top + 1; k <= N_middle + N_top; k++) {
      m[k] = m[k] - loss_m;
      H_abs[k] = H_abs[k] - loss_H_abs;
      S_abs[k] = S_abs[k] - loss_S_abs;
      BGC_EACH(q) bgc[q][k] = bgc[q][k] - loss_bgc[q];
      shift = thick_0 * (double)(float)(N_middle - k + N_top) / (double)(float)(N_middle); /* :293 */
      loss_m = shift * rho[k + 1];
      loss_S_abs = loss_m * S_bu[k + 1];
      loss_H_abs = loss_m * H[k + 1];
      BGC_EACH(q) loss_bgc[q] = loss_m * bulk[q][k + 1];
      m[k] = m[k] + loss_m;
      H_abs[k] = H_abs[k] + loss_H_abs;
      S_abs[k] = S_abs[k] + loss_S_abs;
      BGC_EACH(q) bgc[q][k] = bgc[q][k] + loss_bgc[q];
    }
    for (k = N_top + 1; k <= N_top + N_middle; k++) thick[k] = thick[k] - thick_0 / (double)(float)(N_middle); /* :311-313 */
  }

  if (thick_0 * (c->N_active + 0.501) <= sum_arr(thick, 1, Nlayer) && c->N_active < Nlayer) SAM_STOP(c, 7889); /* :318-321 */
}

/* bottom_melt, mo_layer_dynamics.f90:341-420 */
static void bottom_melt(sam_col* c, real* rho, real* H, real* S_bu) {
  const int Nlayer = c->Nlayer, N_middle = c->N_middle, N_top = c->N_top;
  real *m = c->m, *S_abs = c->S_abs, *H_abs = c->H_abs, *thick = c->thick;
  real loss_m = 0.0, loss_S_abs = 0.0, loss_H_abs = 0.0, shift;
  real loss_bgc[3] = {0.0, 0.0, 0.0};
  int k, q;
  BGC_DECL(c);
  for (k = N_top + 1; k <= Nlayer; k++) { /* :364-370 */
    rho[k] = m[k] / thick[k];
    S_bu[k] = S_abs[k] / m[k];
    H[k] = H_abs[k] / m[k];
    BGC_EACH(q) bulk[q][k] = bgc[q][k] / m[k];
  }
  for (k = N_top + 1; k <= N_top + N_middle; k++) { /* :378-400 */
    m[k] = m[k] + loss_m;
    H_abs[k] = H_abs[k] + loss_H_abs;
    S_abs[k] = S_abs[k] + loss_S_abs;
    BGC_EACH(q) bgc[q][k] = bgc[q][k] + loss_bgc[q];
    shift = thick[Nlayer] * (k - N_top) / (double)(float)(N_middle); /* :387 */
    loss_m = shift * rho[k];
    loss_H_abs = loss_m * H[k];
    loss_S_abs = loss_m * S_bu[k];
    BGC_EACH(q) loss_bgc[q] = loss_m * bulk[q][k];
    m[k] = m[k] - loss_m;
    H_abs[k] = H_abs[k] - loss_H_abs;
    S_abs[k] = S_abs[k] - loss_S_abs;
    BGC_EACH(q) bgc[q][k] = bgc[q][k] - loss_bgc[q];
  }
  for (k = N_top + 1; k <= N_top + N_middle; k++) thick[k] = thick[k] - thick[Nlayer] / (double)(float)(N_middle); /* :405-407 */
  for (k = N_top + N_middle + 1; k <= Nlayer; k++) { /* :410-415 */
    H_abs[k] = rho[k - 1] * thick[k] * H[k - 1];
    S_abs[k] = rho[k - 1] * thick[k] * S_bu[k - 1];
    m[k] = rho[k - 1] * thick[k];
    BGC_EACH(q) bgc[q][k] = rho[k - 1] * thick[k] * bulk[q][k - 1];
  }
}

/* bottom_growth, mo_layer_dynamics.f90:438-520 */
static void bottom_growth(sam_col* c, real* rho, real* H, real* S_bu) {
  const int Nlayer = c->Nlayer, N_middle = c->N_middle, N_top = c->N_top, N_bottom = c->N_bottom;
  real *m = c->m, *S_abs = c->S_abs, *H_abs = c->H_abs, *thick = c->thick;
  real gain_m = 0.0, gain_S_abs = 0.0, gain_H_abs = 0.0, shift;
  real gain_bgc[3] = {0.0, 0.0, 0.0};
  int k, q;
  BGC_DECL(c);
  for (k = N_top + 1; k <= N_top + N_middle + 1; k++) { /* :463-468 */
    rho[k] = m[k] / thick[k];
    S_bu[k] = S_abs[k] / m[k];
    H[k] = H_abs[k] / m[k];
    BGC_EACH(q) bulk[q][k] = bgc[q][k] / m[k];
  }
  for (k = N_top + 1; k <= N_top + N_middle; k++) { /* :476-495 */
    m[k] = m[k] - gain_m;
    H_abs[k] = H_abs[k] - gain_H_abs;
    S_abs[k] = S_abs[k] - gain_S_abs;
    BGC_EACH(q) bgc[q][k] = bgc[q][k] - gain_bgc[q];
    shift = thick[Nlayer] * (k - N_top) / (double)(float)(N_middle); /* :483 */
    gain_m = shift * rho[k + 1];
    gain_H_abs = gain_m * H[k + 1];
    gain_S_abs = gain_m * S_bu[k + 1];
    BGC_EACH(q) gain_bgc[q] = gain_m * bulk[q][k + 1];
    m[k] = m[k] + gain_m;
    H_abs[k] = H_abs[k] + gain_H_abs;
    S_abs[k] = S_abs[k] + gain_S_abs;
    BGC_EACH(q) bgc[q][k] = bgc[q][k] + gain_bgc[q];
  }
  for (k = N_top + 1; k <= N_top + N_middle; k++) thick[k] = thick[k] + thick[Nlayer] / (double)(float)(N_middle); /* :498-500 */
  for (k = Nlayer - N_bottom + 1; k <= Nlayer - 1; k++) { /* :503-508 */
    H_abs[k] = H_abs[k + 1];
    S_abs[k] = S_abs[k + 1];
    m[k] = m[k + 1];
    BGC_EACH(q) bgc[q][k] = bgc[q][k + 1];
  }
  m[Nlayer] = thick[Nlayer] * rho_l; /* :511-513 */
  H_abs[Nlayer] = m[Nlayer] * c->T_bottom * c_l;
  S_abs[Nlayer] = m[Nlayer] * c->S_bu_bottom;
  BGC_EACH(q) bgc[q][Nlayer] = m[Nlayer] * c->bgc_bottom[q]; /* :516 */
}

/* bottom_growth_simple, mo_layer_dynamics.f90:537-561 */
static void bottom_growth_simple(sam_col* c) {
  real *m = c->m, *S_abs = c->S_abs, *H_abs = c->H_abs, *thick = c->thick;
  c->N_active = c->N_active + 1;
  thick[c->N_active] = c->thick_0;
  m[c->N_active] = thick[c->N_active] * rho_l;
  H_abs[c->N_active] = m[c->N_active] * c->T_bottom * c_l;
  S_abs[c->N_active] = m[c->N_active] * c->S_bu_bottom;
  if (c->bgc_flag == 2) {
    int q;
    for (q = 1; q <= c->N_bgc; q++) c->bgc_abs[q][c->N_active] = c->bgc_bottom[q] * m[c->N_active]; /* :557 */
  }
}

/* bottom_melt_simple, mo_layer_dynamics.f90:573-590 */
static void bottom_melt_simple(sam_col* c) {
  c->thick[c->N_active] = 0.0;
  c->m[c->N_active] = 0.0;
  c->S_abs[c->N_active] = 0.0;
  c->H_abs[c->N_active] = 0.0;
  if (c->bgc_flag == 2) {
    int q;
    for (q = 1; q <= c->N_bgc; q++) c->bgc_abs[q][c->N_active] = 0.0; /* :586 */
  }
  c->N_active = c->N_active - 1;
}

/* top_grow, mo_layer_dynamics.f90:607-716 */
static void top_grow(sam_col* c, real* rho, real* H, real* S_bu) {
  const int Nlayer = c->Nlayer, N_middle = c->N_middle, N_top = c->N_top;
  const real thick_0 = c->thick_0;
  real *m = c->m, *S_abs = c->S_abs, *H_abs = c->H_abs, *thick = c->thick;
  real loss_m, loss_S_abs, loss_H_abs, shift;
  real loss_bgc[3] = {0.0, 0.0, 0.0};
  int k, kmax, q;
  BGC_DECL(c);
  for (k = 1; k <= c->N_active; k++) { /* :631-636 */
    rho[k] = m[k] / thick[k];
    S_bu[k] = S_abs[k] / m[k];
    H[k] = H_abs[k] / m[k];
    BGC_EACH(q) bulk[q][k] = bgc[q][k] / m[k];
  }
  loss_m = thick_0 * rho[1]; /* :639-642 */
  loss_S_abs = loss_m * S_bu[1];
  loss_H_abs = loss_m * H[1];
  BGC_EACH(q) loss_bgc[q] = loss_m * bulk[q][1];
  m[1] = m[1] - loss_m; /* :644-648 */
  S_abs[1] = S_abs[1] - loss_S_abs;
  H_abs[1] = H_abs[1] - loss_H_abs;
  BGC_EACH(q) bgc[q][1] = bgc[q][1] - loss_bgc[q];
  thick[1] = thick[1] - thick_0;

  kmax = (N_top < c->N_active) ? N_top : c->N_active;
  for (k = 2; k <= kmax; k++) { /* :651-656 */
    m[k] = rho[k - 1] * thick_0;
    S_abs[k] = S_bu[k - 1] * rho[k - 1] * thick_0;
    H_abs[k] = H[k - 1] * rho[k - 1] * thick_0;
    BGC_EACH(q) bgc[q][k] = bulk[q][k - 1] * rho[k - 1] * thick_0;
  }

  if (c->N_active <= N_top) { /* :659-665 */
    EV(c, TOP_GROW_A);
    c->N_active = c->N_active + 1;
    m[c->N_active] = rho[c->N_active - 1] * thick_0;
    S_abs[c->N_active] = S_bu[c->N_active - 1] * thick_0 * rho[c->N_active - 1];
    H_abs[c->N_active] = H[c->N_active - 1] * thick_0 * rho[c->N_active - 1];
    BGC_EACH(q) bgc[q][c->N_active] = bulk[q][c->N_active - 1] * thick_0 * rho[c->N_active - 1];
    thick[c->N_active] = thick_0;
  } else if (c->N_active > N_top && c->N_active < Nlayer) { /* :668-680 */
    EV(c, TOP_GROW_B);
    for (k = N_top + 1; k <= c->N_active; k++) {
      m[k] = rho[k - 1] * thick_0;
      S_abs[k] = S_bu[k - 1] * rho[k - 1] * thick_0;
      H_abs[k] = H[k - 1] * rho[k - 1] * thick_0;
      BGC_EACH(q) bgc[q][k] = bulk[q][k - 1] * rho[k - 1] * thick_0;
    }
    c->N_active = c->N_active + 1;
    m[c->N_active] = rho[c->N_active - 1] * thick_0;
    S_abs[c->N_active] = S_bu[c->N_active - 1] * thick_0 * rho[c->N_active - 1];
    H_abs[c->N_active] = H[c->N_active - 1] * thick_0 * rho[c->N_active - 1];
    BGC_EACH(q) bgc[q][c->N_active] = bulk[q][c->N_active - 1] * thick_0 * rho[c->N_active - 1];
    thick[c->N_active] = thick_0;
  } else if (c->N_active == Nlayer) { /* :682-711 */
    EV(c, TOP_GROW_C);
    loss_m = thick_0 * rho[N_top];
    loss_S_abs = loss_m * S_bu[N_top];
    loss_H_abs = loss_m * H[N_top];
    BGC_EACH(q) loss_bgc[q] = loss_m * bulk[q][N_top];
    for (k = N_top + 1; k <= N_middle + N_top; k++) {
      m[k] = m[k] + loss_m;
      H_abs[k] = H_abs[k] + loss_H_abs;
      S_abs[k] = S_abs[k] + loss_S_abs;
      BGC_EACH(q) bgc[q][k] = bgc[q][k] + loss_bgc[q];
      shift = thick_0 * (double)(float)(N_middle - k + N_top) / (double)(float)(N_middle); /* :692 */
      loss_m = shift * rho[k];
      loss_S_abs = loss_m * S_bu[k];
      loss_H_abs = loss_m * H[k];
      BGC_EACH(q) loss_bgc[q] = loss_m * bulk[q][k];
      m[k] = m[k] - loss_m;
      H_abs[k] = H_abs[k] - loss_H_abs;
      S_abs[k] = S_abs[k] - loss_S_abs;
      BGC_EACH(q) bgc[q][k] = bgc[q][k] - loss_bgc[q];
    }
    for (k = N_top + 1; k <= N_top + N_middle; k++) thick[k] = thick[k] + thick_0 / (double)(float)(N_middle); /* :707-709 */
  }
}

/* layer_dynamics, mo_layer_dynamics.f90:64-175 */
static void layer_dynamics(sam_col* c) {
  const int Nlayer = c->Nlayer, N_top = c->N_top;
  const real thick_0 = c->thick_0;
  real *phi = c->phi, *thick = c->thick;
  const int N_active = c->N_active;
  const int bottom_flag = c->bottom_flag;
  int nm1 = (N_active - 1 > 1) ? N_active - 1 : 1; /* MAX(N_active-1,1) */
  real* rho = c->scr[3];
  real* H = c->scr[4];
  real* S_bu = c->scr[5];
  c->stat_layer_events++;
  if (phi[Nlayer - 1] <= psi_s_min / 2.0 && phi[N_active] < 0.00001 && N_active == Nlayer &&
      thick[N_top + 1] / thick_0 > 1.000001 && bottom_flag == 1) { /* :85-86 */
    EV(c, BOTTOM_MELT);
    bottom_melt(c, rho, H, S_bu);
  } else if (N_active > 1 && N_active < Nlayer && phi[N_active] < 0.00001 && phi[nm1] <= psi_s_min / 2.0 &&
             bottom_flag == 1) { /* :95-96 */
    EV(c, BOTTOM_MELT_SIMPLE_A);
    bottom_melt_simple(c);
  } else if (N_active > 1 && phi[N_active] < 0.00001 && phi[nm1] <= psi_s_min / 2.0 &&
             (thick[N_top + 1] / thick_0) < 1.01 && bottom_flag == 1) { /* :106-107 */
    EV(c, BOTTOM_MELT_SIMPLE_B);
    bottom_melt_simple(c);
  } else if (phi[N_active] > psi_s_min && N_active < Nlayer && bottom_flag == 1) { /* :122 */
    EV(c, BOTTOM_GROWTH_SIMPLE);
    bottom_growth_simple(c);
  } else if (phi[Nlayer] > psi_s_min && bottom_flag == 1) { /* :132 */
    EV(c, BOTTOM_GROWTH);
    bottom_growth(c, rho, H, S_bu);
  } else if (thick[1] > 1.5 * thick_0) { /* :145 */
    c->melt_thick_output[3] = c->melt_thick_output[3] - thick[1];
    top_grow(c, rho, H, S_bu);
    c->melt_thick_output[3] = c->melt_thick_output[3] + thick[1];
  } else if (thick[1] < 0.5 * thick_0) { /* :160 */
    c->melt_thick_output[3] = c->melt_thick_output[3] - thick[1];
    top_melt(c, rho, H, S_bu);
    c->melt_thick_output[3] = c->melt_thick_output[3] + thick[1];
  } else {
    c->stat_layer_events--; /* guard true but no operation matched */
  }
}

/* ==========================================================================================
 * mo_testcase_specifics.f90
 * ======================================================================================== */

/* sub_test1, mo_testcase_specifics.f90:42-89: T_top = -10 at t = 12,36,...,228 h; -5 at t = 24,...,240 h */
static void sub_test1(sam_col* c) {
  int j;
  for (j = 1; j <= 20; j++) {
    if (r_abs(c->time - (12.0 * j * 3600.0)) < F32(0.01)) {
      c->T_top = (j % 2 == 1) ? c->ttop_cold : c->ttop_warm;
      break;
    }
  }
}

/* sub_test4, mo_testcase_specifics.f90:197-202 */
static void sub_test4(sam_col* c) {
  c->fl_q_bottom = -c->oflux_amp * M_SIN(c->time * (2.0 * pi_sp) / (86400.0 * 365.0)) + c->oflux_amp;
}

/* sub_test2, mo_testcase_specifics.f90:92-101 */
static void sub_test2(sam_col* c) {
  if (c->time > 86400.0 * 25.0) c->T2m = 15.0;
  else if (c->time > 86400.0 * 15.0) c->T2m = 1.0;
}

/* sub_test9, mo_testcase_specifics.f90:105-116 */
static void sub_test9(sam_col* c) {
  if (c->time < (19.75 * 3600.0)) c->T2m = 0.0;
  else if (c->time < (86400.0 * 3.0 + 2.25 * 3600.0)) c->T2m = -15.0;
  else c->T2m = 1.0;
}

/* sub_test34, mo_testcase_specifics.f90:146-161 */
static void sub_test34(sam_col* c) {
  if (c->time < 2.0 * 3600.0) c->T2m = 0.0;
  else if (c->time < (86400.0 * 5.0)) c->T2m = -15.0;
  else if (c->time < (86400.0 * 7.0)) c->T2m = -5.0;
  else c->T2m = 1.0;
}

/* testcase 99, mo_grotz.f90:547-563: air temperature schedule; from day 3 on the snow cover is reset every step */
static void hook_test99(sam_col* c) {
  if (c->time < (double)(86400.f * 3.f)) {
    c->T2m = -40.0;
  } else {
    c->T2m = (c->time > (double)(86400.f * 5.f)) ? 5.0 : -5.0;
    c->thick_snow = 0.2;
    c->T_snow = -5.0;
    c->m_snow = 30.0;
    c->H_abs_snow = -c->m_snow * latent_heat;
  }
}

/* sub_test3, mo_testcase_specifics.f90:170-185 (the day counter it computes is unused) */
static void sub_test3(sam_col* c) {
  c->liquid_precip = 0.0;
  c->solid_precip = 0.15 / 86400.0 / 356.0;
}

/* sub_test6, mo_testcase_specifics.f90:218-243 */
static void sub_test6(sam_col* c) {
  const real t = c->time;
  if (t > 1714.0 * 60.0) c->T2m = -19.0;
  else if (t > 1676.0 * 60.0) c->T2m = -5.0;
  else if (t > 1525.0 * 60.0) c->T2m = -18.0;
  else if (t > 1483.0 * 60.0) c->T2m = -5.0;
  else if (t > 1385.0 * 60.0) c->T2m = -18.0;
  else if (t > 1349.0 * 60.0) c->T2m = -5.0;
  else if (t > 1160.0 * 60.0) c->T2m = -18.0;
  else if (t > 1100.0 * 60.0) c->T2m = -5.0;
}

/* ==========================================================================================
 * mo_heat_fluxes.f90
 * ======================================================================================== */

/* sub_heat_fluxes, mo_heat_fluxes.f90:69-312 */
static void sub_heat_fluxes(sam_col* c) {
  const int N_active = c->N_active, Nlayer = c->Nlayer;
  real *psi_s = c->psi_s, *psi_l = c->psi_l, *psi_g = c->psi_g, *thick = c->thick, *T = c->T, *fl_Q = c->fl_Q,
       *fl_rad = c->fl_rad, *H_abs = c->H_abs, *S_abs = c->S_abs, *m = c->m;
  const real dt = c->dt, thick_min = c->thick_min;
  real T_old, emi, pen, temp, temp1, temp2;
  int k;

  if (c->boundflux_flag == 1) { /* :77-86 */
    fl_Q[1] = sub_fl_Q_0(c, psi_s[1], psi_l[1], psi_g[1], thick[1], T[1], c->T_top, -1);
    if (r_abs(fl_Q[1]) > c->max_flux_plate) {
      fl_Q[1] = fl_Q[1] / r_abs(fl_Q[1]) * c->max_flux_plate;
    }
  }

  if (c->boundflux_flag == 2) { /* :90-195 */
    c->albedo = sam_func_albedo(c->thick_snow, c->T_snow, psi_l[1], thick_min, c->albedo_flag); /* :94 */
    if (c->atmoflux_flag == 1) {
      EV(c, NOTZFLUX);
      sub_notzflux(c->time + 86400.0 * 180.0, &c->fl_sw, &c->fl_rest);
    } else if (c->atmoflux_flag == 2) { /* :97-111 */
      const int tc = c->time_counter;
      if (c->time == c->time_input[tc]) {
        c->fl_sw = c->fl_sw_input[tc];
        c->fl_lw = c->fl_lw_input[tc];
      } else {
        temp = (c->time - c->time_input[tc - 1]) / (c->time_input[tc] - c->time_input[tc - 1]);
        c->fl_sw = (1.0 - temp) * c->fl_sw_input[tc - 1] + temp * c->fl_sw_input[tc];
        c->fl_lw = (1.0 - temp) * c->fl_lw_input[tc - 1] + temp * c->fl_lw_input[tc];
      }
      c->fl_sen = 0.0;
      c->fl_lat = 0.0;
      c->fl_rest = c->fl_lw + c->fl_sen + c->fl_lat;
    }

    if (c->thick_snow < thick_min) T_old = T[1]; else T_old = c->T_snow; /* :114-118 */
    if (c->thick_snow < thick_min) { /* :119-129 */
      emi = emissivity_ice;
      pen = penetr;
    } else {
      emi = emissivity_snow;
      pen = 0.0;
    }
    T_old = T_old + zeroK; /* :131 */

    temp1 = (1.0 - c->albedo) * (1.0 - pen) * c->fl_sw + c->fl_rest; /* :135-139 */
    temp1 = temp1 + emi * 3.0 * sigma * P4(T_old);
    temp1 = temp1 / (emi * 4.0 * sigma * P3(T_old));
    temp1 = temp1 - zeroK;

    T_old = temp1 + zeroK; /* :141-146 */
    temp1 = (1.0 - c->albedo) * (1.0 - pen) * c->fl_sw + c->fl_rest;
    temp1 = temp1 + emi * 3.0 * sigma * P4(T_old);
    temp1 = temp1 / (emi * 4.0 * sigma * P3(T_old));
    temp1 = temp1 - zeroK;

    c->T_top = temp1; /* :148 */

    temp2 = pen * (1.0 - c->albedo) * c->fl_sw; /* :151-155 */
    for (k = 1; k <= N_active; k++) {
      fl_rad[k] = temp2 - temp2 * M_EXP(-extinc * thick[k]);
      temp2 = temp2 * M_EXP(-extinc * thick[k]);
    }

    if (c->thick_snow >= thick_min / 100.0) { /* :158-162 */
      c->T_freeze = 0.0;
    } else {
      c->T_freeze = sam_func_T_freeze(S_abs[1] / m[1], c->salt_flag);
    }

    if (c->T_top > c->T_freeze && N_active > 1) { /* :167-180 */
      EV(c, HEAT_MELT);
      temp1 = emi * sigma * P4(c->T_freeze + zeroK) - (1.0 - c->albedo) * (1.0 - pen) * c->fl_sw - c->fl_rest;
      if (c->thick_snow >= thick_min) {
        c->fl_q_snow = temp1;
        fl_Q[1] = sub_fl_Q_snow(c->m_snow, c->thick_snow, c->T_snow, psi_s[1], psi_l[1], thick[1], T[1]);
      } else if (c->thick_snow >= thick_min / 100.0) {
        c->fl_q_snow = temp1;
        fl_Q[1] = 0.0;
      } else {
        fl_Q[1] = temp1;
      }
      c->T_top = c->T_freeze;
    } else { /* :185-193 */
      if (c->thick_snow >= thick_min) {
        fl_Q[1] = sub_fl_Q_snow(c->m_snow, c->thick_snow, c->T_snow, psi_s[1], psi_l[1], thick[1], T[1]);
        c->fl_q_snow = sub_fl_Q_0_snow(c->m_snow, c->thick_snow, c->T_snow, c->T_top);
      } else if (c->thick_snow > thick_min / 100.0 && c->thick_snow < thick_min) {
        fl_Q[1] = 0.0;
        c->fl_q_snow = sub_fl_Q_0_snow_thin(c->m_snow, c->thick_snow, c->T_snow, psi_s[1], psi_l[1], psi_g[1], thick[1], c->T_top);
      } else {
        fl_Q[1] = sub_fl_Q_0(c, psi_s[1], psi_l[1], psi_g[1], thick[1], T[1], c->T_top, -1);
      }
    }
  }

  if (c->boundflux_flag == 3) { /* :202-258 */
    if (c->lab_snow_flag == 0 || c->thick_snow <= thick_min / 100.0) { /* :206-219 */
      c->T_freeze = r_min(sam_func_T_freeze(S_abs[N_active] / m[N_active], c->salt_flag), 0.0);
      c->T_top = T[1];
      fl_Q[1] = c->alpha_flux_instable * (c->T_top - c->T2m);
      if (fl_Q[1] < 0.0) {
        c->T_top = r_max(c->T_freeze, T[1]);
        fl_Q[1] = c->alpha_flux_stable * (c->T_top - c->T2m);
      }
      if (c->thick_snow == 0.0 && c->lab_snow_flag == 1 && c->styropor_flag == 1) {
        EV(c, STYROPOR);
        fl_Q[1] = fl_Q[1] * c->k_styropor; /* sub_fl_Q_styropor, mo_thermo_functions.f90:276-287 */
      }
    } else if (c->lab_snow_flag == 1) { /* :224-256 */
      c->T_freeze = sam_func_T_freeze(c->S_abs_snow / c->m_snow, c->salt_flag);
      c->T_top = c->T_snow;
      temp1 = c->alpha_flux_instable * (c->T_top - c->T2m);
      if (temp1 >= 0.0) {
        if (c->thick_snow >= thick_min) {
          c->fl_q_snow = temp1;
          fl_Q[1] = sub_fl_Q_snow(c->m_snow, c->thick_snow, c->T_snow, psi_s[1], psi_l[1], thick[1], T[1]);
        } else if (c->thick_snow >= thick_min / 100.0) {
          c->fl_q_snow = sub_fl_Q_0_snow_thin(c->m_snow, c->thick_snow, c->T_snow, psi_s[1], psi_l[1], psi_g[1], thick[1],
                                              (c->T2m + c->T_top) / 2.0);
          fl_Q[1] = 0.0;
        }
      } else {
        temp1 = c->alpha_flux_stable * (c->T_top - c->T2m);
        if (c->thick_snow >= thick_min) {
          c->fl_q_snow = temp1;
          fl_Q[1] = sub_fl_Q_snow(c->m_snow, c->thick_snow, c->T_snow, psi_s[1], psi_l[1], thick[1], T[1]);
        } else if (c->thick_snow >= thick_min / 100.0) {
          c->fl_q_snow = temp1;
          fl_Q[1] = 0.0;
        }
      }
    }
  }

  fl_Q[N_active + 1] = c->fl_q_bottom; /* :262 */

  temp1 = sum_arr(H_abs, 1, Nlayer) + c->H_abs_snow; /* :269 */

  for (k = 2; k <= N_active; k++) { /* :272-274 */
    fl_Q[k] = sub_fl_Q(psi_s[k - 1], psi_l[k - 1], psi_g[k - 1], thick[k - 1], T[k - 1], psi_s[k], psi_l[k], psi_g[k],
                       thick[k], T[k]);
  }
  for (k = 1; k <= N_active; k++) H_abs[k] = H_abs[k] + (fl_Q[k + 1] - fl_Q[k]) * dt; /* :277-279 */
  for (k = 1; k <= N_active; k++) { /* :282-285 (sic: fl_rad(N_active)) */
    H_abs[k] = H_abs[k] + fl_rad[N_active] * dt;
    temp1 = temp1 + fl_rad[N_active] * dt;
  }

  if (c->thick_snow >= thick_min / 100.0 && c->thick_snow < thick_min) { /* :291-295 */
    EV(c, HEAT_THIN_SNOW);
    c->H_abs_snow = c->H_abs_snow - c->fl_q_snow * dt;
    snow_coupling(c, &c->H_abs_snow, &c->phi_s, &c->T_snow, &H_abs[1], &c->H[1], &c->phi[1], &T[1], c->m_snow,
                  c->S_abs_snow, m[1], c->S_bu[1]);
    temp1 = temp1 + c->fl_q_bottom * dt - c->fl_q_snow * dt;
  } else if (c->thick_snow >= thick_min) { /* :296-299 */
    c->H_abs_snow = c->H_abs_snow + (fl_Q[1] - c->fl_q_snow) * dt;
    temp1 = temp1 + c->fl_q_bottom * dt - c->fl_q_snow * dt;
  } else {
    temp1 = temp1 + c->fl_q_bottom * dt - fl_Q[1] * dt; /* :302 */
  }

  temp2 = sum_arr(H_abs, 1, Nlayer) + c->H_abs_snow; /* :305 */
  if (r_abs((temp1 - temp2) / dt) > 0.00001) SAM_STOP(c, 431); /* :307-310 */
}

/* ==========================================================================================
 * mo_grotz.f90:182-835 -- one iteration of the time loop
 * ======================================================================================== */
static void one_step(sam_col* c) {
  real *m = c->m, *S_abs = c->S_abs, *H_abs = c->H_abs, *thick = c->thick, *T = c->T, *S_bu = c->S_bu, *H = c->H,
       *phi = c->phi, *S_br = c->S_br, *psi_s = c->psi_s, *psi_l = c->psi_l, *psi_g = c->psi_g, *V_ex = c->V_ex;
  const real dt = c->dt;
  const int Nlayer = c->Nlayer;
  real temp, temp2, temp_2017_H, temp_2017_m;
  int k, jj;

  /* ---- S0 vital signs :192-223 ---- */
  c->energy_stored = c->H_abs_snow + sum_arr(H_abs, 1, c->N_active) - c->T_bottom * sum_arr(m, 1, c->N_active) * c_l;
  c->freshwater = sum_arr(m, 1, c->N_active) / rho_l;
  c->freshwater = c->freshwater * (1.0 - sum_arr(S_abs, 1, c->N_active) / sum_arr(m, 1, c->N_active) / ref_salinity);
  c->freshwater = c->freshwater + c->m_snow / rho_l;
  c->total_resist = 0.0;
  for (jj = 1; jj <= c->N_active - 1; jj++) c->total_resist = c->total_resist + thick[jj] / (psi_l[jj] * k_l + psi_s[jj] * k_s);
  c->total_resist = c->total_resist +
                    thick[c->N_active] * psi_s[c->N_active] / psi_s_min * (psi_s_min * k_s + 1.0 - psi_s_min * k_l);
  if (c->thick_snow > c->thick_min / 110.0) c->total_resist = c->total_resist + c->thick_snow / sam_func_k_snow(c->m_snow, c->thick_snow);
  if (c->N_active > 1) c->thickness = sum_arr(thick, 1, c->N_active - 1); else c->thickness = 0.0;
  c->thickness = c->thickness + thick[c->N_active] * psi_s[c->N_active] / psi_s_min;
  if (c->N_active > 1) {
    c->bulk_salin = sum_arr(S_abs, 1, c->N_active - 1) + S_abs[c->N_active] * psi_s[c->N_active] / psi_s_min;
    c->bulk_salin = c->bulk_salin / (sum_arr(m, 1, c->N_active - 1) + m[c->N_active] * psi_s[c->N_active] / psi_s_min);
  } else {
    c->bulk_salin = S_abs[1] / m[1];
  }

  /* ---- S1 forcing :229-246 ---- */
  if (c->atmoflux_flag == 2) {
    if (c->time > c->time_input[c->time_counter]) c->time_counter = c->time_counter + 1;
    if (c->time == c->time_input[c->time_counter]) {
      c->T2m = c->T2m_input[c->time_counter];
      c->liquid_precip = c->precip_input[c->time_counter];
    } else {
      const int tc = c->time_counter;
      temp = (c->time - c->time_input[tc - 1]) / (c->time_input[tc] - c->time_input[tc - 1]);
      c->T2m = (1.0 - temp) * c->T2m_input[tc - 1] + temp * c->T2m_input[tc];
      c->liquid_precip = (1.0 - temp) * c->precip_input[tc - 1] + temp * c->precip_input[tc];
    }
  }
  if (c->boundflux_flag == 3 && c->lab_snow_flag == 1) { /* :244-246 */
    c->solid_precip = c->precipinput[(long)floor(1 + c->time / dt)];
  }

  /* ---- S2 snow fall :251-265 ---- */
  if (c->precip_flag == 1) {
    if (r_max(c->liquid_precip, c->solid_precip) > 0.0 && c->N_active > 1) {
      EV(c, SNOW_PRECIP);
      snow_precip(&c->m_snow, &c->H_abs_snow, &c->thick_snow, dt, c->liquid_precip, c->T2m, 0, 0.0);
    } else if (r_max(c->liquid_precip, c->solid_precip) > 0.0 && c->N_active == 1) {
      EV(c, SNOW_PRECIP_0);
      snow_precip_0(&H_abs[1], &S_abs[1], m[1], T[1], dt, c->liquid_precip, c->T2m, 0, 0.0);
    }
  } else if (c->precip_flag == 0) {
    if (r_max(c->liquid_precip, c->solid_precip) > 0.0 && c->N_active > 1) {
      EV(c, SNOW_PRECIP);
      snow_precip(&c->m_snow, &c->H_abs_snow, &c->thick_snow, dt, c->liquid_precip, c->T2m, 1, c->solid_precip);
    } else if (r_max(c->liquid_precip, c->solid_precip) > 0.0 && c->N_active == 1) {
      EV(c, SNOW_PRECIP_0);
      snow_precip_0(&H_abs[1], &S_abs[1], m[1], T[1], dt, c->liquid_precip, c->T2m, 1, c->solid_precip);
    }
  }

  /* ---- S3 snow thermodynamics :273-292 ---- */
  snow_thermo_block(c);

  /* ---- S4 inner layer thermodynamics and expulsion :298-307 ---- */
  c->T_test = c->T_bottom;
  for (k = c->N_active; k >= 1; k--) {
    S_bu[k] = S_abs[k] / m[k];
    H[k] = H_abs[k] / m[k];
    sam_getT(c, H[k], S_bu[k], c->T_test, &T[k], &phi[k], k);
    c->T_test = T[k];
    S_br[k] = sam_func_S_br2(c, T[k], S_bu[k]);
    Expulsion(phi[k], thick[k], m[k], &psi_s[k], &psi_l[k], &psi_g[k], &V_ex[k]);
  }

  /* ---- S5 brine flux due to expulsion :312-321 ---- */
  expulsion_flux(c);
  if (c->i != 1) {
    mass_transfer(c, T, H_abs, S_abs, S_bu, c->fl_m);
    if (c->bgc_flag == 2) { /* :316-320 */
      for (k = 1; k <= c->N_active; k++) FB(c, k, k + 1) = -c->fl_m[k + 1];
    }
  }

  /* ---- S7 :333-335 ---- */
  for (k = c->N_active; k >= 1; k--) S_bu[k] = S_abs[k] / m[k];

  /* ---- S8 standard output :340-398 ---- */
  if (c->n_time_out == c->i_time_out || c->i == 1) {
    if (c->N_active > 1) c->freeboard = sam_func_freeboard(c); else c->freeboard = 0.0;
    if (c->grav_flag == 2) {
      if (c->grav_drain == 0.0) c->grav_temp = 0.0; else c->grav_temp = c->grav_temp / c->grav_drain;
      c->grav_salt = c->grav_salt / c->time_out;
      c->grav_drain = c->grav_drain / c->time_out;
    }
    if (c->on_output) c->on_output(c, c->on_output_user);
    c->n_outputs++;
    c->grav_drain = 0.0;
    c->grav_salt = 0.0;
    c->grav_temp = 0.0;
    c->melt_thick_output[1] = 0.0; c->melt_thick_output[2] = 0.0; c->melt_thick_output[3] = 0.0;
    c->n_time_out = 0;
  } else {
    c->n_time_out = c->n_time_out + 1;
  }

  /* ---- S9 gas in the lowest layer :405-410 ---- */
  if (psi_g[c->N_active] > 0.0) {
    EV(c, GAS_REFILL);
    temp2 = psi_g[c->N_active] * thick[c->N_active] * rho_l;
    m[c->N_active] = m[c->N_active] + temp2;
    S_abs[c->N_active] = S_abs[c->N_active] + temp2 * c->S_bu_bottom;
    H_abs[c->N_active] = H_abs[c->N_active] + temp2 * c_l * c->T_bottom;
  }

  /* ---- S10 thin snow coupling :418-420 ---- */
  if (c->m_snow > 0.0 && c->thick_snow < c->thick_min) {
    snow_coupling(c, &c->H_abs_snow, &c->phi_s, &c->T_snow, &H_abs[1], &H[1], &phi[1], &T[1], c->m_snow, c->S_abs_snow,
                  m[1], S_bu[1]);
  }

  /* ---- S11 flooding :428-445 ---- */
  if (c->N_active > 1 && c->flood_flag > 1) {
    c->freeboard = sam_func_freeboard(c);
    if (c->freeboard < 0.0) {
      if (c->flood_flag == 2) {
        flood(c);
      } else if (c->flood_flag == 3 && c->freeboard < neg_free) {
        flood_simple(c);
      }
    }
  }

  /* ---- S12 turbulence :450-457 ---- */
  if (c->turb_flag == 2) {
    EV(c, TURB);
    if (c->bgc_flag == 2) { /* :451-453, mo_functions.f90:355-360: turb from the S_abs before its update */
      const int Na = c->N_active;
      int q;
      const real turb = Turb_A * M_EXP(Turb_B * (-sam_func_density(c->T_bottom, c->S_bu_bottom) + sam_func_density(T[Na], S_abs[Na] / m[Na]))) * dt;
      S_abs[Na] = S_abs[Na] - turb * (S_abs[Na] / m[Na] - c->S_bu_bottom);
      for (q = 1; q <= c->N_bgc; q++) c->bgc_abs[q][Na] = c->bgc_abs[q][Na] - turb * (c->bgc_abs[q][Na] / m[Na] - c->bgc_bottom[q]);
    } else
    sub_turb_flux(c->T_bottom, c->S_bu_bottom, T[c->N_active], &S_abs[c->N_active], m[c->N_active], dt);
  }

  /* ---- S13 gravity drainage :463-477 ---- */
  if (c->grav_flag == 2 && c->N_active > 1) {
    fl_grav_drain(c);
  } else if (c->grav_flag == 3 && c->N_active > 1) {
    EV(c, GRAV_DRAIN_SIMPLE);
    fl_grav_drain_simple(c);
  }

  /* ---- S14 prescribed salinity :482-497 (prescribe_flag 2; none of the configs) ---- */
  if (c->prescribe_flag == 2) {
    EV(c, PRESCRIBE);
    k = c->N_active;
    while (k > 1 && sum_arr(thick, k, c->N_active) < 0.15) {
      S_bu[k] = c->S_bu_bottom - sum_arr(thick, k, c->N_active) / 0.15 * (c->S_bu_bottom - 4.0);
      k = k - 1;
    }
    while (k > 1 && sum_arr(thick, k, c->N_active) >= 0.15) {
      S_bu[k] = 4.0 - 4.0 * (sum_arr(thick, k, c->N_active) - 0.15) / (sum_arr(thick, 1, c->N_active) - 0.15);
      k = k - 1;
      S_bu[1] = 0.0;
    }
    S_bu[c->N_active] = c->S_bu_bottom;
    for (k = 1; k <= Nlayer; k++) S_abs[k] = S_bu[k] * m[k];
  }

  /* ---- S15 testcase hooks :503-563 ---- */
  if (c->testcase == 1) {
    sub_test1(c);
  } else if (c->testcase >= 101 && c->testcase <= 105) { /* :521-530 */
    const long idx = (long)floor(1 + c->time / dt);
    const real Sb = S_bu[c->N_active + 1];
    c->T2m = c->Tinput[idx];
    c->solid_precip = c->precipinput[idx];
    c->fl_q_bottom = c->ocean_flux_input[idx];
    c->T_bottom = -F32(0.0575) * Sb + F32(1.710523e-3) * M_POW(Sb, 3.0 / 2.0) - F32(2.154996e-4) * P2(Sb) -
                  F32(7.53e-4) * sum_arr(thick, 1, c->N_active - 1);
    c->styropor_flag = (int)c->styropor_input[idx];
  } else if (c->testcase == 2) {
    sub_test2(c);
  } else if (c->testcase == 9) {
    sub_test9(c);
  } else if (c->testcase == 34) {
    sub_test34(c);
  } else if (c->testcase == 99) {
    hook_test99(c);
  } else if (c->testcase == 3) {
    sub_test3(c);
  } else if (c->testcase == 4 || c->testcase == 7) {
    sub_test4(c);
  } else if (c->testcase == 6) {
    sub_test6(c);
  } else if (c->testcase == 111) { /* :505-506 */
    const long idx = (long)floor(1 + c->time / c->dt);
    if (idx < 1 || idx > c->length_input_lab || !c->Tinput) SAM_STOP(c, 8001); /* the reference would read Ttop_input out of bounds */
    c->T_top = c->Tinput[idx];
  } else if (c->testcase == 8) { /* :539-544 */
    if (c->time < (double)(3600.f * 12.f * 11.f)) {
      const long idx = (long)floor(1 + c->time / 60);
      if (idx < 1 || idx > c->length_input_lab) SAM_STOP(c, 8001); /* the reference would read Tinput out of bounds */
      c->T_top = c->Tinput[idx];
    } else {
      c->T_top = -15.0;
    }
  } else if (c->testcase == 5 && c->i == 2) { /* :541-542 */
    for (k = 1; k <= Nlayer; k++) S_abs[k] = 5.0 * m[k];
  }

  /* ---- S16 tank :573-578 ---- */
  if (c->tank_flag == 2) {
    EV(c, TANK);
    c->S_bu_bottom = (c->S_total - sum_arr(S_abs, 1, Nlayer)) / (c->m_total - sum_arr(m, 1, Nlayer));
    if (c->bgc_flag == 2) { /* :575-577 (sic: every tracer gets the value computed from tracer 1) */
      int q;
      const real v = (c->bgc_total[1] - sum_arr(c->bgc_abs[1], 1, Nlayer)) / (c->m_total - sum_arr(m, 1, Nlayer));
      for (q = 1; q <= c->N_bgc; q++) c->bgc_bottom[q] = v;
    }
  }

  /* ---- S17 heat fluxes :584 ---- */
  sub_heat_fluxes(c);

  /* ---- S18 :592-598 ---- */
  c->T_test = c->T_bottom;
  for (k = c->N_active; k >= 1; k--) {
    S_bu[k] = S_abs[k] / m[k];
    H[k] = H_abs[k] / m[k];
    sam_getT(c, H[k], S_bu[k], c->T_test, &T[k], &phi[k], k);
    c->T_test = T[k];
  }

  /* ---- S19 :600-625 ---- */
  temp_2017_H = H_abs[1] + c->H_abs_snow + c->melt_thick_snow * rho_l * c_l * c->T_snow;
  temp_2017_m = m[1] + c->m_snow + c->melt_thick_snow * rho_l;
  (void)temp_2017_H; (void)temp_2017_m; /* only feed PRINT statements (:688-692) */
  c->melt_thick_snow_old = c->melt_thick_snow;
  snow_thermo_block(c);
  c->melt_thick_snow = c->melt_thick_snow_old + c->melt_thick_snow;

  /* ---- S20 flushing preparations :632-664 ---- */
  if (c->N_active > 1 && c->flush_flag > 2) {
    if (c->boundflux_flag == 2) {
      c->T_freeze = sam_func_T_freeze(S_abs[1] / m[1], c->salt_flag);
      c->melt_thick = 0.0;
      if (sam_func_freeboard(c) > 0.0000000000001) {
        if (psi_s[1] < psi_s_top_min || c->T_top >= c->T_freeze) {
          EV(c, MELT_THICK);
          if (sub_melt_thick(psi_l[1], psi_s[1], psi_g[1], T[1], c->T_freeze, c->T_top, c->fl_Q[1], c->thick_snow, dt,
                             &c->melt_thick, &thick[1], c->thick_min)) EV(c, MELT_THICK_GAS);
          if (c->thick_snow >= c->thick_min / 100.0 && c->melt_thick > 0.00000000001 && c->melt_thick_snow == 0.0) {
            if (sub_melt_snow(&c->melt_thick, &thick[1], &c->thick_snow, &H_abs[1], &c->H_abs_snow, &m[1], &c->m_snow,
                              &c->psi_g_snow)) EV(c, MELT_SNOW_ALL); else EV(c, MELT_SNOW_PART);
          }
        }
      }
    }
    if (c->boundflux_flag == 3) {
      c->T_freeze = sam_func_T_freeze(S_abs[1] / m[1], c->salt_flag);
      c->melt_thick = 0.0;
      if (sam_func_freeboard(c) > 0.0000000000001) {
        if (psi_s[1] < psi_s_top_min || c->T2m >= c->T_freeze) {
          EV(c, MELT_THICK);
          if (sub_melt_thick(psi_l[1], psi_s[1], psi_g[1], T[1], c->T_freeze, c->T2m, c->fl_Q[1], c->thick_snow, dt,
                             &c->melt_thick, &thick[1], c->thick_min)) EV(c, MELT_THICK_GAS);
          c->melt_thick = r_max(c->melt_thick, 0.0);
          if (c->thick_snow >= c->thick_min / 100.0 && c->melt_thick > 0.00000000001 && c->melt_thick_snow == 0.0) {
            if (sub_melt_snow(&c->melt_thick, &thick[1], &c->thick_snow, &H_abs[1], &c->H_abs_snow, &m[1], &c->m_snow,
                              &c->psi_g_snow)) EV(c, MELT_SNOW_ALL); else EV(c, MELT_SNOW_PART);
          }
        }
      }
    }
  }

  /* ---- S21 flushing :670-737 ---- */
  c->freeboard = sam_func_freeboard(c);
  c->melt_thick_output[1] = c->melt_thick_output[1] + c->melt_thick;
  c->melt_thick_output[2] = c->melt_thick_output[2] + c->melt_thick_snow;
  c->melt_thick = c->melt_thick + c->melt_thick_snow;
  if (c->melt_thick_snow > 0.0) { /* :677-685 */
    EV(c, SNOW_MELTWATER_TO_ICE);
    H_abs[1] = H_abs[1] + c->melt_thick_snow * rho_l * c_l * c->T_snow;
    S_abs[1] = S_abs[1] + c->melt_thick_snow * rho_l * sam_func_S_br2(c, c->T_snow, c->S_abs_snow / c->m_snow);
    thick[1] = thick[1] + c->melt_thick_snow;
    m[1] = m[1] + c->melt_thick_snow * rho_l;
    S_bu[1] = S_abs[1] / m[1];
    H[1] = H_abs[1] / m[1];
  }
  for (k = 1; k <= Nlayer; k++) { /* :697-701 */
    c->flush_v_old[k] = c->flush_v[k];
    c->flush_h_old[k] = c->flush_h[k];
    c->flush_v[k] = 0.0;
    c->flush_h[k] = 0.0;
  }
  if (c->N_active > 1 && c->freeboard > 0.001) {
    if (c->flush_flag == 4) { /* :704-713 */
      if (c->melt_thick > 0.000000000001 && c->N_active > 2) {
        EV(c, FLUSH_INLINE);
        H_abs[1] = H_abs[1] - c->melt_thick * rho_l * c_l * T[1];
        S_abs[1] = S_abs[1] * (1.0 - (c->melt_thick * rho_l) / m[1]);
        thick[1] = thick[1] - c->melt_thick;
        m[1] = m[1] - c->melt_thick * rho_l;
      }
    } else if (c->flush_flag == 5) { /* :715-728 */
      if (c->melt_thick > 0.000000000001 && c->N_active > 2 && c->freeboard > 0.0) {
        c->freeboard = sam_func_freeboard(c);
        flush3(c);
      }
    } else if (c->flush_flag == 6) { /* :729-733 */
      if (c->melt_thick > 0.000000000001 && c->N_active > 2 && c->thick_snow < c->thick_0) flush4(c);
    }
  }
  for (k = 1; k <= Nlayer; k++) { /* :736-737 */
    c->flush_v[k] = c->flush_v[k] + c->flush_v_old[k];
    c->flush_h[k] = c->flush_h[k] + c->flush_h_old[k];
  }

  /* ---- S22 tracer advection :742-747 ---- */
  if (c->bgc_flag == 2) {
    bgc_advection(c);
    memset(c->fl_brine_bgc, 0, sizeof(real) * (size_t)(Nlayer + 2) * (size_t)(Nlayer + 2));
  }

  /* ---- S23 layer dynamics :755-795 ---- */
  if (c->N_active > 1) {
    if (phi[c->N_active] > psi_s_min || phi[c->N_active - 1] <= psi_s_min / 2.0 || thick[1] / c->thick_0 > 1.5 ||
        thick[1] / c->thick_0 < 0.5) {
      layer_dynamics(c);
    }
    {
      const int kn = (c->N_active + 1 < Nlayer) ? c->N_active + 1 : Nlayer;
      if (c->N_active < Nlayer && thick[kn] == 0) { /* :772-783 */
        EV(c, SCRUB);
        T[c->N_active + 1] = c->T_bottom;
        S_bu[c->N_active + 1] = c->S_bu_bottom;
        H[c->N_active + 1] = 0.0;
        psi_l[c->N_active + 1] = 1.0;
        psi_s[c->N_active + 1] = 0.0;
        if (c->bgc_flag == 2) { /* :778-780 */
          int q;
          for (q = 1; q <= c->N_bgc; q++) c->bgc_abs[q][c->N_active + 1] = 0.0;
        }
      }
    }
  } else {
    if (phi[1] > psi_s_min) layer_dynamics(c);
  }

  /* ---- S24 timestep and health check :802-819 ---- */
  c->time = c->time + dt;
  {
    real mn = psi_s[1], ms = S_abs[1];
    for (k = 1; k <= c->N_active; k++) {
      mn = r_min(mn, psi_s[k]);
      ms = r_min(ms, S_abs[k]);
    }
    if (mn < 0.0) {
      SAM_STOP(c, 1337);
    } else if (ms < 0.0) {
      EV(c, SALT_CLAMP);
      for (k = 1; k <= c->N_active; k++) S_abs[k] = r_max(S_abs[k], 0.0);
    }
  }
}

int sam_step(sam_col* c, long nsteps) {
  long s;
  if (c->status != 0) return c->status;
  if (setjmp(c->jb) != 0) return c->status;
  for (s = 0; s < nsteps; s++) {
    c->i = c->i + 1;
    one_step(c);
  }
  return 0;
}

/* ==========================================================================================
 * mo_init.f90
 * ======================================================================================== */
static real* alloc_arr(int n) { return (real*)calloc((size_t)(n + 3), sizeof(real)); }

static void sub_allocate(sam_col* c, int Nlayer) { /* mo_init.f90:2040-2088 */
  c->H = alloc_arr(Nlayer); c->H_abs = alloc_arr(Nlayer); c->T = alloc_arr(Nlayer);
  c->S_abs = alloc_arr(Nlayer); c->S_bu = alloc_arr(Nlayer + 1); c->S_br = alloc_arr(Nlayer);
  c->thick = alloc_arr(Nlayer); c->m = alloc_arr(Nlayer); c->V_ex = alloc_arr(Nlayer); c->phi = alloc_arr(Nlayer);
  c->perm = alloc_arr(Nlayer); c->flush_v = alloc_arr(Nlayer); c->flush_h = alloc_arr(Nlayer);
  c->flush_v_old = alloc_arr(Nlayer); c->flush_h_old = alloc_arr(Nlayer);
  c->psi_s = alloc_arr(Nlayer); c->psi_l = alloc_arr(Nlayer); c->psi_g = alloc_arr(Nlayer); c->fl_rad = alloc_arr(Nlayer);
  c->fl_Q = alloc_arr(Nlayer + 1); c->fl_m = alloc_arr(Nlayer + 1); c->ray = alloc_arr(Nlayer);
  {
    int q;
    for (q = 0; q < 8; q++) c->scr[q] = alloc_arr(Nlayer + 1); /* automatic arrays of the callees */
  }
  /* sub_allocate_bgc (mo_init.f90:2093-2110); allocated for every column so that bgc_flag can be switched on by a test */
  c->bgc_abs[0] = NULL; c->bgc_abs[1] = alloc_arr(Nlayer + 1); c->bgc_abs[2] = alloc_arr(Nlayer + 1);
  c->fl_brine_bgc = (real*)calloc((size_t)(Nlayer + 2) * (size_t)(Nlayer + 2), sizeof(real));
}

sam_col* sam_create(int testcase) {
  sam_col* c;
  int k;
  int is_lab = (testcase >= 101 && testcase <= 105);
  if (!(testcase == 1 || testcase == 2 || testcase == 3 || testcase == 4 || testcase == 5 || testcase == 6 || testcase == 7 ||
        testcase == 8 || testcase == 9 || testcase == 33 || testcase == 34 || testcase == 50 || testcase == 99 || testcase == 111 ||
        is_lab))
    return NULL;
  c = (sam_col*)calloc(1, sizeof(sam_col));
  c->testcase = testcase;
  /* defaults, mo_init.f90:83-132 */
  c->boundflux_flag = 1; c->atmoflux_flag = 1; c->albedo_flag = 2;
  c->grav_heat_flag = 1; c->flush_heat_flag = 1; c->flood_flag = 2; c->flush_flag = 5; c->grav_flag = 2; c->harmonic_flag = 2;
  c->prescribe_flag = 1; c->salt_flag = 1;
  c->turb_flag = 2; c->bottom_flag = 1; c->tank_flag = 1;
  c->precip_flag = 0; c->freeboard_snow_flag = 0; c->snow_flush_flag = 1; c->snow_precip_flag = 1;
  c->debug_flag = 1; c->bgc_flag = 1;
  c->lab_snow_flag = 0; c->styropor_flag = 0; /* not initialised by init: zero storage */
  c->max_flux_plate = 10000.0; c->k_snow_flush = 0.75; c->k_styropor = 0.8; /* mo_parameters.f90:107-112 */
  c->ttop_warm = -5.0; c->ttop_cold = -10.0; c->oflux_amp = 7.0;

  if (testcase == 1) { /* mo_init.f90:865-945 */
    c->Nlayer = 90; c->N_active = 1; c->N_top = 5; c->N_bottom = 5;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->turb_flag = 1; c->boundflux_flag = 1; c->grav_heat_flag = 1; c->flush_flag = 1; c->salt_flag = 2;
    c->T_top = -5.0; c->T_bottom = -1.; c->S_bu_bottom = 34.; c->fl_q_bottom = 0.0;
    c->thick_0 = 0.002; c->dt = 1.0; c->time = 0.0; c->time_out = 3600.0; c->time_total = c->time_out * 72.0;
    c->thick[1] = c->thick_0;
    c->m[1] = c->thick[1] * rho_l;
    c->S_abs[1] = c->S_bu_bottom * c->m[1];
    c->H_abs[1] = c->m[1] * (c->T_bottom) * c_l;
    c->bgc_flag = 2; /* :921-944 */
    c->N_bgc = 2;
    c->bgc_bottom[1] = 400.0; c->bgc_bottom[2] = 500.0;
    c->bgc_abs[1][1] = c->bgc_bottom[1] * c->m[1]; c->bgc_abs[2][1] = c->bgc_bottom[2] * c->m[1];
  } else if (testcase == 8) { /* mo_init.f90:1451-1494: field temperatures (input/DNotz_fieldT/Tinput.txt, one value per
                               * minute) prescribe T_top.  The reference marks the settings "likely outdated": its init
                               * does not allocate Tinput and mo_grotz.f90:138-143 builds a file name from testcase-100; the
                               * time loop itself (:539-544) is well defined once Tinput holds the series, and that is what
                               * is restated: the caller supplies the series through sam_set_lab_forcing (kind Tice). */
    c->Nlayer = 50; c->N_active = 1; c->N_bottom = 5; c->N_top = 4;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->T_top = -5.0; c->T_bottom = F32(-1.8); c->S_bu_bottom = 34.0; /* :1461 T_bottom = -1.8 is a default-REAL literal */
    c->boundflux_flag = 1; c->fl_q_bottom = 15.0;
    c->grav_flag = 2; c->flush_flag = 5; c->flood_flag = 2;
    c->thick_0 = 0.005;
    c->thick[1] = c->thick_0;
    c->m[1] = c->thick[1] * rho_l;
    c->S_abs[1] = c->S_bu_bottom * c->m[1];
    c->H_abs[1] = c->m[1] * (c->T_bottom) * c_l;
    c->time = 0.0; c->time_out = 3600.0; c->time_total = c->time_out * 12.0 * 12.0; c->dt = 1.0;
  } else if (testcase == 111) { /* mo_init.f90:141-221: salinity-harp comparison; the uppermost harp sensor prescribes T_top, one
                                 * value per time step (2017_input/Ts_<dt>s.txt, mo_grotz.f90:171-176, not shipped): the caller
                                 * supplies the series through sam_set_lab_forcing (kind Tice) */
    c->Nlayer = 100; c->N_active = 1; c->N_top = 10; c->N_bottom = 10;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    c->length_input_lab = 860333;
    sub_allocate(c, c->Nlayer);
    c->turb_flag = 1; c->boundflux_flag = 1; c->grav_heat_flag = 1; c->flush_flag = 1; c->salt_flag = 2;
    c->T_top = -2.0; c->T_bottom = -1.67; c->S_bu_bottom = 33.4079; c->fl_q_bottom = 0.;
    c->thick_0 = 0.01; c->dt = 3.0; c->time = 0.0; c->time_out = 3600.0 * 2.0; c->time_total = 2580996.0;
    c->thick[1] = c->thick_0;
    c->m[1] = c->thick[1] * rho_l;
    c->S_abs[1] = c->S_bu_bottom * c->m[1];
    c->H_abs[1] = c->m[1] * (c->T_bottom) * c_l;
    c->bgc_flag = 1;
  } else if (testcase == 4) { /* mo_init.f90:1127-1207 */
    c->Nlayer = 100; c->N_bottom = 20; c->N_top = 20; c->N_active = 1;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->atmoflux_flag = 2; c->precip_flag = 1; c->boundflux_flag = 2; c->snow_flush_flag = 1; c->flush_heat_flag = 2;
    c->snow_precip_flag = 1;
    c->T_bottom = -1.0; c->S_bu_bottom = 34.0;
    c->thick_0 = 0.01; c->time = 0.0; c->time_out = 86400.0; c->time_total = c->time_out * 365.0 * 4.5; c->dt = 10.0;
    c->thick[1] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->S_bu_bottom * c->m[k];
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = 0.0;
    c->bgc_flag = 1;
  } else if (testcase == 33 || testcase == 34 || testcase == 99) { /* more cooling-chamber tanks, mo_init.f90:1779-1873, 1876-1970, 768-862 */
    c->fl_q_bottom = (testcase == 99) ? 5.0 : 10.0;
    c->alpha_flux_instable = 22.0; c->alpha_flux_stable = 15.0; c->tank_depth = 0.94;
    if (testcase == 99) { c->Nlayer = 20; c->N_bottom = 5; c->N_top = 5; }
    else { c->Nlayer = 100; c->N_bottom = 10; c->N_top = 3; }
    c->N_active = 1;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->tank_flag = 2; c->boundflux_flag = 3; c->grav_heat_flag = 1;
    if (testcase == 99) { c->precip_flag = 0; c->flush_flag = 1; c->flood_flag = 1; c->grav_flag = 2; }
    if (testcase == 33) {        /* fresh water */
      c->T2m = -15.0; c->T_top = -10.0; c->T_bottom = 0.5; c->S_bu_bottom = 0.13;
      c->thick_0 = 0.005; c->time_out = 60.0 * 5.0; c->time_total = c->time_out * 12.0 * 6.0; c->dt = 10.0;
    } else if (testcase == 34) {
      c->T2m = -15.0; c->T_top = -10.0; c->T_bottom = 0.5; c->S_bu_bottom = 34.9;
      c->thick_0 = 0.005; c->time_out = 60.0 * 10.0; c->time_total = 86400.0 * 10.0; c->dt = 10.0;
    } else {                     /* snow on ice in the chamber */
      c->T2m = -5.0; c->T_top = -2.0; c->T_bottom = -1.8; c->S_bu_bottom = 34.0;
      c->thick_0 = 0.05; c->time_out = 60.0 * 10.0; c->time_total = 3600.0 * 24.0 * 7.0; c->dt = 10.0;
    }
    c->m_total = rho_l * c->tank_depth;
    c->S_total = rho_l * c->S_bu_bottom * c->tank_depth;
    c->thick[1] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->S_bu_bottom * c->m[k];
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = c->m[k] * c->T_bottom; /* sic: no c_l */
    c->bgc_flag = 1;
  } else if (testcase == 50) { /* mo_init.f90:1497-1532: spin-up of a stable profile (Griewank & Notz 2012) */
    c->fl_sw = 0.0; c->boundflux_flag = 2;
    c->fl_rest = sigma * 4106877291.8310046; /* sigma*(zeroK-20._wp)**4._wp, the power folded at compile time (correctly rounded) */
    c->fl_q_bottom = 20.0;
    c->Nlayer = 70; c->N_bottom = 5; c->N_top = 5;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    c->T_top = -20.0; c->T_bottom = -1.72; c->S_bu_bottom = 34.0; c->N_active = 1;
    sub_allocate(c, c->Nlayer);
    c->thick_0 = 0.005;
    c->thick[1] = c->thick_0;
    c->m[1] = c->thick[1] * rho_l;
    c->S_abs[1] = c->S_bu_bottom * c->m[1];
    c->H_abs[1] = c->m[1] * (c->T_bottom) * c_l;
    c->time = 0.0; c->time_out = 3600.0 * 24.0 * 30.0; c->dt = 10.0; c->time_total = c->time_out * 12.0 * 3.0;
  } else if (testcase == 2 || testcase == 6 || testcase == 9) { /* cooling-chamber tanks, mo_init.f90:948-1042, 1278-1357, 1684-1776 */
    if (testcase == 2) {
      c->fl_q_bottom = 10.0; c->alpha_flux_instable = 22.0; c->alpha_flux_stable = 15.0; c->tank_depth = 1.0;
      c->Nlayer = 100; c->N_bottom = 10; c->N_top = 3;
    } else if (testcase == 6) {
      c->fl_q_bottom = 35.0; c->alpha_flux_instable = 22.0; c->alpha_flux_stable = 11.0; c->tank_depth = 0.159;
      c->Nlayer = 40; c->N_bottom = 3; c->N_top = 3;
    } else {
      c->fl_q_bottom = 10.0; c->alpha_flux_instable = 22.0; c->alpha_flux_stable = 15.0; c->tank_depth = 0.8;
      c->Nlayer = 100; c->N_bottom = 10; c->N_top = 3;
    }
    c->N_active = 1;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->tank_flag = 2; c->boundflux_flag = 3; c->grav_heat_flag = 1;
    if (testcase == 2) {
      c->T2m = -20.0; c->T_top = -18.0; c->T_bottom = 0.0; c->S_bu_bottom = 31.2;
      c->thick_0 = 0.01; c->time_out = 3600.0 * 6.0; c->time_total = c->time_out * 4.0 * 30.0; c->dt = 30.0;
    } else if (testcase == 6) {
      c->T2m = -18.0; c->T_top = -18.0; c->T_bottom = 0.0; c->S_bu_bottom = 31.2;
      c->thick_0 = 0.0025; c->time_out = 1800.0 / 2.0; c->time_total = c->time_out * 39.0 * 2.0 * 2.0; c->dt = 0.5;
    } else {
      c->T2m = -15.0; c->T_top = -10.0; c->T_bottom = -0.07; c->S_bu_bottom = 34.6;
      c->thick_0 = 0.005; c->time_out = 3600.0 * 2.0; c->time_total = c->time_out * 12.0 * 6.0; c->dt = 10.0;
    }
    c->m_total = rho_l * c->tank_depth;
    c->S_total = rho_l * c->S_bu_bottom * c->tank_depth;
    c->thick[1] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->S_bu_bottom * c->m[k];
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = c->m[k] * c->T_bottom;
    c->bgc_flag = (testcase == 9) ? 1 : 2;
    if (c->bgc_flag == 2) { /* :1016-1040 (2: two tracers), :1331-1355 (6: one tracer) */
      int q;
      c->N_bgc = (testcase == 2) ? 2 : 1;
      for (q = 1; q <= c->N_bgc; q++) {
        c->bgc_bottom[q] = 385.0;
        c->bgc_total[q] = c->bgc_bottom[q] * rho_l * c->tank_depth;
        c->bgc_abs[q][1] = c->bgc_bottom[q] * c->m[1];
      }
    }
  } else if (testcase == 3) { /* mo_init.f90:1045-1124: climatological forcing (notzflux), constant oceanic heat flux */
    c->Nlayer = 20; c->N_bottom = 5; c->N_top = 5; c->N_active = 1;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->atmoflux_flag = 1; c->precip_flag = 0; c->boundflux_flag = 2;
    c->fl_q_bottom = 8.0; c->T_bottom = -1.0; c->S_bu_bottom = 34.0;
    c->thick_0 = 0.03; c->time = 0.0; c->time_out = 86400.0 * 3.5; c->time_total = c->time_out * 54.0 * 2.0 * 2.0; c->dt = 60.0;
    c->thick[1] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->S_bu_bottom * c->m[k];
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = 0.0;
    c->bgc_flag = 1;
  } else if (testcase == 5) { /* mo_init.f90:1210-1275: top melt of a 1 m block of cold fresh ice */
    c->Nlayer = 100; c->N_active = c->Nlayer; c->N_bottom = 10; c->N_top = 20;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->boundflux_flag = 2; c->atmoflux_flag = 3; c->flush_heat_flag = 2;
    c->flush_flag = 5; c->grav_flag = 1; c->flood_flag = 1;
    c->fl_sw = 0.0; c->fl_rest = 7072810000.0 * sigma; /* 290._wp**4*sigma, an integer power folded at compile time */ c->fl_q_bottom = 15.0;
    c->S_bu_bottom = 5.0; c->T_bottom = 0.;
    c->thick_0 = 0.01; c->time = 0.0; c->time_out = 3600.0 * 3.0; c->time_total = c->time_out * 24.0 * 10.0; c->dt = 10.0;
    for (k = 1; k <= c->Nlayer; k++) c->thick[k] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->m[k] * c->S_bu_bottom;
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = c->m[k] * (-90.0) * c_l;
    c->bgc_flag = 1;
  } else if (testcase == 7) { /* mo_init.f90:1360-1448: testcase 4 with the simple parametrisations */
    c->Nlayer = 100; c->N_bottom = 20; c->N_top = 20; c->N_active = 1;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    sub_allocate(c, c->Nlayer);
    c->atmoflux_flag = 2; c->precip_flag = 1; c->boundflux_flag = 2;
    c->albedo_flag = 1; c->grav_heat_flag = 2; c->flush_heat_flag = 2;
    c->flush_flag = 4; c->grav_flag = 3; c->flood_flag = 3;
    c->T_bottom = -1.0; c->S_bu_bottom = 34.0;
    c->thick_0 = 0.01; c->time = 0.0; c->time_out = 86400.0 / 2.0; c->time_total = c->time_out * 365.0 * 9.0; c->dt = 10.0;
    c->thick[1] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->S_bu_bottom * c->m[k];
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = 0.0;
    c->bgc_flag = 1;
  } else { /* testcases 101-105, mo_init.f90:222-767 */
    static const double Sb[5] = {25.6664555556, 26.1336777778, 26.0335888889, 27.0363, 31.5625333333};
    static const double tt[5] = {1625000.0, 1124000.0, 1283000.0, 2439000.0, 1549000.0};
    static const long len[5] = {1628263, 1124187, 1283092, 2439729, 1549323};
    const int q = testcase - 101;
    c->fl_q_bottom = 0.0; c->alpha_flux_instable = 22.0; c->alpha_flux_stable = 21.0; c->tank_depth = 0.94;
    c->Nlayer = 200; c->N_bottom = 10; c->N_top = 5; c->N_active = 1;
    c->N_middle = c->Nlayer - c->N_top - c->N_bottom;
    c->length_input_lab = len[q];
    sub_allocate(c, c->Nlayer);
    c->tank_flag = 2; c->boundflux_flag = 3; c->precip_flag = 0; c->grav_heat_flag = 1; c->flush_flag = 5; c->flood_flag = 2;
    c->grav_flag = 2; c->lab_snow_flag = 1; c->freeboard_snow_flag = 1; c->snow_flush_flag = 1; c->flush_heat_flag = 2;
    c->snow_precip_flag = 1;
    c->T2m = 0.0; c->T_top = 0.0; c->T_bottom = -1.3; c->S_bu_bottom = Sb[q];
    c->thick_0 = 0.01; c->time = 0.0; c->time_out = 60.0 * 60.0; c->time_total = tt[q]; c->dt = 1.0;
    c->m_total = rho_l * c->tank_depth;
    c->S_total = rho_l * c->S_bu_bottom * c->tank_depth;
    c->thick[1] = c->thick_0;
    for (k = 1; k <= c->Nlayer; k++) c->m[k] = c->thick[k] * rho_l;
    for (k = 1; k <= c->Nlayer; k++) c->S_abs[k] = c->S_bu_bottom * c->m[k];
    for (k = 1; k <= c->Nlayer; k++) c->H_abs[k] = c->m[k] * c->T_bottom; /* sic: no c_l (:299) */
    c->bgc_flag = 1;
  }

  /* common tail, mo_init.f90:1982-2009 */
  for (k = 1; k <= c->Nlayer; k++) {
    c->T[k] = c->T_bottom;
    c->S_bu[k] = c->S_bu_bottom;
    c->psi_s[k] = 0.0;
    c->phi[k] = 0.0;
    c->psi_l[k] = 1.0;
    c->fl_rad[k] = 0.0;
  }
  c->thick_min = c->thick_0 / 2.0;
  c->i_time = (int)(c->time_total / c->dt);
  c->i_time_out = (int)(c->time_out / c->dt);
  c->n_time_out = 0;
  c->melt_thick = 0;
  c->thickness = 0.0;
  c->bulk_salin = sum_arr(c->S_abs, 1, c->N_active) / sum_arr(c->m, 1, c->N_active);
  c->time_counter = 1; /* mo_grotz.f90:133 */
  c->i = 0;
  return c;
}

void sam_destroy(sam_col* c) {
  if (!c) return;
  free(c->H); free(c->H_abs); free(c->T); free(c->S_abs); free(c->S_bu); free(c->S_br); free(c->thick); free(c->m);
  free(c->V_ex); free(c->phi); free(c->perm); free(c->flush_v); free(c->flush_h); free(c->flush_v_old); free(c->flush_h_old);
  free(c->psi_s); free(c->psi_l); free(c->psi_g); free(c->fl_rad); free(c->fl_Q); free(c->fl_m); free(c->ray);
  free(c->time_input); free(c->T2m_input); free(c->precip_input); free(c->fl_sw_input); free(c->fl_lw_input);
  free(c->Tinput); free(c->precipinput); free(c->ocean_flux_input); free(c->styropor_input);
  { int q; for (q = 0; q < 8; q++) free(c->scr[q]); }
  free(c->bgc_abs[1]); free(c->bgc_abs[2]); free(c->fl_brine_bgc);
  free(c);
}

static double* dup1(const double* src, long n) {
  double* d = (double*)calloc((size_t)(n + 2), sizeof(double));
  if (src) memcpy(d + 1, src, sizeof(double) * (size_t)n);
  return d;
}

void sam_set_forcing(sam_col* c, int n, const double* fl_sw, const double* fl_lw, const double* T2m, const double* precip) {
  int k;
  free(c->time_input); free(c->T2m_input); free(c->precip_input); free(c->fl_sw_input); free(c->fl_lw_input);
  c->length_input = n;
  c->fl_sw_input = dup1(fl_sw, n);
  c->fl_lw_input = dup1(fl_lw, n);
  c->T2m_input = dup1(T2m, n);
  c->precip_input = dup1(precip, n);
  c->time_input = dup1(NULL, n);
  for (k = 1; k <= n; k++) c->time_input[k] = ((double)(float)k - 1.0) * 3600.0 * 3.0; /* mo_functions.f90:323-325 */
}

void sam_set_lab_forcing(sam_col* c, long n, const double* Tice, const double* snowfall, const double* heat,
                         const double* styropor) {
  long k;
  free(c->Tinput); free(c->precipinput); free(c->ocean_flux_input); free(c->styropor_input);
  c->length_input_lab = n;
  c->Tinput = dup1(Tice, n);
  c->precipinput = dup1(snowfall, n);
  c->ocean_flux_input = dup1(heat, n);
  c->styropor_input = dup1(styropor, n);
  if (c->snow_precip_flag == 0) { /* mo_grotz.f90:147-149 */
    for (k = 1; k <= n; k++) c->precipinput[k] = 0.0;
  }
}

/* ==========================================================================================
 * named access
 * ======================================================================================== */
typedef struct { const char* name; size_t off; int extra; } arr_desc;
#define AOFF(f) offsetof(sam_col, f)
static const arr_desc k_arrays[] = {
    {"H", AOFF(H), 0}, {"H_abs", AOFF(H_abs), 0}, {"fl_Q", AOFF(fl_Q), 1}, {"T", AOFF(T), 0}, {"S_bu", AOFF(S_bu), 0},
    {"S_abs", AOFF(S_abs), 0}, {"S_br", AOFF(S_br), 0}, {"thick", AOFF(thick), 0}, {"m", AOFF(m), 0},
    {"fl_m", AOFF(fl_m), 1}, {"V_ex", AOFF(V_ex), 0}, {"phi", AOFF(phi), 0}, {"psi_s", AOFF(psi_s), 0},
    {"psi_l", AOFF(psi_l), 0}, {"psi_g", AOFF(psi_g), 0}, {"ray", AOFF(ray), -1}, {"perm", AOFF(perm), 0},
    {"flush_v", AOFF(flush_v), 0}, {"flush_h", AOFF(flush_h), 0}, {"flush_v_old", AOFF(flush_v_old), 0},
    {"flush_h_old", AOFF(flush_h_old), 0}, {"fl_rad", AOFF(fl_rad), 0}, {"bgc_abs1", AOFF(bgc_abs[1]), 0},
    {"bgc_abs2", AOFF(bgc_abs[2]), 0}, {NULL, 0, 0}};

static const arr_desc* find_arr(const char* name) {
  const arr_desc* d;
  for (d = k_arrays; d->name; d++)
    if (strcmp(d->name, name) == 0) return d;
  return NULL;
}

int sam_array_len(const sam_col* c, const char* name) {
  const arr_desc* d = find_arr(name);
  if (!d) return -1;
  return c->Nlayer + d->extra;
}

int sam_get_array(const sam_col* c, const char* name, double* out) {
  const arr_desc* d = find_arr(name);
  real* a;
  int n, k;
  if (!d) return -1;
  a = *(real* const*)((const char*)c + d->off);
  n = c->Nlayer + d->extra;
  for (k = 0; k < n; k++) out[k] = (double)a[k + 1];
  return n;
}

int sam_set_array(sam_col* c, const char* name, const double* in) {
  const arr_desc* d = find_arr(name);
  real* a;
  int n, k;
  if (!d) return -1;
  a = *(real**)((char*)c + d->off);
  n = c->Nlayer + d->extra;
  for (k = 0; k < n; k++) a[k + 1] = in[k];
  return n;
}

typedef struct { const char* name; size_t off; } sc_desc;
static const sc_desc k_scalars[] = {
    {"dt", AOFF(dt)}, {"thick_0", AOFF(thick_0)}, {"time", AOFF(time)}, {"freeboard", AOFF(freeboard)},
    {"T_freeze", AOFF(T_freeze)}, {"time_out", AOFF(time_out)}, {"time_total", AOFF(time_total)},
    {"T_bottom", AOFF(T_bottom)}, {"T_top", AOFF(T_top)}, {"S_bu_bottom", AOFF(S_bu_bottom)}, {"T2m", AOFF(T2m)},
    {"fl_q_bottom", AOFF(fl_q_bottom)}, {"psi_s_snow", AOFF(psi_s_snow)}, {"psi_l_snow", AOFF(psi_l_snow)},
    {"psi_g_snow", AOFF(psi_g_snow)}, {"phi_s", AOFF(phi_s)}, {"S_abs_snow", AOFF(S_abs_snow)},
    {"H_abs_snow", AOFF(H_abs_snow)}, {"m_snow", AOFF(m_snow)}, {"T_snow", AOFF(T_snow)}, {"thick_snow", AOFF(thick_snow)},
    {"liquid_precip", AOFF(liquid_precip)}, {"solid_precip", AOFF(solid_precip)}, {"fl_q_snow", AOFF(fl_q_snow)},
    {"energy_stored", AOFF(energy_stored)}, {"total_resist", AOFF(total_resist)}, {"freshwater", AOFF(freshwater)},
    {"thickness", AOFF(thickness)}, {"bulk_salin", AOFF(bulk_salin)}, {"thick_min", AOFF(thick_min)},
    {"albedo", AOFF(albedo)}, {"fl_sw", AOFF(fl_sw)}, {"fl_lw", AOFF(fl_lw)}, {"fl_sen", AOFF(fl_sen)},
    {"fl_lat", AOFF(fl_lat)}, {"fl_rest", AOFF(fl_rest)}, {"grav_drain", AOFF(grav_drain)}, {"grav_salt", AOFF(grav_salt)},
    {"grav_temp", AOFF(grav_temp)}, {"melt_thick", AOFF(melt_thick)}, {"melt_thick_snow", AOFF(melt_thick_snow)},
    {"melt_thick_snow_old", AOFF(melt_thick_snow_old)}, {"melt_thick_output1", AOFF(melt_thick_output[1])},
    {"melt_thick_output2", AOFF(melt_thick_output[2])}, {"melt_thick_output3", AOFF(melt_thick_output[3])},
    {"alpha_flux_instable", AOFF(alpha_flux_instable)}, {"alpha_flux_stable", AOFF(alpha_flux_stable)},
    {"m_total", AOFF(m_total)}, {"S_total", AOFF(S_total)}, {"tank_depth", AOFF(tank_depth)}, {"melt_err", AOFF(melt_err)},
    {"max_flux_plate", AOFF(max_flux_plate)}, {"k_snow_flush", AOFF(k_snow_flush)}, {"k_styropor", AOFF(k_styropor)},
    {"ttop_warm", AOFF(ttop_warm)}, {"ttop_cold", AOFF(ttop_cold)}, {"oflux_amp", AOFF(oflux_amp)},
    {"bgc_bottom1", AOFF(bgc_bottom[1])}, {"bgc_bottom2", AOFF(bgc_bottom[2])}, {"bgc_total1", AOFF(bgc_total[1])},
    {"bgc_total2", AOFF(bgc_total[2])}, {NULL, 0}};

static const sc_desc k_ints[] = {
    {"testcase", AOFF(testcase)}, {"Nlayer", AOFF(Nlayer)}, {"N_top", AOFF(N_top)}, {"N_middle", AOFF(N_middle)},
    {"N_bottom", AOFF(N_bottom)}, {"N_active", AOFF(N_active)}, {"i", AOFF(i)}, {"i_time", AOFF(i_time)},
    {"i_time_out", AOFF(i_time_out)}, {"n_time_out", AOFF(n_time_out)}, {"time_counter", AOFF(time_counter)},
    {"length_input", AOFF(length_input)}, {"styropor_flag", AOFF(styropor_flag)}, {"atmoflux_flag", AOFF(atmoflux_flag)},
    {"grav_flag", AOFF(grav_flag)}, {"prescribe_flag", AOFF(prescribe_flag)}, {"grav_heat_flag", AOFF(grav_heat_flag)},
    {"flush_heat_flag", AOFF(flush_heat_flag)}, {"turb_flag", AOFF(turb_flag)}, {"salt_flag", AOFF(salt_flag)},
    {"boundflux_flag", AOFF(boundflux_flag)}, {"flush_flag", AOFF(flush_flag)}, {"flood_flag", AOFF(flood_flag)},
    {"bottom_flag", AOFF(bottom_flag)}, {"debug_flag", AOFF(debug_flag)}, {"precip_flag", AOFF(precip_flag)},
    {"harmonic_flag", AOFF(harmonic_flag)}, {"tank_flag", AOFF(tank_flag)}, {"albedo_flag", AOFF(albedo_flag)},
    {"lab_snow_flag", AOFF(lab_snow_flag)}, {"freeboard_snow_flag", AOFF(freeboard_snow_flag)},
    {"snow_flush_flag", AOFF(snow_flush_flag)}, {"snow_precip_flag", AOFF(snow_precip_flag)}, {"bgc_flag", AOFF(bgc_flag)},
    {"N_bgc", AOFF(N_bgc)}, {"status", AOFF(status)}, {NULL, 0}};

int sam_get_scalar(const sam_col* c, const char* name, double* out) {
  const sc_desc* d;
  for (d = k_scalars; d->name; d++)
    if (strcmp(d->name, name) == 0) {
      *out = (double)*(const real*)((const char*)c + d->off);
      return 0;
    }
  return -1;
}
int sam_set_scalar(sam_col* c, const char* name, double v) {
  const sc_desc* d;
  for (d = k_scalars; d->name; d++)
    if (strcmp(d->name, name) == 0) {
      *(real*)((char*)c + d->off) = v;
      return 0;
    }
  return -1;
}
int sam_get_int(const sam_col* c, const char* name, int* out) {
  const sc_desc* d;
  for (d = k_ints; d->name; d++)
    if (strcmp(d->name, name) == 0) {
      *out = *(const int*)((const char*)c + d->off);
      return 0;
    }
  return -1;
}
int sam_set_int(sam_col* c, const char* name, int v) {
  const sc_desc* d;
  for (d = k_ints; d->name; d++)
    if (strcmp(d->name, name) == 0) {
      *(int*)((char*)c + d->off) = v;
      return 0;
    }
  return -1;
}
static const char* const k_event_names[SAM_EV_COUNT] = {
    "flood", "flood_neg_free", "flood_simple", "flush3", "flush4", "flush_inline", "styropor", "snow_thermo",
    "snow_thermo_meltwater", "snow_wet", "snow_merge", "snow_compaction", "snow_coupling_iter", "snow_coupling_warm1",
    "snow_coupling_warm2", "snow_precip", "snow_precip_0", "melt_snow_all", "melt_snow_part", "bottom_melt",
    "bottom_melt_simple_a", "bottom_melt_simple_b", "bottom_growth_simple", "bottom_growth", "top_grow_a", "top_grow_b",
    "top_grow_c", "top_melt_a", "top_melt_b", "top_melt_c", "grav_drained", "salt_clamp",
    "gas_refill", "getT_Tfr_fallback", "getT_saltfree", "getT_liquid", "heat_melt", "heat_thin_snow", "melt_thick_gas",
    "snow_meltwater_to_ice", "prescribe", "grav_drain_simple", "notzflux", "flush3_clamp", "scrub", "melt_thick", "turb",
    "tank", "two_pass_step"};
const char* sam_event_name(int id) { return (id >= 0 && id < SAM_EV_COUNT) ? k_event_names[id] : NULL; }

long sam_get_stat(const sam_col* c, const char* name) {
  if (!strncmp(name, "ev_", 3)) {
    int q;
    for (q = 0; q < SAM_EV_COUNT; q++)
      if (!strcmp(name + 3, k_event_names[q])) return c->ev[q];
    return -1;
  }
  if (!strcmp(name, "getT_calls")) return c->stat_getT_calls;
  if (!strcmp(name, "newton_fr")) return c->stat_newton_fr;
  if (!strcmp(name, "newton_T")) return c->stat_newton_T;
  if (!strcmp(name, "layer_events")) return c->stat_layer_events;
  if (!strcmp(name, "flush_calls")) return c->stat_flush_calls;
  if (!strcmp(name, "flood_calls")) return c->stat_flood_calls;
  if (!strcmp(name, "coupling_iters")) return c->stat_coupling_iters;
  if (!strcmp(name, "n_outputs")) return c->n_outputs;
  return -1;
}

/* ==========================================================================================
 * helpers for the test harness (not in the reference)
 * ======================================================================================== */
void sam_set_output_hook(sam_col* c, sam_output_fn fn) {
  c->on_output = fn;
  c->on_output_user = NULL;
}

/* unit KAT drivers: evaluate one physics function over n inputs */
void sam_kat_getT(int salt_flag, int n, const double* H, const double* S_bu, const double* T_in, double* T_out,
                  double* phi_out) {
  sam_col c;
  int q;
  memset(&c, 0, sizeof c);
  c.salt_flag = salt_flag;
  for (q = 0; q < n; q++) {
    real T = 0.0, phi = 0.0;
    if (setjmp(c.jb) != 0) {
      T_out[q] = NAN;
      phi_out[q] = NAN;
      c.status = 0;
      continue;
    }
    sam_getT(&c, H[q], S_bu[q], T_in[q], &T, &phi, 0);
    T_out[q] = (double)T;
    phi_out[q] = (double)phi;
  }
}

/* fn: 0 S_br(a) 1 S_br(a,b) 2 ddT_S_br(a) 3 density(a,b) 4 T_freeze(a) 5 k_snow(a,b) 6 albedo(a=thick_snow,b=T_snow;
 * psi_l = 0.1, thick_min = 0.005, albedo_flag 2) */
void sam_kat_scalar(int fn, int salt_flag, int n, const double* a, const double* b, double* out) {
  sam_col c;
  int q;
  memset(&c, 0, sizeof c);
  c.salt_flag = salt_flag;
  for (q = 0; q < n; q++) {
    switch (fn) {
      case 0: out[q] = (double)sam_func_S_br(&c, a[q]); break;
      case 1: out[q] = (double)sam_func_S_br2(&c, a[q], b[q]); break;
      case 2: out[q] = (double)sam_func_ddT_S_br(&c, a[q]); break;
      case 3: out[q] = (double)sam_func_density(a[q], b[q]); break;
      case 4: out[q] = (double)sam_func_T_freeze(a[q], salt_flag); break;
      case 5: out[q] = (double)sam_func_k_snow(a[q], b[q]); break;
      case 6: out[q] = (double)sam_func_albedo(a[q], b[q], 0.1, 0.005, 2); break;
      default: out[q] = NAN;
    }
  }
}

/* CPU baseline driver: one column per OS thread */
#include <pthread.h>
typedef struct { sam_col** cols; int n, tid, nthreads; long nsteps; int rc; } batch_arg;
static void* batch_worker(void* p) {
  batch_arg* a = (batch_arg*)p;
  int q;
  for (q = a->tid; q < a->n; q += a->nthreads) {
    int rc = sam_step(a->cols[q], a->nsteps);
    if (rc != 0) a->rc = rc;
  }
  return NULL;
}
int sam_run_batch(sam_col** cols, int n, long nsteps, int nthreads) {
  pthread_t th[256];
  batch_arg args[256];
  int t, rc = 0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  for (t = 0; t < nthreads; t++) {
    args[t].cols = cols; args[t].n = n; args[t].tid = t; args[t].nthreads = nthreads; args[t].nsteps = nsteps; args[t].rc = 0;
    pthread_create(&th[t], NULL, batch_worker, &args[t]);
  }
  for (t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (args[t].rc) rc = args[t].rc;
  }
  return rc;
}
