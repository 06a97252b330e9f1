/* samsim_oracle.h -- CPU oracle for the SAMSIM column timestep.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a line-by-line scalar restatement of the
 * reference's time-loop body (mo_grotz.f90:182-835) and of the physics modules
 * it calls.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, link or call it; nothing under samsim_b200/
 * does.  Parity pin: validated against reference_output/Reference_testcase1_with_Version_2
 * and reference_output/Reference_SHEBA_with_Version_2 (see tests/test_oracle_golden.py
 * and tests/golden/).  The reference itself (Fortran) cannot be compiled in this
 * image (no gfortran/f951), see DESIGN.md.
 *
 * Layout: one `sam_col` holds what mo_data.f90:34-203 holds for one column.  Arrays
 * are 1-based like the Fortran (element 0 is unused) so that indices in this file
 * can be compared with the reference literally.
 */
#ifndef SAMSIM_ORACLE_H
#define SAMSIM_ORACLE_H

#include <setjmp.h>
#include <stdint.h>

#ifdef SAMSIM_COUNT_OPS
#include "count_real.h" /* C++ only: counting wrapper */
#else
typedef double real;
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sam_col sam_col;

/* called at S8 (mo_grotz.f90:340-398) right where the reference calls output() */
typedef void (*sam_output_fn)(sam_col* c, void* user);

struct sam_col {
  /* ---- grid / run control (mo_data.f90:57-74) ---- */
  int testcase;
  int Nlayer, N_top, N_middle, N_bottom, N_active;
  int i;          /* loop index of the step being executed, 1-based (mo_grotz.f90:182) */
  int i_time, i_time_out, n_time_out;
  int time_counter, length_input;
  int styropor_flag;
  /* ---- flags (mo_data.f90:136-155) ---- */
  int atmoflux_flag, grav_flag, prescribe_flag, grav_heat_flag, flush_heat_flag, turb_flag, salt_flag,
      boundflux_flag, flush_flag, flood_flag, bottom_flag, debug_flag, precip_flag, harmonic_flag, tank_flag,
      albedo_flag, lab_snow_flag, freeboard_snow_flag, snow_flush_flag, snow_precip_flag, bgc_flag;
  /* ---- scalars ---- */
  real dt, thick_0, time, freeboard, T_freeze, time_out, time_total;
  real T_bottom, T_top, S_bu_bottom, T2m, fl_q_bottom;
  real psi_s_snow, psi_l_snow, psi_g_snow, phi_s, S_abs_snow, H_abs_snow, m_snow, T_snow, thick_snow;
  real liquid_precip, solid_precip, fl_q_snow;
  real energy_stored, total_resist, freshwater, thickness, bulk_salin;
  real thick_min, T_test;
  real albedo, fl_sw, fl_lw, fl_sen, fl_lat, fl_rest;
  real grav_drain, grav_salt, grav_temp;
  real melt_thick, melt_thick_snow, melt_thick_snow_old;
  real melt_thick_output[4]; /* 1-based, 3 used */
  real alpha_flux_instable, alpha_flux_stable;
  real m_total, S_total, tank_depth;
  real melt_err;
  /* mutable "parameters" (mo_parameters.f90:107-112) */
  real max_flux_plate, k_snow_flush, k_styropor;
  /* ---- per-column knobs that are literals in the reference (identity values reproduce it) ----
   * ttop_warm/ttop_cold: the -5/-10 levels of sub_test1 (mo_testcase_specifics.f90:46-87);
   * oflux_amp: the 7 W/m2 amplitude of sub_test4 (:200). */
  real ttop_warm, ttop_cold, oflux_amp;
  /* ---- arrays, 1-based ---- */
  real *H, *H_abs, *fl_Q, *T, *S_bu, *S_abs, *S_br, *thick, *m, *fl_m, *V_ex, *phi, *psi_s, *psi_l, *psi_g, *ray,
      *perm, *flush_v, *flush_h, *flush_v_old, *flush_h_old, *fl_rad;
  real* scr[8]; /* scratch standing in for the callees' automatic arrays */
  /* ---- passive tracers, bgc_flag == 2 (mo_data.f90 bgc_*; mo_init.f90 sub_allocate_bgc) ---- */
  int N_bgc;                 /* 0..2 */
  real* bgc_abs[3];          /* [tracer 1..2][layer 1..Nlayer] */
  real bgc_bottom[3], bgc_total[3];
  real* fl_brine_bgc;        /* (Nlayer+1) x (Nlayer+1), FB(i,j), 1-based */
  /* ---- forcing (atmoflux_flag==2), 1-based, length_input records ---- */
  double *time_input, *T2m_input, *precip_input, *fl_sw_input, *fl_lw_input;
  /* ---- lab forcing (testcases 101-105), 1-based ---- */
  long length_input_lab;
  double *Tinput, *precipinput, *ocean_flux_input, *styropor_input;
  /* ---- error state: reference STOP codes ---- */
  int status;
  jmp_buf jb;
  /* ---- output hook ---- */
  sam_output_fn on_output;
  void* on_output_user;
  long n_outputs;
  /* ---- statistics (not in the reference) ---- */
  long stat_getT_calls, stat_newton_fr, stat_newton_T, stat_layer_events, stat_flush_calls, stat_flood_calls,
      stat_coupling_iters;
  /* per-branch execution counters (not in the reference): how often each rarely taken branch of the path ran,
   * so that a parity test can assert that the branch it claims to cover was executed.  Order = SAM_EV_* below =
   * samsim_event_id of include/samsim_b200.h (checked by tests/test_cabi_cpu.py). */
  long ev[64];
};

/* branch ids, each named after the reference lines it marks (see sam_event_name) */
enum {
  SAM_EV_FLOOD = 0, SAM_EV_FLOOD_NEG_FREE, SAM_EV_FLOOD_SIMPLE, SAM_EV_FLUSH3, SAM_EV_FLUSH4, SAM_EV_FLUSH_INLINE,
  SAM_EV_STYROPOR, SAM_EV_SNOW_THERMO, SAM_EV_SNOW_THERMO_MELTWATER, SAM_EV_SNOW_WET, SAM_EV_SNOW_MERGE,
  SAM_EV_SNOW_COMPACTION, SAM_EV_SNOW_COUPLING_ITER, SAM_EV_SNOW_COUPLING_WARM1, SAM_EV_SNOW_COUPLING_WARM2,
  SAM_EV_SNOW_PRECIP, SAM_EV_SNOW_PRECIP_0, SAM_EV_MELT_SNOW_ALL, SAM_EV_MELT_SNOW_PART, SAM_EV_BOTTOM_MELT,
  SAM_EV_BOTTOM_MELT_SIMPLE_A, SAM_EV_BOTTOM_MELT_SIMPLE_B, SAM_EV_BOTTOM_GROWTH_SIMPLE, SAM_EV_BOTTOM_GROWTH,
  SAM_EV_TOP_GROW_A, SAM_EV_TOP_GROW_B, SAM_EV_TOP_GROW_C, SAM_EV_TOP_MELT_A, SAM_EV_TOP_MELT_B, SAM_EV_TOP_MELT_C,
  SAM_EV_GRAV_DRAINED, SAM_EV_SALT_CLAMP,
  SAM_EV_GAS_REFILL, SAM_EV_GETT_TFR_FALLBACK, SAM_EV_GETT_SALTFREE, SAM_EV_GETT_LIQUID, SAM_EV_HEAT_MELT,
  SAM_EV_HEAT_THIN_SNOW, SAM_EV_MELT_THICK_GAS, SAM_EV_SNOW_MELTWATER_TO_ICE, SAM_EV_PRESCRIBE,
  SAM_EV_GRAV_DRAIN_SIMPLE, SAM_EV_NOTZFLUX, SAM_EV_FLUSH3_CLAMP, SAM_EV_SCRUB, SAM_EV_MELT_THICK, SAM_EV_TURB,
  SAM_EV_TANK, SAM_EV_TWO_PASS_STEP /* device only: never counted here */,
  SAM_EV_COUNT
};
/* "flood", "flood_neg_free", ... in the order above; NULL beyond SAM_EV_COUNT */
const char* sam_event_name(int id);

/* Allocate a column and run the reference's init(testcase) for testcases 1, 4, 101-105
 * (mo_init.f90:83-132, :865-945, :1127-1207, :222-767, :1982-2009).  Returns NULL for
 * other testcases (reference: STOP 4321 only for unknown ones; the others are out of scope). */
sam_col* sam_create(int testcase);
void sam_destroy(sam_col* c);

/* atmoflux_flag==2 forcing, as sub_input builds it (mo_functions.f90:304-327): n records,
 * time_input(k) = (k-1)*10800.  Arrays are copied. */
void sam_set_forcing(sam_col* c, int n, const double* fl_sw, const double* fl_lw, const double* T2m,
                     const double* precip);
/* lab forcing (mo_grotz.f90:138-169), n per-dt records each. Arrays are copied. */
void sam_set_lab_forcing(sam_col* c, long n, const double* Tice, const double* snowfall, const double* heat,
                         const double* styropor);

/* Advance nsteps iterations of the loop body.  Returns 0 or the reference STOP code. */
int sam_step(sam_col* c, long nsteps);

/* named access for the Python tests: arrays return the 0-based contents (Fortran element 1 first) */
int sam_array_len(const sam_col* c, const char* name);
int sam_get_array(const sam_col* c, const char* name, double* out);
int sam_set_array(sam_col* c, const char* name, const double* in);
int sam_get_scalar(const sam_col* c, const char* name, double* out);
int sam_set_scalar(sam_col* c, const char* name, double v);
int sam_get_int(const sam_col* c, const char* name, int* out);
int sam_set_int(sam_col* c, const char* name, int v);
long sam_get_stat(const sam_col* c, const char* name);

/* individual physics routines, exported for unit KATs against the CUDA device functions */
void sam_getT(sam_col* c, real H, real S_bu, real T_in, real* T, real* phi, int k);
real sam_func_S_br(const sam_col* c, real T);
real sam_func_S_br2(const sam_col* c, real T, real S_bu);
real sam_func_ddT_S_br(const sam_col* c, real T);
real sam_func_density(real T, real S);
real sam_func_T_freeze(real S_bu, int salt_flag);
real sam_func_albedo(real thick_snow, real T_snow, real psi_l, real thick_min, int albedo_flag);
real sam_func_freeboard(const sam_col* c);
real sam_func_k_snow(real m_snow, real thick_snow);

/* harness helpers */
void sam_set_output_hook(sam_col* c, sam_output_fn fn);
void sam_kat_getT(int salt_flag, int n, const double* H, const double* S_bu, const double* T_in, double* T_out,
                  double* phi_out);
void sam_kat_scalar(int fn, int salt_flag, int n, const double* a, const double* b, double* out);
int sam_run_batch(sam_col** cols, int n, long nsteps, int nthreads);

/* the math backend this build uses: "libm" or "det" */
const char* sam_math_backend(void);
double sam_math_pow(double x, double y);
double sam_math_exp(double x);
double sam_math_sin(double x);

#ifdef __cplusplus
}
#endif
#endif
