// kernel_on_host.cpp -- TEST HARNESS, not part of the product.
//
// Compiles the CUDA device functions of samsim_b200/csrc/{physics,step}.cuh for the HOST (g++ -ffp-contract=off) so
// that the no-GPU test stage can diff the kernel's per-column logic -- sub-step order, fusions, memoised sums, lazy
// Rayleigh numbers, tracer replay -- bit for bit against the CPU oracle (tests/test_kernel_on_host_cpu.py).
// One column, the same `Col` / `Lay` views the kernel builds, ncol_pad = 1.
//
// It is built only by that test, into tests/hostbuild/_build/; it exports no samsim_b200_* symbol, nothing under
// samsim_b200/ refers to it, and the product still has no CPU path (samsim_b200_create fails without a GPU).
#include <stdint.h>
#include <string.h>

#include <vector>

#define SAMSIM_HOST_BUILD 1
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define SAMSIM_SYNC 0  // phase barriers are a scheduling device of the GPU build only
#define SAMSIM_LOOP    // "#pragma unroll 1" likewise

#include "../../include/samsim_b200.h"
#include "../../samsim_b200/csrc/step.cuh"

using namespace samsim;

namespace {

struct HostKernel {
  samsim_config_t cfg;
  DevCfg g;
  int LS;
  std::vector<double> arr;  // one tile, lane 0: element (slot a, layer k) at [(k*AR_COUNT + a)*SAMSIM_TILE]
  double sc[SC_COUNT];
  int N_active, status, styropor_flag;
  unsigned ev0 = 0, ev1 = 0;
  double time;
  long long i;
  int n_time_out, time_counter;
  std::vector<double> series;  // [4][nrec]
  int nrec;
  double fscale[4], foffset[4];
  std::vector<double> lab;     // [4][lab_nrec]
  long long lab_nrec;
  std::vector<double> snap_sc, snap_arr;
};

}  // namespace

extern "C" {

// same cfg -> DevCfg mapping as samsim_b200_create (samsim_b200.cu); the parity tests would expose a divergence
void* hostk_create(const samsim_config_t* cfg) {
  HostKernel* h = new HostKernel();
  h->cfg = *cfg;
  DevCfg& d = h->g;
  memset(&d, 0, sizeof d);
  d.testcase = cfg->testcase; d.Nlayer = cfg->Nlayer; d.N_top = cfg->N_top; d.N_middle = cfg->N_middle; d.N_bottom = cfg->N_bottom;
  d.atmoflux_flag = cfg->atmoflux_flag; d.grav_flag = cfg->grav_flag; d.prescribe_flag = cfg->prescribe_flag;
  d.grav_heat_flag = cfg->grav_heat_flag; d.flush_heat_flag = cfg->flush_heat_flag; d.turb_flag = cfg->turb_flag;
  d.salt_flag = cfg->salt_flag; d.boundflux_flag = cfg->boundflux_flag; d.flush_flag = cfg->flush_flag;
  d.flood_flag = cfg->flood_flag; d.bottom_flag = cfg->bottom_flag; d.precip_flag = cfg->precip_flag;
  d.harmonic_flag = cfg->harmonic_flag; d.tank_flag = cfg->tank_flag; d.albedo_flag = cfg->albedo_flag;
  d.lab_snow_flag = cfg->lab_snow_flag; d.freeboard_snow_flag = cfg->freeboard_snow_flag;
  d.snow_flush_flag = cfg->snow_flush_flag; d.snow_precip_flag = cfg->snow_precip_flag;
  d.i_time_out = cfg->i_time_out;
  d.n_bgc = cfg->N_bgc;
  d.dt = cfg->dt; d.thick_0 = cfg->thick_0; d.thick_min = cfg->thick_min; d.time_out = cfg->time_out;
  d.alpha_flux_instable = cfg->alpha_flux_instable; d.alpha_flux_stable = cfg->alpha_flux_stable; d.m_total = cfg->m_total;
  d.max_flux_plate = cfg->max_flux_plate; d.k_snow_flush = cfg->k_snow_flush; d.k_styropor = cfg->k_styropor;
  d.pf = 2;
  if (cfg->salt_flag == 1) {
    d.c2 = -18.7; d.c3 = -0.519; d.c4 = -0.00535; d.d2 = -21.4; d.d3x2 = 2.0 * -0.886; d.d4x3 = 3.0 * -0.0170;
  } else {
    d.c2 = -17.6; d.c3 = -0.389; d.c4 = -0.00362; d.d2 = -17.6; d.d3x2 = 2.0 * -0.389; d.d4x3 = 3.0 * -0.00362;
  }
  h->LS = cfg->Nlayer + 2;
  h->arr.assign((size_t)AR_COUNT * h->LS * SAMSIM_TILE, 0.0);
  memset(h->sc, 0, sizeof h->sc);
  h->N_active = 1; h->status = 0; h->styropor_flag = 0;
  h->time = 0.0; h->i = 0; h->n_time_out = 0; h->time_counter = 1;
  h->nrec = 0; h->lab_nrec = 0;
  for (int k = 0; k < 4; k++) { h->fscale[k] = 1.0; h->foffset[k] = 0.0; }
  h->snap_sc.assign(SAMSIM_SNAPSC_COUNT, 0.0);
  h->snap_arr.assign((size_t)SAMSIM_SNAPARR_COUNT * h->LS, 0.0);
  return h;
}

void hostk_destroy(void* p) { delete (HostKernel*)p; }
void hostk_set_tuning(void* p, int two_pass) { ((HostKernel*)p)->g.two_pass = two_pass; }

static int slot_of_public(int id) { return (id < AR_STATE_COUNT) ? id : AR_BGC1 + (id - AR_STATE_COUNT); }

// layer k (1-based) of public array `id` lives at arr[(k*AR_COUNT + slot)*SAMSIM_TILE] (lane 0 of the one tile)
void hostk_set_array(void* p, int id, const double* v, int n) {
  HostKernel* h = (HostKernel*)p;
  for (int k = 0; k < n; k++) h->arr[((size_t)(k + 1) * AR_COUNT + slot_of_public(id)) * SAMSIM_TILE] = v[k];
}
void hostk_get_array(void* p, int id, double* v, int n) {
  HostKernel* h = (HostKernel*)p;
  for (int k = 0; k < n; k++) v[k] = h->arr[((size_t)(k + 1) * AR_COUNT + slot_of_public(id)) * SAMSIM_TILE];
}
double* hostk_scalars(void* p) { return ((HostKernel*)p)->sc; }
void hostk_set_ints(void* p, int N_active, int status, int styropor_flag) {
  HostKernel* h = (HostKernel*)p;
  h->N_active = N_active; h->status = status; h->styropor_flag = styropor_flag;
}
void hostk_get_ints(void* p, int* out5) {
  HostKernel* h = (HostKernel*)p;
  out5[0] = h->N_active; out5[1] = h->status; out5[2] = h->styropor_flag; out5[3] = (int)h->ev0; out5[4] = (int)h->ev1;
}
void hostk_set_clock(void* p, double time, long long i, int n_time_out, int time_counter) {
  HostKernel* h = (HostKernel*)p;
  h->time = time; h->i = i; h->n_time_out = n_time_out; h->time_counter = time_counter;
}
void hostk_get_clock(void* p, double* time, long long* i, int* n_time_out, int* time_counter) {
  HostKernel* h = (HostKernel*)p;
  *time = h->time; *i = h->i; *n_time_out = h->n_time_out; *time_counter = h->time_counter;
}
void hostk_set_forcing(void* p, int nrec, const double* series /* [4][nrec] */, const double* scale4, const double* offset4) {
  HostKernel* h = (HostKernel*)p;
  h->nrec = nrec;
  h->series.assign(series, series + (size_t)4 * nrec);
  for (int k = 0; k < 4; k++) { h->fscale[k] = scale4 ? scale4[k] : 1.0; h->foffset[k] = offset4 ? offset4[k] : 0.0; }
}
void hostk_set_lab_forcing(void* p, long long nrec, const double* series /* [4][nrec] */) {
  HostKernel* h = (HostKernel*)p;
  h->lab_nrec = nrec;
  h->lab.assign(series, series + (size_t)4 * nrec);
  if (h->cfg.snow_precip_flag == 0) std::fill(h->lab.begin() + nrec, h->lab.begin() + 2 * nrec, 0.0);  // mo_grotz.f90:147-149
}
const double* hostk_snapshot_scalars(void* p) { return ((HostKernel*)p)->snap_sc.data(); }
// row `id` of the S8 record, layers 1..Nlayer
void hostk_get_snapshot_array(void* p, int id, double* v, int n) {
  HostKernel* h = (HostKernel*)p;
  for (int k = 0; k < n; k++) v[k] = h->snap_arr[(size_t)id * h->LS + k + 1];
}

// getT of the device code, elementwise (known-answer test of the lazy freezing point); st = STOP code, ev = word-1 event bits
void hostk_kat_getT(int salt_flag, int n, const double* H, const double* S_bu, const double* T_in, double* T_out,
                    double* phi_out, int* st, int* ev) {
  DevCfg g;
  memset(&g, 0, sizeof g);
  g.salt_flag = salt_flag;
  if (salt_flag == 1) { g.c2 = -18.7; g.c3 = -0.519; g.c4 = -0.00535; g.d2 = -21.4; g.d3x2 = 2.0 * -0.886; g.d4x3 = 3.0 * -0.0170; }
  else { g.c2 = -17.6; g.c3 = -0.389; g.c4 = -0.00362; g.d2 = -17.6; g.d3x2 = 2.0 * -0.389; g.d4x3 = 3.0 * -0.00362; }
  for (int q = 0; q < n; q++) {
    double T = 0.0, phi = 0.0;
    int status = 0;
    unsigned ev1 = 0;
    samsim_host_cfg = &g;
    getT(H[q], S_bu[q], T_in[q], T, phi, status, ev1);
    T_out[q] = T; phi_out[q] = phi; st[q] = status; ev[q] = (int)ev1;
  }
}

// one "launch" of nsteps steps: what samsim_step_kernel does for one thread (samsim_b200.cu), whole series as window
int hostk_step(void* p, long long nsteps) {
  HostKernel* h = (HostKernel*)p;
  Col c;
  c.base = h->arr.data();
  c.ls = (unsigned)(AR_COUNT * SAMSIM_TILE);
  for (int q = 0; q < SC_COUNT; q++) c.sc[q] = h->sc[q];
  c.N_active = h->N_active; c.status = h->status; c.styropor_flag = h->styropor_flag;
  c.ev0 = h->ev0; c.ev1 = h->ev1;
  c.time = h->time; c.i = h->i; c.n_time_out = h->n_time_out; c.time_counter = h->time_counter;
  c.fsw0 = c.fsw1 = c.flw0 = c.flw1 = c.ftime0 = c.ftime1 = 0.0;
  c.thermo_valid = false;
  c.pre.valid = false;
  c.want_state = false;
  c.fb.tot_valid = c.fb.suf_valid = c.fb.res_valid = false; c.fb.k_last = 0; c.fb.ks = 0;
  c.min_psi_s = 0.0; c.min_S_abs_2 = 0.0;
  c.fb_x = 0.0;

  Forcing f;
  f.win = h->series.empty() ? nullptr : h->series.data();
  f.win_len = h->nrec; f.win_first = 1; f.site = 0;
  for (int k = 0; k < 4; k++) { f.scale[k] = h->fscale[k]; f.offset[k] = h->foffset[k]; }
  f.lab = h->lab.empty() ? nullptr : h->lab.data();
  f.lab_nrec = h->lab_nrec; f.lab_set = 0;

  samsim_host_cfg = &h->g;
  SnapOut snap;
  snap.scalars = h->snap_sc.data(); snap.arrays = h->snap_arr.data(); snap.ncol_pad = 1; snap.col = 0;

  for (long long s = 0; s < nsteps; s++) {
    if (h->g.two_pass) column_step<true>(c, f, s == nsteps - 1, snap);
    else column_step<false>(c, f, s == nsteps - 1, snap);
  }

  for (int q = 0; q < SC_COUNT; q++) h->sc[q] = c.sc[q];
  h->N_active = c.N_active; h->status = c.status; h->styropor_flag = c.styropor_flag;
  h->ev0 = c.ev0; h->ev1 = c.ev1;
  h->time = c.time; h->i = c.i; h->n_time_out = c.n_time_out; h->time_counter = c.time_counter;
  return c.status;
}

}  // extern "C"
