// params.cuh -- constant table of the column timestep (mo_parameters.f90:33-112).
//
// The reference declares several constants through default-REAL (single precision) literals;
// their double values are the widened float bit patterns, reproduced here with float casts
// (SURVEY.md section 8a-notes).  Everything is constexpr so the values are folded into the
// SASS as immediates.
#pragma once

namespace samsim {

#define SAMSIM_F32(x) ((double)(float)(x))

constexpr double pi_sp = (double)3.1415f;   // mo_parameters.f90:38  REAL, PARAMETER :: pi
constexpr double grav = (double)9.8061f;    // :39
constexpr double k_s = 2.2;                 // :46
constexpr double k_l = 0.523;               // :47
constexpr double c_s = 2020.0;              // :49
constexpr double c_s_beta = 7.6973;         // :50
constexpr double c_l = 3400.0;              // :51
constexpr double rho_s = 920.0;             // :52
constexpr double rho_l = 1028.0;            // :53
constexpr double latent_heat = 333500.0;    // :54
constexpr double zeroK = 273.15;            // :55
constexpr double bbeta = 0.8 * (double)1e-3f;      // :56  0.8_wp*1e-3
constexpr double mu = 2.55 * (double)1e-3f;        // :57
constexpr double kappa_l = k_l / rho_l / c_l;      // :58
constexpr double sigma = 5.6704 * (double)1e-8f;   // :59
constexpr double psi_s_min = 0.05;          // :69
constexpr double neg_free = -0.05;          // :70
constexpr double x_grav = 0.000584;         // :74
constexpr double ray_crit = 4.89;           // :75
constexpr double para_flush_horiz = 1.0;    // :79
constexpr double para_flush_gamma = 0.9;    // :81
constexpr double psi_s_top_min = 0.40;      // :83
constexpr double ratio_flood = 1.50;        // :85
constexpr double ref_salinity = 34.0;       // :87
constexpr double rho_snow = 330.0;          // :91
constexpr double gas_snow_ice2 = 0.20;      // :93
constexpr double emissivity_ice = 0.95;     // :96
constexpr double emissivity_snow = 1.00;    // :97
constexpr double penetr = 0.30;             // :98
constexpr double extinc = 2.00;             // :99
constexpr double Turb_A = 0.1 * 0.05 * rho_l / 86400.0;  // :102
constexpr double Turb_B = 0.05;             // :103

// liquidus polynomials, mo_thermo_functions.f90:321-336 (func_S_br) and :393-402 (func_ddT_S_br);
// index 0 = seawater (salt_flag 1), 1 = NaCl (salt_flag 2)
struct Liquidus {
  double c2, c3, c4;     // S_br = c2 T + c3 T^2 + c4 T^3
  double d2, d3x2, d4x3; // dS_br/dT = d2 + (2 d3) T + (3 d4) T^2   (old seawater coefficients!)
  double dcrit;          // derivative frozen below -20 C, :408-412
};

}  // namespace samsim
