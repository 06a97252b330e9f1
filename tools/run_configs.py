"""Full-size runs of BASELINE.json configs 4 and 5 on one B200, with oracle spot checks at scale.

config 4: 10,000-member perturbed ensemble of testcase 1, all 259,200 steps.
config 5: 1,048,576 ERA-style columns, one model day (8,641 steps incl. one S8 output) from the mid-January state.
For each, a few columns are re-run on the CPU oracle and compared bit for bit (state after the run).
Writes gpurun_out/configs.json.
"""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from samsim_b200 import api, grotz
from oracle import oracle, parity_util as pu

out = {}

# ---------------- config 4 ----------------
ncol = 10000
rng = np.random.default_rng(20170301)
d1, d2 = rng.normal(0, 0.5, ncol), rng.normal(0, 0.5, ncol)
Sb = 34.0 + rng.normal(0, 0.5, ncol)
Tb = np.maximum(-1.0 + rng.normal(0, 0.05, ncol), -1.7)
fq = np.abs(rng.normal(0, 2.0, ncol))
st = grotz.init_testcase(1)
cfg = api.Config.from_state(st)
eng = api.Engine(cfg, ncol, 0)
eng.load_column_state(st, 0)
eng.broadcast_column(0, 0, ncol)
c_l = 3400.0
m1 = st["m"][0]
eng.set_scalar("ttop_warm", -5.0 + d1); eng.set_scalar("ttop_cold", -10.0 + d2); eng.set_scalar("T_top", -5.0 + d1)
eng.set_scalar("T_bottom", Tb); eng.set_scalar("S_bu_bottom", Sb); eng.set_scalar("fl_q_bottom", fq)
S = np.tile(st["S_abs"], (ncol, 1)); S[:, 0] = Sb * m1; eng.set_array("S_abs", S)
H = np.tile(st["H_abs"], (ncol, 1)); H[:, 0] = m1 * Tb * c_l; eng.set_array("H_abs", H)
eng.set_array("T", np.repeat(Tb[:, None], 90, 1)); eng.set_array("S_bu", np.repeat(Sb[:, None], 90, 1))
nsteps = st["i_time"]
t0 = time.time(); eng.step(nsteps); wall = time.time() - t0
na = eng.get_int("N_active"); th = eng.get_scalar("thickness")
bad_total = 0
for j in (0, 4999, 9999):
    col = oracle.Column(1, "det")
    col.set_scalar("ttop_warm", -5.0 + d1[j]); col.set_scalar("ttop_cold", -10.0 + d2[j]); col.set_scalar("T_top", -5.0 + d1[j])
    col.set_scalar("T_bottom", Tb[j]); col.set_scalar("S_bu_bottom", Sb[j]); col.set_scalar("fl_q_bottom", fq[j])
    a = col.array("S_abs"); a[0] = Sb[j] * m1; col.set_array("S_abs", a)
    a = col.array("H_abs"); a[0] = m1 * Tb[j] * c_l; col.set_array("H_abs", a)
    a = col.array("T"); a[:] = Tb[j]; col.set_array("T", a)
    a = col.array("S_bu"); a[:] = Sb[j]; col.set_array("S_bu", a)
    assert col.step(nsteps) == 0
    bad = pu.compare_column(col, eng, j)
    bad_total += len(bad)
    if bad: print("config 4 member", j, bad[:3])
out["config4"] = {"columns": ncol, "steps": nsteps, "column_steps": ncol * nsteps, "wall_s": wall,
                  "Mcolsteps_per_s": ncol * nsteps / wall / 1e6, "failed": eng.count_failed(),
                  "N_active_min_mean_max": [int(na.min()), float(na.mean()), int(na.max())],
                  "thickness_m_min_mean_max": [float(th.min()), float(th.mean()), float(th.max())],
                  "oracle_spot_check_members": [0, 4999, 9999], "oracle_bitwise_mismatches": bad_total}
print(json.dumps(out["config4"]), flush=True)
eng.close()

# ---------------- config 5 ----------------
ncol = bench.TOTAL_COLUMNS
st = bench.load_state(bench.START_RECORD)
sites = bench.load_sites(64)
cfg = api.Config.from_state(st)
eng = api.Engine(cfg, ncol, 0)
eng.load_column_state(st, 0)
eng.broadcast_column(0, 0, ncol)
site, scale, offset, amp = bench.perturbations(0, ncol)
eng.set_forcing(sites, site, scale, offset)
eng.set_scalar("oflux_amp", amp)
eng.set_snapshot_mode(api.SNAP_SCALARS_ONLY)
nsteps = 8641
t0 = time.time(); eng.step(nsteps); wall = time.time() - t0
ms = eng.last_step_ms()
na = eng.get_int("N_active"); th = eng.get_scalar("thickness")
bad_total = 0
check = [0, 1, 524287, 1048575]
cols = bench.make_oracle_columns(st, sites, np.array(check))
for j, col in zip(check, cols):
    assert col.step(nsteps) == 0
    bad = pu.compare_column(col, eng, j)
    bad_total += len(bad)
    if bad: print("config 5 column", j, bad[:3])
out["config5_one_day"] = {"columns": ncol, "steps": nsteps, "column_steps": ncol * nsteps, "wall_s": wall, "kernel_ms": ms,
                          "Mcolsteps_per_s": ncol * nsteps / (ms * 1e-3) / 1e6, "failed": eng.count_failed(),
                          "launches": eng.launch_count(),
                          "N_active_min_mean_max": [int(na.min()), float(na.mean()), int(na.max())],
                          "thickness_m_min_mean_max": [float(th.min()), float(th.mean()), float(th.max())],
                          "oracle_spot_check_columns": check, "oracle_bitwise_mismatches": bad_total}
print(json.dumps(out["config5_one_day"]), flush=True)
eng.close()

# ---------------- config 5, one model month of one wave of columns (31 S8 records), divergence-driven re-binning ----------------
ncol = bench.YEAR_COLUMNS
eng = bench.make_engine(api, st, sites, ncol, 0, 0)
eng.set_snapshot_mode(api.SNAP_SCALARS_ONLY)
eng.set_rebin_auto(bench.YEAR_REBIN_THRESHOLD)
ndays = 31
t0 = time.time()
kernel_ms = 0.0
for d in range(ndays):
    for _ in range(8):
        eng.step(1080, sync=False)
        eng.synchronize()
        kernel_ms += eng.last_step_ms()
    eng.step(1, sync=False)   # 8 x 1080 + 1 = 8641 = one output period
    eng.synchronize()
    kernel_ms += eng.last_step_ms()
wall = time.time() - t0
nsteps = ndays * 8641
na = eng.get_int("N_active"); th = eng.get_scalar("thickness")
check = [0, 1, ncol // 2, ncol - 1]
cols = bench.make_oracle_columns(st, sites, np.array(check))
bad_total = 0
for j, col in zip(check, cols):
    assert col.step(nsteps) == 0
    bad = pu.compare_column(col, eng, j)
    bad_total += len(bad)
    if bad: print("config 5 month column", j, bad[:3])
out["config5_one_month"] = {"columns": ncol, "steps": nsteps, "column_steps": ncol * nsteps, "wall_s": wall, "kernel_ms": kernel_ms,
                            "Mcolsteps_per_s": ncol * nsteps / (kernel_ms * 1e-3) / 1e6, "failed": eng.count_failed(),
                            "divergence": eng.divergence(),
                            "N_active_min_mean_max": [int(na.min()), float(na.mean()), int(na.max())],
                            "thickness_m_min_mean_max": [float(th.min()), float(th.mean()), float(th.max())],
                            "oracle_spot_check_columns": check, "oracle_bitwise_mismatches": bad_total}
print(json.dumps(out["config5_one_month"]), flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "configs.json", "w"), indent=1)
