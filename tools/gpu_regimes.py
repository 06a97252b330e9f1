"""Throughput of the step kernel in the different regimes of the SHEBA year (not the official bench).

For each oracle restart state: 131072 columns, bench-style per-column perturbations (9 sites, affine forcing,
oceanic flux amplitude), 300 warm-up steps so that the columns really diverge, then 300 timed steps."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from samsim_b200 import api

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
rows = []
for rec, what in [(200, "mid winter, grid full"), (100, "early winter, growing"), (60, "open water / freeze-up"),
                  (330, "melt onset, wet snow, flushing"), (345, "bare-ice melt, top_melt"), (300, "late winter")]:
    st = bench.load_state(rec)
    sites = bench.load_sites(64)
    cfg = api.Config.from_state(st)
    eng = api.Engine(cfg, ncol, 0)
    eng.load_column_state(st, 0)
    eng.broadcast_column(0, 0, ncol)
    site, scale, offset, amp = bench.perturbations(0, ncol)
    eng.set_forcing(sites, site, scale, offset)
    eng.set_scalar("oflux_amp", amp)
    eng.step(300)
    eng.step(300)
    ms = eng.last_step_ms()
    na = eng.get_int("N_active")
    snow = eng.get_scalar("thick_snow")
    row = {"state": rec, "regime": what, "Mcolsteps_per_s": ncol * 300 / (ms * 1e-3) / 1e6, "N_active_min": int(na.min()),
           "N_active_mean": float(na.mean()), "N_active_max": int(na.max()), "snow_frac": float((snow > 0).mean()),
           "failed": eng.count_failed()}
    rows.append(row)
    print(json.dumps(row), flush=True)
    eng.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(rows, open(ROOT / "gpurun_out" / "regimes.json", "w"), indent=1)
