"""Effect of samsim_b200_rebin on a drifted ensemble (not the official bench).

131072 columns start from the oracle's SHEBA freeze-up state (record 60) with a wide per-column T2m offset and the
bench's other perturbations, run `spin` steps so that freeze-up dates, N_active and snow states drift apart, then
the same 300 steps are timed on two identical handles: columns in their original order, and re-binned."""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from samsim_b200 import api

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
spin = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
rows = []
for rec in (60, 100):
    st = bench.load_state(rec)
    sites = bench.load_sites(64)
    cfg = api.Config.from_state(st)
    engs = []
    for rebin in (False, True):
        eng = api.Engine(cfg, ncol, 0)
        eng.load_column_state(st, 0)
        eng.broadcast_column(0, 0, ncol)
        site, scale, offset, amp = bench.perturbations(0, ncol)
        rng = np.random.default_rng(5)
        offset[2] = rng.uniform(-15, 5, ncol)
        eng.set_forcing(sites, site, scale, offset)
        eng.set_scalar("oflux_amp", amp)
        eng.step(spin)
        changed = eng.rebin() if rebin else False
        eng.step(300)
        eng.step(300)
        ms = eng.last_step_ms()
        na = eng.get_int("N_active")
        engs.append(eng)
        row = {"state": rec, "spin_steps": spin, "rebinned": rebin, "order_changed": bool(changed),
               "Mcolsteps_per_s": ncol * 300 / (ms * 1e-3) / 1e6, "N_active_min": int(na.min()),
               "N_active_mean": float(na.mean()), "N_active_max": int(na.max()), "failed": eng.count_failed()}
        rows.append(row)
        print(json.dumps(row), flush=True)
    same = all((engs[0].get_array(n) == engs[1].get_array(n)).all() for n in ("T", "S_abs", "H_abs", "thick"))
    print("identical results:", same, flush=True)
    rows[-1]["identical_to_unbinned"] = bool(same)
    for e in engs: e.close()
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(rows, open(ROOT / "gpurun_out" / "rebin_demo.json", "w"), indent=1)
