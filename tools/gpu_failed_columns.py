"""Explain the failed columns of the drifted freeze-up ensemble (42 failed columns in
profiles/r1j_rebin_demo_1M.json).  The ensemble of tools/gpu_rebin_demo.py is re-run (state 60, the bench's per-column
perturbations plus a T2m offset U(-15, 5) K, `spin` steps); every column whose status is non-zero is then re-run on the
CPU oracle with the same perturbed forcing: the oracle must STOP with the same code at the same step count or earlier
within the same launch.  Writes gpurun_out/failed_columns.json.   python tools/gpu_failed_columns.py [ncol] [spin]"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from samsim_b200 import api
from oracle import oracle

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
spin = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
st = bench.load_state(60)
sites = bench.load_sites(64)
cfg = api.Config.from_state(st)
eng = api.Engine(cfg, ncol, 0)
eng.load_column_state(st, 0)
eng.broadcast_column(0, 0, ncol)
site, scale, offset, amp = bench.perturbations(0, ncol)
offset[2] = np.random.default_rng(5).uniform(-15, 5, ncol)
eng.set_forcing(sites, site, scale, offset)
eng.set_scalar("oflux_amp", amp)
eng.step(spin)
status = eng.get_int("status")
bad = np.nonzero(status)[0]
rows = []
oracle.build()
for c in bad[:200]:
    col = oracle.Column(4, "det")
    col.set_forcing(*[sites[site[c], k] * scale[k, c] + offset[k, c] for k in range(4)])
    col.load_state(st)
    col.set_scalar("oflux_amp", float(amp[c]))
    rc = col.step(spin)
    rows.append({"column": int(c), "gpu_status": int(status[c]), "oracle_status": int(rc), "oracle_stopped_at_step": int(col.int("i")) - int(st["i"]),
                 "T2m_offset": float(offset[2, c]), "site": bench.SITES[site[c]]})
out = {"columns": ncol, "spin_steps": spin, "failed": int(len(bad)), "codes": {int(k): int(v) for k, v in zip(*np.unique(status[bad], return_counts=True))},
       "checked_on_oracle": len(rows), "same_code": int(sum(r["gpu_status"] == r["oracle_status"] for r in rows)), "rows": rows[:60]}
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(ROOT / "gpurun_out" / "failed_columns.json", "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "rows"}))
