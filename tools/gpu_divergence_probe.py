"""What do mixed warps cost?  The year-weighted ensembles of bench.py (per-column forcing, one model day of drift),
timed as they are, after a re-binning, and as a uniform ensemble (every column with the forcing of site 0).
Usage: python tools/gpu_divergence_probe.py [records ...]"""
import sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from samsim_b200 import api

recs = [int(a) for a in sys.argv[1:]] or [100, 330, 200]
sites = bench.load_sites(64)
sites = np.concatenate([sites[:, :, :2920], sites[:, :, :2920]], axis=2)
ncol = bench.YEAR_COLUMNS


def timed(eng):
    eng.step(320)
    eng.step(640, sync=False)
    eng.synchronize()
    return ncol * 640 / (eng.last_step_ms() * 1e-3) / 1e6


for rec in recs:
    st = bench.load_state(rec)
    row = {"record": rec}
    eng = bench.make_engine(api, st, sites, ncol, 0, 0)
    for _ in range(8):
        eng.step(1080)
    row["mixed"] = timed(eng)
    row["divergence"] = eng.divergence()
    eng.rebin()
    eng.clear_events()
    row["rebinned"] = timed(eng)
    # which branches ran in some columns only during the last 960 steps
    e0 = eng.get_int("events0").astype(np.uint32); e1 = eng.get_int("events1").astype(np.uint32)
    share = {}
    for i, name in enumerate(api.EVENT_NAMES):
        bits = (e0 >> np.uint32(i)) & 1 if i < 32 else (e1 >> np.uint32(i - 32)) & 1
        f = float(bits.mean())
        if 0.0 < f < 1.0: share[name] = round(f, 4)
    row["event_share"] = share
    eng.close()
    eng = api.Engine(api.Config.from_state(st), ncol, 0)
    eng.load_column_state(st, 0)
    eng.broadcast_column(0, 0, ncol)
    eng.set_forcing(sites[:1])
    for _ in range(8):
        eng.step(1080)
    row["uniform"] = timed(eng)
    eng.close()
    print(json.dumps(row), flush=True)
