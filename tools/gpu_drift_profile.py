"""Drifted, re-binned ensemble for profiling (see gpu_rebin_demo.py): spin, rebin, then two 300-step launches."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from samsim_b200 import api
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
spin = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
nst = int(sys.argv[3]) if len(sys.argv) > 3 else 100
st = bench.load_state(100)
sites = bench.load_sites(64)
eng = api.Engine(api.Config.from_state(st), ncol, 0)
eng.load_column_state(st, 0)
eng.broadcast_column(0, 0, ncol)
site, scale, offset, amp = bench.perturbations(0, ncol)
offset[2] = np.random.default_rng(5).uniform(-15, 5, ncol)
eng.set_forcing(sites, site, scale, offset)
eng.set_scalar("oflux_amp", amp)
eng.step(spin)
eng.rebin()
print("launches before the timed ones:", eng.launch_count())
eng.step(nst)
eng.step(nst)
print("Mcolsteps/s", ncol * nst / (eng.last_step_ms() * 1e-3) / 1e6)
