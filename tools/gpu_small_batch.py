"""Config-4-like small batch (testcase 1 from init): python tools/gpu_small_batch.py [ncol] [nsteps] [bgc 0|1] [two_pass 0|1]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from samsim_b200 import api, grotz
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
bgc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
two_pass = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pf = int(sys.argv[5]) if len(sys.argv) > 5 else 0
st = grotz.init_testcase(1)
if not bgc:
    st["N_bgc"] = 0
cfg = api.Config.from_state(st)
if not bgc:
    cfg.N_bgc = 0
eng = api.Engine(cfg, ncol, 0)
eng.set_tuning(bool(two_pass), pf)
eng.load_column_state(st, 0)
eng.broadcast_column(0, 0, ncol)
eng.step(40000)      # grow some ice first
eng.synchronize()
t0 = time.time(); eng.step(nsteps); eng.synchronize(); dt = time.time() - t0
print(f"ncol {ncol} bgc {bgc} two_pass {two_pass} pf {pf} N_active {int(eng.get_int('N_active')[0])}: {ncol * nsteps / dt / 1e6:.2f} M column-steps/s ({dt / nsteps * 1e6:.1f} us/step)")
