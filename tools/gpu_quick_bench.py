"""Quick throughput probe (not the official bench): N columns from the oracle's SHEBA state 200."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from samsim_b200 import api
from oracle import oracle, parity_util as pu

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rec = int(sys.argv[3]) if len(sys.argv) > 3 else 200
pf = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # L1 prefetch distance of the layer sweeps (0 = default)
two_pass = int(sys.argv[5]) if len(sys.argv) > 5 else 0
z = np.load(ROOT / 'tests/golden/sheba_oracle_states.npz')
p = f'state{rec}_'
st = {k[len(p):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}
F = np.load(ROOT / 'tests/golden/forcing_era.npz')['sheba']
col = oracle.Column(4, 'det'); col.set_forcing(*F); col.load_state(st)
print('fp64 peak TF/s', api.fp64_peak(0, 1.0))
eng = pu.engine_from_oracle(col, ncol=ncol)
eng.set_forcing(F[None])
if pf or two_pass: eng.set_tuning(bool(two_pass), pf)
eng.step(20)
for rep in range(3):
    t0 = time.time(); eng.step(nsteps); dt = time.time() - t0
    ms = eng.last_step_ms()
    print(f'ncol {ncol} nsteps {nsteps}: wall {dt*1e3:.1f} ms, event {ms:.1f} ms -> {ncol*nsteps/(ms*1e-3)/1e6:.2f} M column-steps/s')
print('failed', eng.count_failed(), 'N_active', eng.reduce_diag()['N_active'])
