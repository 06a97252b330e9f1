"""Run the oracle on testcase 4 (SHEBA forcing) and dump output records + restartable states.

Usage: python tools/run_oracle_sheba.py <backend: libm|det> <outfile.npz> [nrecords]
Reads forcing from /root/reference (this container only); the result feeds tools/make_golden.py.
"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import oracle

backend = sys.argv[1]
out = sys.argv[2]
nrec = int(sys.argv[3]) if len(sys.argv) > 3 else 1643
c = oracle.Column(4, backend)
c.set_forcing(*oracle.read_forcing_dir('/root/reference'))
c.record_outputs()
period = c.int('i_time_out') + 1
save_at = {1, 2, 5, 10, 20, 40, 60, 80, 100, 150, 200, 250, 300, 330, 345, 360, 380, 400, 450, 500, 600, 700, 715, 730, 800,
           1000, 1200, 1400, 1600}
states = {}
t0 = time.time()
done = 0
rc = 0
for j in range(1, nrec + 1):
    target = (j - 1) * period  # steps completed before the step that writes record j
    n = target - done
    if n > 0:
        rc = c.step(n)
        done = target
    if rc:
        print('STOP', rc, 'before record', j); break
    if j in save_at:
        states[j] = c.state()
    rc = c.step(1); done += 1
    if rc:
        print('STOP', rc, 'at record', j); break
    if j % 100 == 0:
        print(j, time.time() - t0, flush=True)
recs = c.records
d = {}
for k in oracle.SNAP_ARRAYS:
    d['rec_' + k] = np.array([r[k] for r in recs])
for k in oracle.SNAP_SCALARS + ['N_active']:
    d['rec_' + k] = np.array([r[k] for r in recs])
for j, st in states.items():
    for k, v in st.items():
        d['state%d_%s' % (j, k)] = np.asarray(v)
d['state_records'] = np.array(sorted(states))
d['stats'] = np.array([c.stat(k) for k in ['getT_calls', 'newton_fr', 'newton_T', 'layer_events', 'flush_calls', 'flood_calls', 'coupling_iters']])
np.savez_compressed(out, **d)
print('done', len(recs), 'records', time.time() - t0, 's; steps', done, 'rc', rc)
