"""Pin the oracle to the reference's SHEBA golden run over its whole length.

Runs the oracle from `init` (testcase 4, open water on 1 July) through all 1,643 output records = 14,191,200 steps
with the committed forcing (tests/golden/forcing_era.npz = the reference's *.txt.input) and compares every record
with tests/golden/sheba_reference.npz (made from reference_output/Reference_SHEBA_with_Version_2):
  * N_active of all 1,643 records (exact),
  * T2m (all printed digits) and T_top (17 printed digits) of all records,
  * the melt / flushing totals (dat_melt, 8 digits).
Usage: python tools/pin_oracle_sheba.py [backend = libm_shebagold] [nrecords = 1643] > profiles/r2_oracle_sheba_pin_<backend>.json
The `*_shebagold` builds carry the one-line snow_precip revision the golden run was made with."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from oracle import oracle

backend = sys.argv[1] if len(sys.argv) > 1 else "libm_shebagold"
nrec = int(sys.argv[2]) if len(sys.argv) > 2 else 1643
gold = np.load(ROOT / "tests" / "golden" / "sheba_reference.npz")
F = np.load(ROOT / "tests" / "golden" / "forcing_era.npz")["sheba"]
oracle.build()
col = oracle.Column(4, backend)
col.set_forcing(*F)
col.record_outputs()
period = col.int("i_time_out") + 1
t0 = time.time()
rc = col.step((nrec - 1) * period + 1)
wall = time.time() - t0
recs = col.records
n = len(recs)
Na = np.array([r["N_active"] for r in recs])
gNa = gold["N_active"][:n]
tt = gold["T2m_T_top"][:n]
Ttop = np.array([r["T_top"] for r in recs]); T2m = np.array([r["T2m"] for r in recs])
melt = np.array([[r["melt_thick_output1"], r["melt_thick_output2"], r["melt_thick_output3"]] for r in recs])
gmelt = gold["melt"][:n]
dT = np.abs(Ttop - tt[:, 1])
relmelt = np.abs(melt - gmelt) / np.maximum(np.abs(gmelt), 1e-12)
first_Na = int(np.argmax(Na != gNa)) + 1 if (Na != gNa).any() else None
big = np.nonzero(dT > 1e-6)[0]
out = {
    "backend": backend, "records": n, "steps": (n - 1) * period + 1, "status": rc, "wall_s": round(wall, 1),
    "N_active_equal_records": int((Na == gNa).sum()), "N_active_first_difference_record": first_Na,
    "N_active_differing_records": [int(j) + 1 for j in np.nonzero(Na != gNa)[0][:20]],
    "T2m_equal_records": int((T2m == tt[:, 0]).sum()),
    "T_top_max_abs_diff_records_1_to_347": float(dT[:347].max()), "T_top_first_record_above_1e-6": (int(big[0]) + 1) if len(big) else None,
    "T_top_records_within_1e-9": int((dT <= 1e-9).sum()), "T_top_records_within_1e-3": int((dT <= 1e-3).sum()),
    "melt_max_rel_diff_records_1_to_347": float(relmelt[:347].max()),
    "events": {k: v for k, v in col.event_counts().items() if v > 0},
}
print(json.dumps(out, indent=1))
