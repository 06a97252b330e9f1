"""Exact algorithmic FP64 operation count of one column-timestep.

Runs the CPU oracle's operation-counting build (oracle/count_real.h: every + - * / of the reference's arithmetic
counts 1; pow/exp/sin calls counted separately) on the bench workload and on other regimes of the SHEBA run.
Usage: python tools/count_flops.py
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle  # noqa: E402


def main():
    L = oracle.lib("count")
    z = np.load(ROOT / "tests" / "golden" / "sheba_oracle_states.npz")
    F = np.load(ROOT / "tests" / "golden" / "forcing_era.npz")["sheba"]

    def state(j):
        p = f"state{j}_"
        return {k[len(p):]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith(p)}

    print(f"{'state':>6} {'N_active':>8} {'add':>9} {'mul':>9} {'div':>8} {'flop':>9} {'pow':>6} {'exp':>6} {'sin':>4}   per column-timestep (mean of 2000 steps)")
    out = {}
    for rec in (200, 100, 60, 330, 345, 400):
        col = oracle.Column(4, "count")
        col.set_forcing(*F)
        col.load_state(state(rec))
        L.sam_reset_op_counts()
        n = 2000
        assert col.step(n) == 0
        buf = (C.c_longlong * 7)()
        L.sam_get_op_counts(buf)
        add, mul, div, cmp_, pw, ex, sn = [b / n for b in buf]
        flop = add + mul + div
        out[rec] = flop
        print(f"{rec:>6} {col.int('N_active'):>8} {add:>9.0f} {mul:>9.0f} {div:>8.0f} {flop:>9.0f} {pw:>6.1f} {ex:>6.1f} {sn:>4.1f}")
    print(f"\nbench workload (state 200): F_ALG_FLOP_PER_COLUMN_STEP = {out[200]:.0f}")


if __name__ == "__main__":
    main()
