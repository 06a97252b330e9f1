"""Build a tuning variant of the library: python tools/build_variant.py NAME -DSAMSIM_SYNC=0 ...
-> samsim_b200/_lib/variants/NAME.so (select it with SAMSIM_B200_LIB=<path>).  Not part of the product build."""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from samsim_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
out = B.LIBDIR / "variants"; out.mkdir(parents=True, exist_ok=True)
cmd = ["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + defs + ["-o", str(out / f"{name}.so"), str(B.SRC), str(B.HOST_SRC)]
r = subprocess.run(cmd, capture_output=True, text=True)
print(name, "ok" if r.returncode == 0 else r.stderr[-2000:])
