for pf in 1 2 3 4 6 8; do for r in 200 345; do echo "== pf $pf state $r"; python tools/gpu_quick_bench.py 303104 64 $r $pf 2>&1 | grep "column-steps" | tail -2; done; done
for v in g56 g48 b256 b1024; do for r in 200 345; do echo "== variant $v state $r"; SAMSIM_B200_LIB=samsim_b200/_lib/variants/$v.so python tools/gpu_quick_bench.py 303104 64 $r 2>&1 | grep "column-steps" | tail -2; done; done
