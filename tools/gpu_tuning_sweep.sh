#!/bin/bash
# Throughput of library variants (tools/build_variant.py) on one box: ./tools/gpu_tuning_sweep.sh "default v1 v2" "200 345"
for rep in 1 2; do
for v in $1; do for r in $2; do
  if [ "$v" = default ]; then L=samsim_b200/_lib/libsamsim_b200.so; else L=samsim_b200/_lib/variants/$v.so; fi
  echo "== variant $v state $r: $(SAMSIM_B200_LIB=$L python tools/gpu_quick_bench.py 303104 64 $r 2>&1 | grep column-steps | tail -1)"
done; done; done
