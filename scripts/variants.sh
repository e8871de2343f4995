#!/bin/bash
# Times kernel variants built under build/ (CVS_B200_LIB override) with scripts/seq_probe.py; run on the GPU box.
# usage: variants.sh name...      env PROBE_ARGS: extra arguments of seq_probe.py, TAG: label suffix
set -u
mkdir -p gpurun_out
out=gpurun_out/variants.log
for lib in "$@"; do
  for d in 10000 100000 500000; do
    echo -n "$lib${TAG:-} " >> $out
    CVS_B200_LIB=$PWD/build/libcvs_$lib.so timeout 300 python scripts/seq_probe.py --density $d --frames ${FRAMES:-300} ${PROBE_ARGS:-} 2>&1 | tail -1 >> $out
  done
done
