#!/bin/bash
# ncu --set full capture of a few launches, exported ON THE GPU BOX to small CSV files (the .ncu-rep itself is
# deleted: gpurun merges at most 64 MiB back).
#   tools/ncu_capture.sh OUT_PREFIX KERNEL_REGEX LAUNCH_SKIP LAUNCH_COUNT -- python tools/profile_more.py gmm9
# writes  gpurun_out/OUT_PREFIX.raw.csv      (ncu --page raw: one row per captured launch, all metrics)
#         gpurun_out/OUT_PREFIX.srcN.csv     (ncu --page source for captured launch N: per-SASS-line counters)
#         gpurun_out/OUT_PREFIX.log
set -u
out=$1; regex=$2; skip=$3; count=$4; shift 5
mkdir -p gpurun_out
rep=/tmp/${out}.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:${regex}" --launch-skip "$skip" -c "$count" \
    -f -o "/tmp/${out}" "$@" > "gpurun_out/${out}.log" 2>&1
rc=$?
if [ -f "$rep" ]; then
    ncu -i "$rep" --page raw --csv > "gpurun_out/${out}.raw.csv" 2>/dev/null
    for ((i = 0; i < count; i++)); do
        ncu -i "$rep" --page source --csv --launch-skip "$i" --launch-count 1 2>/dev/null | python -c "
import csv, sys
w = csv.writer(sys.stdout)
for row in csv.reader(sys.stdin):
    w.writerow(row[:8])" > "gpurun_out/${out}.src${i}.csv"
    done
    rm -f "$rep"
fi
exit $rc
