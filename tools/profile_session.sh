#!/bin/bash
# One GPU-box session that regenerates everything under profiles/ for the current build (run under gpurun):
#   tools/profile_session.sh r02
# 1. bench line (not under a profiler)  2. ncu launch list of the same command  3. ncu --set full of every hot kernel
# 4. in-kernel timeline (needs `make -C .../csrc variant VNAME=tl9 VFLAGS=-DSCC_TIMELINE`)  5. size / variant scans
tag=${1:-r02}
L=spectrogram_cube_clustering_b200
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-extra --no-cpu --single-step-graphs > gpurun_out/${tag}_launches.log 2>&1; echo "launch list rc=$?"
tools/ncu_capture.sh ${tag}_kernels "dec_|gmm_em|gmm_tail" 0 36 -- python tools/profile_all.py; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/${tag}_kernels.raw.csv gpurun_out/${tag}_ncu_kernels.txt gpurun_out/${tag}_traffic_raw.json
if [ -f $L/libscc_b200_tl9.so ]; then
    SCC_LIB=$L/libscc_b200_tl9.so timeout 200 python tools/timeline.py > gpurun_out/${tag}_timeline.txt 2>&1
fi
timeout 200 python tools/scan_sizes.py 9 8 > gpurun_out/${tag}_scan_sizes.txt 2>&1
timeout 200 python tools/d32_scan.py > gpurun_out/${tag}_scan_d32.txt 2>&1
timeout 200 python tools/gmm_scan.py > gpurun_out/${tag}_gmm_scan.txt 2>&1
SCC_GMM_VARIANT=full timeout 200 python tools/gmm_scan.py >> gpurun_out/${tag}_gmm_scan.txt 2>&1
timeout 200 python tools/gmm_scan.py 32 16 2000000 3 >> gpurun_out/${tag}_gmm_scan.txt 2>&1
ls -la gpurun_out | head -60
