#!/bin/bash
# A/B of library variants on the headline step: tools/ab_variants.sh <tag> <spec> [<spec> ...]
#   spec = <variant>[@ENV=VALUE]   ("main" = the production build, else libscc_b200_<variant>.so)
# Bench lines (driver flags, no extras) twice per spec; with SCAN=1 the size scan, with TESTS=1 the GPU parity tests
# of the DEC kernels.  Output: gpurun_out/<tag>_ab.txt (+ <tag>_scan_<spec>.txt)
tag=$1; shift
L=$PWD/spectrogram_cube_clustering_b200
mkdir -p gpurun_out
for spec in "$@"; do
    v=${spec%%@*}; envs=""; [ "$spec" != "$v" ] && envs=${spec#*@}
    lib=$L/libscc_b200.so; [ "$v" != main ] && lib=$L/libscc_b200_$v.so
    for i in 1 2; do
        env SCC_LIB=$lib $envs timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu 2>gpurun_out/${tag}_err.txt \
            | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$spec', 'step_us', round(d['ms_per_step']*1e3,2), 'kernel_us', round(d['roofline']['kernels_ms']['dec_step']*1e3,2), 'frac', round(d['roofline']['frac'],4), 'sm_mhz', d['clocks']['sm_mhz'])" \
            || tail -n 5 gpurun_out/${tag}_err.txt
    done
    [ -n "$SCAN" ] && env SCC_LIB=$lib $envs timeout 300 python tools/scan_sizes.py > gpurun_out/${tag}_scan_$spec.txt 2>&1
    if [ -n "$TESTS" ]; then
        env SCC_LIB=$lib $envs timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -n 3
    fi
done 2>&1 | tee gpurun_out/${tag}_ab.txt
