#!/bin/bash
# One GPU-box session: GPU test suite, smoke(), a short bench line.  Usage (from the repo root, under gpurun):
#   tools/gpu_check.sh TAG [pytest args...]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
tail -8 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/${tag}_bench.err
python tools/show_bench.py gpurun_out/${tag}_bench.json 2>/dev/null | head -60
