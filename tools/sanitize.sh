#!/bin/bash
# compute-sanitizer on the smallest parity cases (ONE tool per gpurun call, see B200_PROFILING.md).
# NOTE (round 1): compute-sanitizer is CLOSED on this GPU pool ("runs under it have left GPUs needing a
# reset"), so no sanitizer evidence could be collected; correctness rests on the parity suite
# (ragged sizes, repeat-determinism checks, 81+ GPU tests).
# usage: tools/sanitize.sh memcheck|racecheck|synccheck
set -e
TOOL=${1:-memcheck}
compute-sanitizer --tool "$TOOL" --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "dec_assign_golden and (k5 or d32) or dec_kl_grad_api_mode_golden and (k5 or d32) or gmm_hard or dec_empty" 2>&1 | tail -15
