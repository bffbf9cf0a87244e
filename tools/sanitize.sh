#!/bin/bash
# compute-sanitizer pass over the shared-memory QR / Jacobi kernels (race detection / sanitizers).
# ONE tool per gpurun call (B200_PROFILING.md): run as
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      (or racecheck / synccheck / initcheck)
# on the smallest cases that reach every kernel family: the smoke test (small leaves, Jacobi, larfb, GEMM, inner product) and the
# forced cluster leaf.  Output: gpurun_out/sanitizer_<tool>.log; the summary line is copied into profiles/ by hand.
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitizer_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool "$TOOL" --print-limit 20 --error-exitcode 3 \
    python -c "
import os, sys
sys.path.insert(0, '.')
import __graft_entry__ as g
g.smoke()
os.environ['QTT_FORCE_CLUSTER_LEAF'] = '1'
g.smoke()
" > "gpurun_out/sanitizer_${TOOL}.log" 2>&1
rc=$?
tail -15 "gpurun_out/sanitizer_${TOOL}.log"
echo "compute-sanitizer --tool $TOOL exit code $rc"
exit $rc
