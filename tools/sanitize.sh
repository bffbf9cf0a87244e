#!/bin/bash
# compute-sanitizer pass over the shared-memory QR / Jacobi kernels (race detection / sanitizers).
# ONE tool per gpurun call (B200_PROFILING.md): run as
#     gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      (or racecheck / synccheck / initcheck)
# on the smallest cases that reach every kernel family: the smoke test (small leaves, Jacobi, larfb, GEMM, inner product) and the
# forced cluster leaf.  Output: gpurun_out/sanitizer_<tool>.log; the summary line is copied into profiles/ by hand.
# NOTE (round 2): this pool's gpurun refuses compute-sanitizer ("closed on this pool ... runs under it have left GPUs needing a reset",
# exit code 86), so the script could not be run here; the same cases run WITHOUT the tool as the first step below and in tests/.
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
del os.environ['QTT_FORCE_CLUSTER_LEAF']
# round-2 kernels: gemm_push_kernel (both B layouts: operator rank 20 > 16 columns, head compression), jacobi_svd_kernel with its vectors
# in shared memory and more than 8 pairs, push_t reading R in place, tsqr2_leaf_kernel + larfb_ts_kernel (400-row panels, threshold 330)
import numpy as np, torch
from oracle import tt_oracle as O
from oracle.make_golden import random_tt
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine
rng = np.random.default_rng(3)
L, N = 6, 2
op = random_tt(rng, [(4, 4)] * L, [20] * (L - 1), scale=0.2)
xs = [random_tt(rng, [(4,)] * L, [4] * (L - 1)) for _ in range(N)]
eng = Engine('cuda:0')
Y = eng.round_sum([(DeviceMPO(op, eng.device), BatchedTT.from_host(xs, eng.device))], caps=24, rel_eps=1e-14).to_host()
for i in range(N):
    ref = O.reapprox(O.matmul(op, xs[i]), ranks_new=24, rel_error=1e-14)
    err = np.linalg.norm(O.dense(Y[i], False) - O.dense(ref, False)) / np.linalg.norm(O.dense(ref, False))
    assert err < 1e-11, err
eng.close()
os.environ['QTT_TS_MIN_ROWS'] = '330'
e2 = Engine('cuda:0')
del os.environ['QTT_TS_MIN_ROWS']
ts = [random_tt(rng, [(4,)] * 5, [100] * 4, scale=0.3) for _ in range(N)]
n2 = e2.norm2(BatchedTT.from_host(ts, e2.device)).cpu().numpy()
for i in range(N):
    assert abs(n2[i] - O.norm_squared(ts[i])) < 1e-11 * n2[i]
R = e2.round_sum([(None, BatchedTT.from_host(ts, e2.device))], caps=5, rel_eps=1e-16).to_host()
e2.close()
print('round-2 kernel cases ok')
" > "gpurun_out/sanitizer_${TOOL}.log" 2>&1
rc=$?
tail -15 "gpurun_out/sanitizer_${TOOL}.log"
echo "compute-sanitizer --tool $TOOL exit code $rc"
exit $rc
