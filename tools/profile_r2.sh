#!/bin/bash
# ncu evidence of one fused round(B @ x) sweep (2D L=12, 148 instances, cap 4): launch list + --set full captures of the
# dominant kernels.  Run under gpurun AFTER the same command exited 0 without ncu (B200_PROFILING.md):
#     gpurun --timeout 2400 -- 'bash tools/profile_r2.sh [tag]'
# Output: gpurun_out/<tag>_*.{csv,ncu-rep,log}; summaries are copied into profiles/ by hand.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/dev_time.py 148 4 0"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 $OUT/${TAG}_plain.log; exit 1; }
tail -4 $OUT/${TAG}_plain.log
# 1. launch list (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc $?"
# 2. full captures.  Launch order inside a sweep runs from the last core to the first, so the tall panels (cores 7..4) come late:
#    skip counts pick launches of the full-size cores.
cap() {   # name regex skip count
    ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count $4 -f -o $OUT/${TAG}_ncu_$1 $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
    echo "$1 rc $?"
    ncu -i $OUT/${TAG}_ncu_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1_raw.csv 2>/dev/null
}
cap larfb2 'larfb2_kernel' 12 3
cap leaf1024 'qr_leaf2_kernel<1024>' 8 3
cap gemm56 'gemm_dmma_kernel<64, 56' 2 2
ls -la $OUT | tail -20
