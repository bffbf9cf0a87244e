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
# 2. full captures of the dominant kernels, each at its LONGEST launch (ncu's -k matches the base name without template
#    arguments, so the launch index among the same-named launches is read from the launch list)
idx() { python - "$OUT/${TAG}_launches.csv" "$1" <<'PY'
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
best, k = (-1.0, 0), 0
for r in rows[hdr + 1:]:
    base = re.sub(r'<.*', '', re.sub(r'\(.*', '', r[4]).replace('void ', ''))
    if base == sys.argv[2]:
        v = float(r[14].replace(',', ''))
        if v > best[0]: best = (v, k)
        k += 1
print(best[1])
PY
}
cap() {   # tag kernel-base-name
    local skip=$(idx $2)
    ncu --set full --clock-control none --import-source on -k "regex:^$2\$" --launch-skip $skip --launch-count 1 -f -o $OUT/${TAG}_ncu_$1 $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
    echo "$1 (launch $skip of $2) rc $?"
    ncu -i $OUT/${TAG}_ncu_$1.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$1_raw.csv 2>/dev/null
}
cap push gemm_push_kernel
cap larfb2 larfb2_kernel
cap leaf qr_leaf2_kernel
ls -la $OUT | tail -12
