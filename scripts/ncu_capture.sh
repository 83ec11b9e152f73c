#!/bin/bash
# ncu --set full captures of the sweep's kernels at sweep 25 of the 10 M benchmark chain (one launch each;
# run only after `python scripts/profile_run.py 1e7 ${BURN:-24} 2 gc` has exited 0 without ncu).
# usage: scripts/ncu_capture.sh TAG "kernel:skip kernel:skip ..."   -> gpurun_out/TAG_ncu_full.txt (+ .ncu-rep of the first kernel)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT /tmp/ncu
python scripts/profile_run.py 1e7 ${BURN:-24} 2 gc > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
: > $OUT/${TAG}_ncu_full.txt
first=1
for ks in $1; do
  k=${ks%%:*}; skip=${ks##*:}
  rep=/tmp/ncu/${TAG}_${k//[^a-zA-Z0-9_]/}.ncu-rep
  ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" --launch-skip $skip -c 1 -f -o ${rep%.ncu-rep} \
      python scripts/profile_run.py 1e7 ${BURN:-24} 2 gc > $OUT/${TAG}_ncu_${k//[^a-zA-Z0-9_]/}.log 2>&1
  python scripts/ncu_metrics.py $rep >> $OUT/${TAG}_ncu_full.txt 2>&1
  if [ $first = 1 ]; then cp $rep $OUT/; first=0; fi
done
cat $OUT/${TAG}_ncu_full.txt
