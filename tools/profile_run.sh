#!/bin/bash
# One evidence run on the GPU box (under gpurun): the default bench line, the ncu launch list of
# the short bench command, and full ncu captures of the dominant kernels, reduced ON THE BOX to the
# raw-metric and per-instruction CSV pages (the .ncu-rep files are too large to travel back).
#   bash tools/profile_run.sh [bench|launches|admix3|dense|mix|digit|digit5 ...]
set -u
O=gpurun_out
mkdir -p $O
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --no-other"
pages() {       # $1 = report stem
    ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
    ncu -i $O/$1.ncu-rep --page source --csv --print-source sass > $O/$1_src.csv 2>/dev/null
    rm -f $O/$1.ncu-rep
}
for what in "$@"; do
case $what in
bench)
    python bench.py > $O/final1.json 2> $O/final1.err ;;
launches)
    $SHORT > $O/plain_l.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file $O/launches_r2.csv $SHORT > $O/ncu_l.log 2>&1 ;;
admix3)
    $SHORT > $O/plain_f.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:admix3 -s 3 -c 1 \
        -o $O/prof_r2f $SHORT > $O/ncu_f.log 2>&1
    pages prof_r2f ;;
dense)
    CMD="python tools/dense_time.py c5 --steps 1"
    $CMD > $O/plain_d.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 3 -c 1 \
        -o $O/prof_r2d $CMD > $O/ncu_d.log 2>&1
    pages prof_r2d ;;
mix)        # the DMMA mixture kernels (the planner's own choice is `digit`)
    CMD="python tools/dense_time.py c2 --steps 2 --kernel 3"
    $CMD > $O/plain_m.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 6 -c 2 \
        -o $O/prof_r2m $CMD > $O/ncu_m.log 2>&1
    pages prof_r2m ;;
digit)
    CMD="python tools/dense_time.py c2 --steps 2"
    $CMD > $O/plain_g.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:digit_kernel -s 6 -c 2 \
        -o $O/prof_r2g $CMD > $O/ncu_g.log 2>&1
    pages prof_r2g ;;
digit5)
    CMD="python tools/dense_time.py c5mix --steps 1"
    $CMD > $O/plain_h.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:digit_kernel -s 6 -c 2 \
        -o $O/prof_r2h $CMD > $O/ncu_h.log 2>&1
    pages prof_r2h ;;
esac
done
ls -la $O | head -30
