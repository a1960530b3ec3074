#!/bin/bash
# Round-1 ncu captures (run under gpurun, one GPU). Every ncu run follows a plain run of the same command.
set -x
O=gpurun_out
python bench.py --steps 20 --warmup 3 > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 > $O/ncu_bench.log 2>&1
python tools/prof_decode.py lattice640k 6 grid > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_grid -s 4 -c 1 -f -o $O/prof_decode_grid_lattice \
    python tools/prof_decode.py lattice640k 6 grid > $O/ncu1.log 2>&1
python tools/prof_decode.py roi 6 grid > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_grid -s 4 -c 1 -f -o $O/prof_decode_grid_roi \
    python tools/prof_decode.py roi 6 grid > $O/ncu1b.log 2>&1
python tools/prof_decode.py uniform640k 6 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_kernel -s 4 -c 1 -f -o $O/prof_decode_uniform \
    python tools/prof_decode.py uniform640k 6 > $O/ncu2.log 2>&1
python tools/prof_encode.py 4 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_reduce -s 3 -c 1 -f -o $O/prof_encode_s1 \
    python tools/prof_encode.py 4 > $O/ncu3.log 2>&1
python tools/prof_lift.py 4 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lift_kernel -s 2 -c 1 -f -o $O/prof_lift \
    python tools/prof_lift.py 4 > $O/ncu4.log 2>&1
tail -2 $O/ncu*.log
