#!/bin/bash
# round-1 refresh after the plane-skip change of the lattice decode kernel: decode captures (lattice640k, roi),
# launch list of the default bench command, final default bench line (run under gpurun, one GPU)
set -x
O=gpurun_out
python tools/prof_decode.py lattice640k 6 grid > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_grid -s 4 -c 1 -f -o $O/prof_decode_grid_lattice \
    python tools/prof_decode.py lattice640k 6 grid > $O/ncu1.log 2>&1
python tools/prof_decode.py roi 6 grid > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_grid -s 4 -c 1 -f -o $O/prof_decode_grid_roi \
    python tools/prof_decode.py roi 6 grid > $O/ncu1b.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 > $O/ncu_bench.log 2>&1
python bench.py > $O/bench_r1_final.log 2>&1
tail -c 300 $O/bench_r1_final.log
