#!/bin/bash
# Round-2 evidence (run under gpurun, one GPU): full default bench line, launch lists, ncu --set full of the dominant kernels.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02_gputests.txt
( time python bench.py ) > $O/r02_bench_all_n1.json 2> $O/r02_bench_all_n1.err
python bench.py --workload decode --steps 20 --warmup 3 > $O/plain_dec.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_decode.csv python bench.py --workload decode --steps 20 --warmup 3 > $O/ncu_dec.log 2>&1
python bench.py --workload range_cam --steps 5 --warmup 3 > $O/plain_rc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_range_cam.csv python bench.py --workload range_cam --steps 5 --warmup 3 > $O/ncu_rc.log 2>&1
python tools/prof_decode.py lattice640k 6 grid > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_grid -s 4 -c 1 -f -o $O/r02_decode_grid_lattice python tools/prof_decode.py lattice640k 6 grid > $O/ncu1.log 2>&1
python tools/prof_sparse.py 3 8 > $O/plain_sp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sparse_ -s 5 -c 5 -f -o $O/r02_sparse python tools/prof_sparse.py 3 8 > $O/ncu_sp.log 2>&1
cat $O/r02_gputests.txt; tail -3 $O/r02_bench_all_n1.err
