#!/bin/bash
# Round-2 evidence, final (run under gpurun, one GPU): tests, the default bench lines, launch lists, ncu --set full of the
# dominant kernels (lattice decode, per-query + segment decode on configs[2], dense encode reduce).
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02c_gputests.txt
( time python bench.py ) > $O/r02c_bench_all_n1.json 2> $O/r02c_bench_all_n1.err
python bench.py --workload decode --steps 20 --warmup 3 > $O/plain_dec.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02c_launches_decode.csv python bench.py --workload decode --steps 20 --warmup 3 > $O/ncu_dec.log 2>&1
python tools/prof_surf_sam.py 3 > $O/plain_ss.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sample3_kernel|sample3_seg" -s 4 -c 2 -f -o $O/r02c_surf_sam python tools/prof_surf_sam.py 3 > $O/ncu_ss.log 2>&1
python tools/prof_decode.py lattice640k 6 grid > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_grid -s 4 -c 1 -f -o $O/r02c_decode_grid_lattice python tools/prof_decode.py lattice640k 6 grid > $O/ncu1.log 2>&1
python tools/prof_encode_dense.py 2 2 > $O/plain_ed.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_reduce -s 1 -c 1 -f -o $O/r02c_encode_dense python tools/prof_encode_dense.py 2 2 > $O/ncu_ed.log 2>&1
python tools/prof_backward_all.py 2 > $O/plain_bw.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"backward|sample3_seg" -s 5 -c 5 -f -o $O/r02c_backward python tools/prof_backward_all.py 2 > $O/ncu_bw.log 2>&1
cat $O/r02c_gputests.txt; tail -3 $O/r02c_bench_all_n1.err
