set -x
O=gpurun_out
python tools/prof_lift.py 4 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lift_kernel -s 2 -c 1 -f -o $O/prof_lift \
    python tools/prof_lift.py 4 > $O/ncu4.log 2>&1
python tools/prof_encode.py 4 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_reduce -s 3 -c 1 -f -o $O/prof_encode_s1 \
    python tools/prof_encode.py 4 > $O/ncu3.log 2>&1
python tools/prof_encode.py 3 350000 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_reduce -s 1 -c 1 -f -o $O/prof_enc350k_split \
    python tools/prof_encode.py 3 350000 > $O/ncu5.log 2>&1
python bench.py --steps 20 --warmup 3 > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 20 --warmup 3 > $O/ncu_bench.log 2>&1
python bench.py --steps 400 --warmup 10 > $O/bench_r1_final.log 2>&1
tail -c 300 $O/bench_r1_final.log
