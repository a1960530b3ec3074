#!/bin/bash
# Round-2 ncu captures (run under gpurun, one GPU). Every ncu run follows a plain run of the same command.
O=gpurun_out
python tools/prof_surf_sam.py 4 > $O/plain_ss.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sample3_kernel -s 2 -c 1 -f -o $O/r02_range_bs8 python tools/prof_surf_sam.py 4 > $O/ncu_ss1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sample3_seg -s 2 -c 1 -f -o $O/r02_seg python tools/prof_surf_sam.py 4 > $O/ncu_ss2.log 2>&1
python tools/prof_sparse.py 3 > $O/plain_sp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sparse_ -s 5 -c 5 -f -o $O/r02_sparse python tools/prof_sparse.py 3 > $O/ncu_sp.log 2>&1
tail -2 $O/ncu_ss1.log $O/ncu_ss2.log $O/ncu_sp.log $O/plain_ss.log $O/plain_sp.log
