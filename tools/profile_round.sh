#!/usr/bin/env bash
# One gpurun call: launch list of one factorization (time + DRAM bytes per launch) and ncu --set full captures of
# the kernels named in VERDICT/DESIGN, exported to CSV on the box (the .ncu-rep files are too large to bring back).
set -x
O=gpurun_out
A="python tools/ab_flags.py lap2d_1024 0"
exp () { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null ; rm -f $O/$1.ncu-rep ; }
timeout 200 $A > $O/r02w_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 2102 -c 2102 --csv --log-file $O/r02w_launches.csv $A > $O/r02w_ncu0.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"k_pack|k_assemble|k_zero_fronts|k_front_small|k_front_setup|k_front_finish" -s 150 -c 14 -f -o $O/r02w_asm $A > $O/r02w_ncu1.log 2>&1 ; exp r02w_asm
timeout 400 ncu --set full --clock-control none -k regex:k_update_dmma -s 1150 -c 2 -f -o $O/r02w_update $A > $O/r02w_ncu2.log 2>&1 ; exp r02w_update
timeout 400 ncu --set full --clock-control none -k regex:"k_panel_grid|k_wide_vtc|k_wide_apply_rows" -s 300 -c 4 -f -o $O/r02w_grid python tools/ab_flags.py lap3d_64 0 > $O/r02w_ncu3.log 2>&1 ; exp r02w_grid
ls -la $O ; du -sh $O
