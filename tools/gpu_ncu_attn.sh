#!/bin/bash
OUT=gpurun_out/${1:-ncuattn}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_mma -s 29 -c 1 -o $OUT/prof_wattn python tools/profile_forward.py --batch 8 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof_wattn.ncu-rep --page details > $OUT/prof_wattn_details.txt 2>/dev/null
ncu -i $OUT/prof_wattn.ncu-rep --page raw --csv > $OUT/prof_wattn_raw.csv 2>/dev/null
ncu -i $OUT/prof_wattn.ncu-rep --page source --csv --print-source sass > $OUT/prof_wattn_source.csv 2>/dev/null
grep -E "Duration|Issue Slots Busy|Registers Per|Achieved Occ|Grid Size|No Eligible|Warp Cycles Per Issued" $OUT/prof_wattn_details.txt
