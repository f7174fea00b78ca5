#!/bin/bash
OUT=gpurun_out/${1:-tcfg}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_gpu.log
python tools/profile_forward.py --batch 8 --model swin_b_full > $OUT/profile_full_b8.txt 2>&1; sed -n 3,12p $OUT/profile_full_b8.txt
python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1; sed -n 3,5p $OUT/profile_b8.txt; grep -E "msda|mha" $OUT/profile_b8.txt
