#!/bin/bash
# usage: ncu_k1.sh <variant ldg|bulk> <tag>
GPUJPEG_B200_K1=$1 ncu --set full --import-source on --clock-control none -k regex:k_fdct_rgb444 -s 3 -c 1 -o gpurun_out/$2 -f python profiles/k1_times.py > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
