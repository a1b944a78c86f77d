#!/bin/bash
# usage: ncu_k3.sh <kind> <lanes> <tag>
ncu --set full --import-source on --clock-control none -k regex:k_huff_decode_sync -s 3 -c 1 -o gpurun_out/$3 -f python profiles/stage_times.py $1 $2 > gpurun_out/$3.log 2>&1
tail -2 gpurun_out/$3.log
