#!/bin/bash
TAG=${1:-v1}
mkdir -p gpurun_out/r2
export GROOVE_B200_DEV_FLAGS="-DGT_T256_TIMELINE"
GT_T256_DBG=3 timeout -s KILL 300 python tools/t256_timeline.py c3 4096 > gpurun_out/r2/timeline_c3_$TAG.txt 2>&1
GT_T256_DBG=3 timeout -s KILL 300 python tools/t256_timeline.py c4 4096 > gpurun_out/r2/timeline_c4_$TAG.txt 2>&1
grep -c timeline gpurun_out/r2/timeline_c3_$TAG.txt gpurun_out/r2/timeline_c4_$TAG.txt
