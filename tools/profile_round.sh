#!/bin/bash
# One gpurun call: launch lists (gpu__time_duration) + one `--set full` capture of the top kernels.
# Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
R=${1:-r01}
O=gpurun_out/$R
mkdir -p $O
P="python tools/decode_probe.py"
KR='regex:gemm|attn|attention|rmsnorm|qkv_post|sample|embedding|advance'
$P --fast 1 --batch 64 --ctx 2048 --steps 2 > $O/plain_fast_b64.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 700 --csv --log-file $O/launches_fast_b64.csv \
    $P --fast 1 --batch 64 --ctx 2048 --steps 2 > $O/ncu_fast_b64.log 2>&1
$P --fast 0 --batch 1 --ctx 96 --steps 2 > $O/plain_ref_b1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 700 --csv --log-file $O/launches_ref_b1.csv \
    $P --fast 0 --batch 1 --ctx 96 --steps 2 > $O/ncu_ref_b1.log 2>&1
$P --fast 1 --batch 64 --ctx 2048 --steps 1 > $O/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:attn_decode_fast|gemm_tcgen05' -s 12 -c 8 -o $O/prof_fast \
    $P --fast 1 --batch 64 --ctx 2048 --steps 1 > $O/ncu_full.log 2>&1
ls -la $O
