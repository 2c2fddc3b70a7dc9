#!/bin/bash
# Per-kernel `ncu --set full` evidence in ONE gpurun call, with outputs small enough to travel back (gpurun_out <= 64 MiB):
# the big report stays in /tmp on the box, only its summaries (md / json / raw csv) and one small report with source are kept.
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture.sh'
set -u
B=${1:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_ops_b$B python tools/ncu_ops.py --batch $B > gpurun_out/r2_ncu_ops_b$B.log 2>&1
tail -3 gpurun_out/r2_ncu_ops_b$B.log | cut -c1-300
python tools/ncu_summary.py /tmp/r2_ops_b$B.ncu-rep gpurun_out/r2_kernels_b$B
ncu -i /tmp/r2_ops_b$B.ncu-rep --page raw --csv > gpurun_out/r2_ops_b${B}_raw.csv 2>/dev/null
# the dominant tensor-core kernel with source correlation (3 launches of the 3x3 256->256 @112x224 layer)
SHAPES=0,4 IMPLS=4 NMUL=4 ncu --set full --clock-control none --import-source on -k regex:conv_f16s -s 22 -c 1 -f -o gpurun_out/r2_f16s_1x1 python tools/conv_bench.py 0 > gpurun_out/r2_ncu_f16s_1x1.log 2>&1
SHAPES=4 IMPLS=4 NMUL=4 ncu --set full --clock-control none --import-source on -k regex:conv_f16s -s 22 -c 1 -f -o gpurun_out/r2_f16s_3x3 python tools/conv_bench.py 0 > gpurun_out/r2_ncu_f16s_3x3.log 2>&1
ls -la gpurun_out | tail -8
