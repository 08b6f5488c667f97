#!/bin/bash
# refresh of the launch list after bench.py gained the FP32 probe (one B200); kernels unchanged since r2_measure2.sh
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2g_plain.json 2> $O/r2g_plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r2g_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2g_ncu_l.log 2>&1
python bench.py --steps 5 --warmup 3 > $O/r2g_bench_c3.json 2> $O/r2g_bench_c3.err
tail -c 300 $O/r2g_*.err
