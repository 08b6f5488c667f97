#!/bin/bash
# final build of round 2 (guided batches in k_wf_trace, narrowed upload): launch list, full capture of k_wf_trace,
# bench line (one B200)
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2k_plain.json 2> $O/r2k_plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r2k_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2k_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 7 -c 1 -f -o $O/r2k_prof_wf_trace \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2k_ncu_t.log 2>&1
python bench.py --steps 5 --warmup 3 > $O/r2k_bench_c3.json 2> $O/r2k_bench_c3.err
tail -c 300 $O/r2k_*.err
