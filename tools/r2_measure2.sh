#!/bin/bash
# final-build ncu artefacts (one B200): launch list of the bench command + full captures of the two top kernels
set -x
O=gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2f_plain.json 2> $O/r2f_plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r2f_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2f_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 4 -c 1 -f -o $O/r2f_prof_wf_trace \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2f_ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 4 -c 1 -f -o $O/r2f_prof_wf_shade \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2f_ncu_s.log 2>&1
python bench.py --steps 5 --warmup 3 > $O/r2f_bench_c3.json 2> $O/r2f_bench_c3.err
for w in c1 c2 c5; do python bench.py --workload $w --steps 3 --warmup 3 > $O/r2f_bench_$w.json 2> $O/r2f_bench_$w.err; done
tail -c 200 $O/r2f_*.err
