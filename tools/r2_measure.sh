#!/bin/bash
# round-2 measurement session on one B200: bench lines, ncu launch list, ncu full captures (run on the GPU box)
set -x
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2_plain.json 2> $O/r2_plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r2_launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_trace -s 7 -c 1 -f -o $O/r2_prof_wf_trace \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2_ncu_t.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_shade -s 7 -c 1 -f -o $O/r2_prof_wf_shade \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/r2_ncu_s.log 2>&1
python bench.py --workload c1 --steps 5 --warmup 3 > $O/r2_bench_c1.json 2> $O/r2_bench_c1.err
python bench.py --workload c2 --steps 3 --warmup 3 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err
python bench.py --workload c5 --steps 5 --warmup 3 > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err
python bench.py --workload c4 --scale 0.125 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_bench_c4_eighth.json 2> $O/r2_bench_c4.err
python bench.py --workload c1 --integrator whitted --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_c1_whitted.json 2> $O/r2_bench_whitted.err
python bench.py --impl reference --workload c1 --steps 2 --warmup 0 > $O/r2_bench_ref_c1.json 2> $O/r2_bench_ref_c1.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/r2_bench_ref_c3.json 2> $O/r2_bench_ref_c3.err
tail -c 300 $O/r2_bench_*.err
