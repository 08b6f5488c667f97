#!/bin/bash
# final bench lines of round 2 after the narrowed upload (one B200); kernels unchanged since r2_measure3.sh,
# so the ncu launch list and captures of r2_measure2/3 stand
set -x
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/r2i_bench_c3.json 2> $O/r2i_bench_c3.err
for w in c1 c2 c5; do python bench.py --workload $w --steps 3 --warmup 3 > $O/r2i_bench_$w.json 2> $O/r2i_bench_$w.err; done
python bench.py --workload c4 --scale 0.125 --steps 3 --warmup 3 --no-cpu-baseline > $O/r2i_bench_c4_eighth.json 2> $O/r2i_bench_c4.err
python bench.py --workload c1 --integrator whitted --steps 5 --warmup 3 --no-cpu-baseline > $O/r2i_bench_c1_whitted.json 2> $O/r2i_bench_whitted.err
python bench.py --impl reference --workload c1 --steps 2 --warmup 0 > $O/r2i_bench_ref_c1.json 2> $O/r2i_bench_ref_c1.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/r2i_bench_ref_c3.json 2> $O/r2i_bench_ref_c3.err
tail -c 200 $O/r2i_bench_*.err
