#!/bin/bash
# Round-2 evidence, part 2 (one B200): final bench lines, ncu launch list of the bench command, ncu --set full captures.
O=gpurun_out; mkdir -p $O
python bench.py > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; cut -c1-160 $O/r02_bench_c4.json
python bench.py --impl reference > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; cut -c1-160 $O/r02_bench_reference_arm.json
python bench.py --workload c3 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; cut -c1-160 $O/r02_bench_c3.json
python bench.py --steps 2 --warmup 3 --no-long-window --no-cpu-baseline > $O/r02_bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-long-window --no-cpu-baseline > $O/r02_ncu_launches.log 2>&1
wc -l $O/r02_launches_bench.csv
scripts/ncu_capture.sh r02 "k_rounds_tail:24 k_cand_fast:49 k_prep:24 k_entities:24 k_entities_general:24 k_entities_multi:24 k_chain_build:72 k_cand_long:24 k_wide_turns:24 k_wide_emit:24 k_distort:24"
