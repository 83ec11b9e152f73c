#!/bin/bash
# Round-2 evidence, part 1 (one B200): GPU tests, bench lines, windows with traces, soak.  Outputs in gpurun_out/r02_*.
O=gpurun_out; mkdir -p $O
python -m pytest tests -q -m gpu > $O/r02_gputests.txt 2>&1; tail -2 $O/r02_gputests.txt
python bench.py > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; cut -c1-200 $O/r02_bench_c4.json
python bench.py --impl reference > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; cut -c1-200 $O/r02_bench_reference_arm.json
python bench.py --distort-prior beta11 > $O/r02_bench_c4_beta11.json 2> $O/r02_bench_c4_beta11.err; cut -c1-200 $O/r02_bench_c4_beta11.json
python bench.py --workload c3 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; cut -c1-200 $O/r02_bench_c3.json
python scripts/windows_probe.py 1e7 gen_coupon "25:10,200:10" > $O/r02_windows.out 2> $O/r02_windows.trace; cat $O/r02_windows.out | cut -c1-250
(python scripts/soak.py 1e7 600 100 gen_coupon; python scripts/soak.py 1e6 1000 250 pitman_yor) > $O/r02_soak.txt 2>&1; tail -3 $O/r02_soak.txt
compute-sanitizer --version > $O/r02_sanitizer.txt 2>&1; compute-sanitizer --tool memcheck python -c "print(1)" >> $O/r02_sanitizer.txt 2>&1; tail -2 $O/r02_sanitizer.txt
