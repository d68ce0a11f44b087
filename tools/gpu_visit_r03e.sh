#!/bin/bash
# round 2, last visit: the SR-helper tests (one new: against the unmodified reference file on the GPU) and a short bench line
O=gpurun_out/r03e; mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_sr_helpers.py -x -q -m gpu 2>&1 | tail -5 > $O/sr_tests.log; echo "pytest rc=${PIPESTATUS[0]}" > $O/rc.txt
timeout 150 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/bench_short.json 2> $O/bench_short.err; echo "bench rc=$?" >> $O/rc.txt
cat $O/sr_tests.log $O/rc.txt; cut -c1-1500 $O/bench_short.json; tail -3 $O/bench_short.err
