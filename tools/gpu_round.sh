#!/bin/bash
# One GPU-box visit: tests, smoke, both bench arms, ncu launch list.  Logs -> gpurun_out/<tag>/
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?" | tee -a $out/rc.txt
tail -c 3000 $out/pytest_gpu.log
tail -c 600 $out/smoke.log
head -c 6000 $out/bench.json
tail -c 1500 $out/bench.err
