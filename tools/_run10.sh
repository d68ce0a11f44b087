for m in adafactor came; do python tools/kprof.py $m 2>&1 | tail -1; done
bash tools/gpu_round.sh r02l 2>&1 | head -c 6000
