#!/bin/bash
# ncu evidence for one round: launch list of a step + --set full captures of the hot kernels.  usage: gpu_profile.sh <tag> [modes...]
tag=${1:-r02}; shift
modes=${@:-"ka ka1280 kaqkv kgradb kb1 kb2 kc_bf16 adafactor came dora"}
out=gpurun_out/$tag
mkdir -p $out
python tools/step_profile.py sdxl > $out/step_plain.log 2>&1 || { echo "step_profile failed"; tail -5 $out/step_plain.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_step.csv \
    python tools/step_profile.py sdxl > $out/step_ncu.log 2>&1; echo "launch list rc=$?"
for m in $modes; do
  python tools/kprof.py $m > $out/kprof_$m.log 2>&1 || { echo "$m plain run failed"; tail -3 $out/kprof_$m.log; continue; }
  cat $out/kprof_$m.log | tail -1
  timeout 600 ncu --set full --clock-control none --import-source on -c 6 -k "regex:lora_|noise_|loss_|adamw_|adafactor|came|dora|factored" \
      -o $out/ncu_$m -f python tools/kprof.py $m 2 > $out/ncu_$m.log 2>&1; echo "ncu $m rc=$?"
done
ls -la $out
