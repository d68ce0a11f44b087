#!/bin/bash
# ncu evidence for one round: launch list of a step + --set full captures of the hot kernels, SUMMARISED ON THE BOX
# (gpurun brings back at most 64 MiB: the .ncu-rep files stay behind except the ones named in KEEP).
# usage: gpu_profile.sh <tag> [modes...]
tag=${1:-r02}; shift
modes=${@:-"ka ka1280 kaqkv kgradb kb1 kb2 kc_bf16 adafactor came dora"}
KEEP=${KEEP:-"ka1280"}
out=gpurun_out/$tag
tmp=/tmp/ncu_$tag
mkdir -p $out $tmp
python tools/step_profile.py sdxl > $out/step_plain.log 2>&1 || { echo "step_profile failed"; tail -5 $out/step_plain.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $tmp/launches_step.csv \
    python tools/step_profile.py sdxl > $out/step_ncu.log 2>&1; echo "launch list rc=$?"
python tools/ncu_summary.py list $tmp/launches_step.csv > $out/launches_step.md 2>$out/launches_step.err
for m in $modes; do
  python tools/kprof.py $m > $out/kprof_$m.log 2>&1 || { echo "$m plain run failed"; tail -3 $out/kprof_$m.log; continue; }
  tail -1 $out/kprof_$m.log
  timeout 600 ncu --set full --clock-control none --import-source on -c 3 -k "regex:lora_|noise_|loss_|adamw_|adafactor|came|dora|factored" \
      -o $tmp/ncu_$m -f python tools/kprof.py $m 2 > $out/ncu_$m.log 2>&1; echo "ncu $m rc=$?"
  python tools/ncu_summary.py rep $tmp/ncu_$m.ncu-rep > $out/ncu_$m.md 2>$out/ncu_$m.err
  for k in $KEEP; do [ "$k" = "$m" ] && cp $tmp/ncu_$m.ncu-rep $out/; done
done
du -sh $out; ls $out
