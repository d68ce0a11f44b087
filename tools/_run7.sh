mkdir -p gpurun_out/r02i
for m in kac640 ka1280 ka; do
python tools/kprof.py $m | tail -1
ncu --set full --clock-control none --import-source on -c 1 -k regex:lora_linear -o /tmp/ncu_$m -f python tools/kprof.py $m 2 > gpurun_out/r02i/ncu_$m.log 2>&1
python tools/ncu_summary.py lines /tmp/ncu_$m.ncu-rep > gpurun_out/r02i/ncu_${m}_lines.md
python tools/ncu_summary.py rep /tmp/ncu_$m.ncu-rep > gpurun_out/r02i/ncu_${m}.md
done
