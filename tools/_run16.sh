bash tools/gpu_round.sh r02u 2>&1 | head -c 3000
KEEP="" bash tools/gpu_profile.sh r02v ka ka1280 kac640 kaqkv kgradb adafactor came dora 2>&1 | tail -30
for m in ka ka1280 kac640; do
ncu --set full --clock-control none --import-source on -c 1 -k regex:lora_linear -o /tmp/ncul_$m -f python tools/kprof.py $m 2 > /dev/null 2>&1
python tools/ncu_summary.py lines /tmp/ncul_$m.ncu-rep > gpurun_out/r02v/ncu_${m}_lines.md
done
