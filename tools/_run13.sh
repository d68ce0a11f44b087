mkdir -p gpurun_out/r02p
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/c4_probe.py 8 > gpurun_out/r02p/c4_n8_$name.log 2>&1
  echo "$name: $(grep 'C4 full' gpurun_out/r02p/c4_n8_$name.log | sed 's/.*GPU: //;s/, loss.*//')"; }
run kz_default OTB_SHARDED_OPT=1
run nccl_replicated OTB_SHARDED_OPT=0
run kz_overlap_b592 OTB_SHARDED_OPT=1 OTB_KZ_OVERLAP=1
run kz_p2p OTB_SHARDED_OPT=1 OTB_KZ_MODE=p2p
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 tools/shard_check.py big > gpurun_out/r02p/shard_check_n8.log 2>&1; echo "shard_check rc=$?"; grep "GPUs,\|shard_check:\|nvls bfloat16" gpurun_out/r02p/shard_check_n8.log | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02p/bench_n8.json 2> gpurun_out/r02p/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p/bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(d.get('exchange_check'))
for c in d.get('configs',[]): print({k:c[k] for k in c if k in ('config','ms_per_step','steps_per_sec','exchange','error')})
PY
