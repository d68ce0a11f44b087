mkdir -p gpurun_out/r02z
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02z/bench_n2.json 2> gpurun_out/r02z/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z/bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['exchange_check']['ok'])
for c in d.get('configs',[]): print({k:c[k] for k in c if k in ('config','ms_per_step','steps_per_sec','exchange','error')})
PY
timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -2
