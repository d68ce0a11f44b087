mkdir -p gpurun_out/r02d
timeout 600 python -m pytest tests/test_gpu_factored_optim.py -x -q > gpurun_out/r02d/pytest_factored.log 2>&1; echo "factored rc=$?"; tail -3 gpurun_out/r02d/pytest_factored.log
for m in adafactor came; do python tools/kprof.py $m 2>&1 | tail -1; done | tee gpurun_out/r02d/kprof_factored.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/shard_check.py > gpurun_out/r02d/shard_check.log 2>&1; echo "shard_check rc=$?"
grep "nvls bfloat16\|shard_check:" gpurun_out/r02d/shard_check.log
for s in 1 0; do
OTB_SHARDED_OPT=$s timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/c4_probe.py 8 > gpurun_out/r02d/c4_n2_sharded$s.log 2>&1; echo "c4 sharded=$s rc=$?"; grep "C4 full" gpurun_out/r02d/c4_n2_sharded$s.log
done
OTB_SHARDED_OPT=1 OTB_KZ_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 tools/c4_probe.py 8 > gpurun_out/r02d/c4_n2_sharded_nooverlap.log 2>&1; grep "C4 full" gpurun_out/r02d/c4_n2_sharded_nooverlap.log
timeout 300 python tools/c4_probe.py 8 > gpurun_out/r02d/c4_n1.log 2>&1; grep "C4 full" gpurun_out/r02d/c4_n1.log
