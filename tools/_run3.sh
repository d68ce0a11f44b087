mkdir -p gpurun_out/r02c
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/shard_check.py big > gpurun_out/r02c/shard_check.log 2>&1; echo "shard_check rc=$?"
grep -v "Warning\|warn" gpurun_out/r02c/shard_check.log | tail -40
timeout 900 python -m pytest tests/test_ref_callers.py -m gpu -x -q > gpurun_out/r02c/pytest_callers.log 2>&1; echo "callers rc=$?"
grep -v "Warning\|warn" gpurun_out/r02c/pytest_callers.log | tail -15
