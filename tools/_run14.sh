mkdir -p gpurun_out/r02q
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 tools/shard_check.py > gpurun_out/r02q/shard_check_n4.log 2>&1; echo "shard_check rc=$?"
grep -v "Warning\|warn(" gpurun_out/r02q/shard_check_n4.log | grep "rank 0/4\|shard_check:" | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/c4_probe.py 8 > gpurun_out/r02q/c4_n4_kz.log 2>&1; grep "C4 full" gpurun_out/r02q/c4_n4_kz.log | sed 's/.*GPU: //;s/, loss.*//'
