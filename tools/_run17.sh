mkdir -p gpurun_out/r02w
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/shard_check.py > gpurun_out/r02w/shard_check.log 2>&1; echo "shard_check rc=$?"
grep -v "Warning\|warn(" gpurun_out/r02w/shard_check.log | grep "rank 0/2 step\|shard_check:\|Error\|error" | cut -c1-260
for f in 1 0; do
OTB_KZ_FUSED_GATHER=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/c4_probe.py 8 > gpurun_out/r02w/c4_n2_fused$f.log 2>&1; echo "fused=$f: $(grep 'C4 full' gpurun_out/r02w/c4_n2_fused$f.log | sed 's/.*GPU: //;s/, loss.*//')"
done
OTB_KZ_MODE=nvls OTB_KZ_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 tools/c4_probe.py 8 > gpurun_out/r02w/c4_n2_fused_nvls.log 2>&1; echo "fused nvls: $(grep 'C4 full' gpurun_out/r02w/c4_n2_fused_nvls.log | sed 's/.*GPU: //;s/, loss.*//')"
