mkdir -p gpurun_out/r02y
run() { name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/c4_probe.py 8 > gpurun_out/r02y/c4_n4_$name.log 2>&1
  echo "$name: $(grep 'C4 full' gpurun_out/r02y/c4_n4_$name.log | sed 's/.*GPU: //;s/, loss.*//')"; }
run nvls_nooverlap OTB_KZ_OVERLAP=0
run p2p_nooverlap OTB_KZ_MODE=p2p OTB_KZ_OVERLAP=0
