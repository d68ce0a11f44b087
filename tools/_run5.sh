mkdir -p gpurun_out/r02e
run() { # name env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/c4_probe.py 8 > gpurun_out/r02e/c4_$name.log 2>&1
  echo "$name: $(grep 'C4 full' gpurun_out/r02e/c4_$name.log | sed 's/.*GPU: //;s/, loss.*//')"
}
run ov_b8 OTB_KZ_OVERLAP=1 OTB_KZ_BLOCKS=8
run ov_b16 OTB_KZ_OVERLAP=1 OTB_KZ_BLOCKS=16
run ov_b32 OTB_KZ_OVERLAP=1 OTB_KZ_BLOCKS=32
run ov_b64 OTB_KZ_OVERLAP=1 OTB_KZ_BLOCKS=64
run noov_b64 OTB_KZ_OVERLAP=0 OTB_KZ_BLOCKS=64
run noov_b148 OTB_KZ_OVERLAP=0 OTB_KZ_BLOCKS=148
run noov_b592 OTB_KZ_OVERLAP=0 OTB_KZ_BLOCKS=592
run ov_b32_bk32 OTB_KZ_OVERLAP=1 OTB_KZ_BLOCKS=32 OTB_KZ_BUCKETS=32
timeout 300 python -m pytest tests/test_gpu_dora.py -x -q 2>&1 | tail -2
python tools/kprof.py dora 2>&1 | tail -1
