mkdir -p gpurun_out/r02j
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_parity_fullsize.py -x -q > gpurun_out/r02j/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02j/pytest.log
python tools/ka_bench.py > gpurun_out/r02j/ka_bench.log 2>&1; cat gpurun_out/r02j/ka_bench.log
KA_STATIC=1 OTB_KA_PREFETCH_W=2 python tools/ka_bench.py > gpurun_out/r02j/ka_bench_l2pf.log 2>&1; cat gpurun_out/r02j/ka_bench_l2pf.log
KA_STATIC=1 OTB_KA_PREFETCH_W=1 python tools/ka_bench.py > gpurun_out/r02j/ka_bench_smempf.log 2>&1; tail -1 gpurun_out/r02j/ka_bench_smempf.log
