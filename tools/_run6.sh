mkdir -p gpurun_out/r02h
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_parity_fullsize.py tests/test_gpu_dora.py -x -q > gpurun_out/r02h/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02h/pytest.log
python tools/ka_bench.py > gpurun_out/r02h/ka_bench.log 2>&1; cat gpurun_out/r02h/ka_bench.log
OTB_KA_LOOKAHEAD=4 python tools/ka_bench.py > gpurun_out/r02h/ka_bench_la4.log 2>&1; tail -1 gpurun_out/r02h/ka_bench_la4.log
