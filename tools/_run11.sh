mkdir -p gpurun_out/r02n
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_parity_fullsize.py -x -q > gpurun_out/r02n/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02n/pytest.log
python tools/ka_bench.py > gpurun_out/r02n/ka_bench.log 2>&1; cat gpurun_out/r02n/ka_bench.log
OTB_KA_TAIL=0 python tools/ka_bench.py > gpurun_out/r02n/ka_bench_notail.log 2>&1; tail -1 gpurun_out/r02n/ka_bench_notail.log
python tools/ka_trace.py 4096 1280 1280 > gpurun_out/r02n/trace_1280.log 2>&1; tail -12 gpurun_out/r02n/trace_1280.log
