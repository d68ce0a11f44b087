mkdir -p gpurun_out/r02k
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py tests/test_gpu_parity_fullsize.py -x -q > gpurun_out/r02k/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02k/pytest.log
python tools/ka_bench.py > gpurun_out/r02k/ka_bench.log 2>&1; cat gpurun_out/r02k/ka_bench.log
python tools/ka_bench.py flux > gpurun_out/r02k/ka_bench_flux.log 2>&1; cat gpurun_out/r02k/ka_bench_flux.log
