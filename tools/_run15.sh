mkdir -p gpurun_out/r02r
timeout 600 python -m pytest tests/test_gpu_modules.py -x -q -k "bucketed or graphed" > gpurun_out/r02r/pytest_bucketed.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02r/pytest_bucketed.log | cut -c1-300
KA_STATIC=1 OTB_KA_PREFETCH_W=1 python tools/ka_trace.py 4096 1280 1280 > gpurun_out/r02r/trace_1280_pf1.log 2>&1; tail -12 gpurun_out/r02r/trace_1280_pf1.log
