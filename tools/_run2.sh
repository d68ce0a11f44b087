mkdir -p gpurun_out/r02b
timeout 900 python -m pytest tests/test_ref_callers.py -m gpu -x -q > gpurun_out/r02b/pytest_callers.log 2>&1; echo "callers rc=$?"
tail -15 gpurun_out/r02b/pytest_callers.log
timeout 600 python tools/ka_bench.py 0 0 ref > gpurun_out/r02b/ka_bench_ref.log 2>&1; echo "ka_bench rc=$?"
cat gpurun_out/r02b/ka_bench_ref.log
bash tools/gpu_profile.sh r02prof
