mkdir -p gpurun_out
./scripts/tc_gemm_probe > gpurun_out/r2w_tc_gemm_probe.log 2>&1
./scripts/fp64_rate_probe > gpurun_out/r2w_fp64_rate_probe.log 2>&1
timeout 300 python scripts/prefill_profile.py moshi7b q4_k > gpurun_out/r2w_prefill_profile.log 2>&1
timeout 300 python scripts/prefill_bench.py moshi7b q4_k 1024 >> gpurun_out/r2w_prefill_profile.log 2>&1
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
tail -c 3000 gpurun_out/r2w_bench.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_gpu_tests.log 2>&1
tail -3 gpurun_out/r2w_gpu_tests.log
