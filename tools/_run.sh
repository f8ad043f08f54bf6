set -x
timeout 300 python tools/prefill_bench.py > gpurun_out/r02_prefill_default.jsonl 2>&1; tail -1 gpurun_out/r02_prefill_default.jsonl | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_prefill.py tests/test_gpu_gemm_tc.py tests/test_gpu_host_cpp.py -q --timeout 600 -x 2>&1 | tail -3
timeout 300 python tools/prefill_bench.py --model llama3-8b > gpurun_out/r02_prefill_l3.jsonl 2>&1; tail -1 gpurun_out/r02_prefill_l3.jsonl | cut -c1-200
