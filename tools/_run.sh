set -x
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_gemm_tc.py tests/test_gpu_prefill.py -q --timeout 600 -k "range_checked or gemm or prefill" > gpurun_out/r02_t_gemm2.log 2>&1; tail -5 gpurun_out/r02_t_gemm2.log
TCE_W4_GEMM=pair timeout 120 python tools/gemm_pair_check.py > gpurun_out/r02_gemm_pair_v2.jsonl 2>&1; tail -5 gpurun_out/r02_gemm_pair_v2.jsonl
timeout 300 python tools/prefill_bench.py > gpurun_out/r02_prefill_default_v2.jsonl 2>&1; tail -1 gpurun_out/r02_prefill_default_v2.jsonl | cut -c1-330
