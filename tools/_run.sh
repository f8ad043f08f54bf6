set -x
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_attention.py -q --timeout 600 > gpurun_out/r02_t_llama.log 2>&1; tail -4 gpurun_out/r02_t_llama.log
