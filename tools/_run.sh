set -x
timeout 900 python -m pytest tests/test_gpu_sampling.py -q --timeout 600 > gpurun_out/r02_t_sampling.log 2>&1; tail -40 gpurun_out/r02_t_sampling.log
