set -x
timeout 900 python -m pytest tests/test_gpu_opt_attention.py tests/test_gpu_callsites.py tests/test_gpu_misc.py tests/test_gpu_host_cpp.py -q --timeout 600 > gpurun_out/r02_t_opt.log 2>&1; tail -5 gpurun_out/r02_t_opt.log
timeout 300 python tools/w8a8_layer_bench.py > gpurun_out/r02_w8a8_layer.json 2>&1; tail -1 gpurun_out/r02_w8a8_layer.json | cut -c1-700
