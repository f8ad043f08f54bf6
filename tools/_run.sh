set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python tools/debug_pk.py tiny-gqa 4 1 > gpurun_out/r02_dbg1.log 2>&1; tail -8 gpurun_out/r02_dbg1.log
python tools/debug_pk.py tiny-mha 3 2 > gpurun_out/r02_dbg2.log 2>&1; tail -8 gpurun_out/r02_dbg2.log
timeout 900 python -m pytest tests/test_gpu_w4a16.py -q --timeout 300 > gpurun_out/r02_t_w4a16.log 2>&1; tail -15 gpurun_out/r02_t_w4a16.log
timeout 900 python -m pytest tests/test_gpu_llama.py -q --timeout 600 > gpurun_out/r02_t_llama.log 2>&1; tail -25 gpurun_out/r02_t_llama.log
timeout 600 python -m pytest tests/test_gpu_attention.py -q --timeout 300 -k "benchmarked" > gpurun_out/r02_t_attn.log 2>&1; tail -5 gpurun_out/r02_t_attn.log
TCE_PERSISTENT=1 timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_pk1.json 2> gpurun_out/r02_bench_pk1.err; cat gpurun_out/r02_bench_pk1.json; tail -3 gpurun_out/r02_bench_pk1.err
TCE_PERSISTENT=0 timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/r02_bench_pk0.json 2> gpurun_out/r02_bench_pk0.err; cat gpurun_out/r02_bench_pk0.json; tail -3 gpurun_out/r02_bench_pk0.err
