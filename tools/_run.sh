set -x
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_w4a16.py -q --timeout 600 -x > gpurun_out/r02_t_llama.log 2>&1; tail -6 gpurun_out/r02_t_llama.log
TCE_PK_DEBUG=1 python tools/pk_timeline.py --ctx 2048 --out gpurun_out/r02_pk_timeline_ctx2048.json 2>/dev/null > gpurun_out/r02_pk_timeline.txt; cat gpurun_out/r02_pk_timeline.txt
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_pk1.json 2> gpurun_out/r02_bench_pk1.err; cat gpurun_out/r02_bench_pk1.json; tail -3 gpurun_out/r02_bench_pk1.err
