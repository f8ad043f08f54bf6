set -x
python tools/debug_pk.py tiny-gqa 4 1 > gpurun_out/r02_dbg1.log 2>&1; tail -5 gpurun_out/r02_dbg1.log
python tools/debug_pk.py tiny-mha 3 2 > gpurun_out/r02_dbg2.log 2>&1; tail -4 gpurun_out/r02_dbg2.log
timeout 900 python -m pytest tests/test_gpu_llama.py -q --timeout 600 -x > gpurun_out/r02_t_llama.log 2>&1; tail -8 gpurun_out/r02_t_llama.log
TCE_PK_DEBUG=1 python tools/pk_timeline.py --ctx 2048 --out gpurun_out/r02_pk_timeline_ctx2048.json > gpurun_out/r02_pk_timeline.txt 2>&1; cat gpurun_out/r02_pk_timeline.txt
TCE_PK_DEBUG=1 python tools/pk_timeline.py --ctx 4095 > gpurun_out/r02_pk_timeline_4095.txt 2>&1; cat gpurun_out/r02_pk_timeline_4095.txt
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_pk1.json 2> gpurun_out/r02_bench_pk1.err; cat gpurun_out/r02_bench_pk1.json; tail -3 gpurun_out/r02_bench_pk1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 3 -c 1 -o gpurun_out/r02_pk_full python bench.py --steps 1 --warmup 3 --ctx 2048 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu.log 2>&1; tail -3 gpurun_out/r02_ncu.log
ls -la gpurun_out/ | tail -6
