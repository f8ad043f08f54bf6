set -x
timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_misc.py -q --timeout 600 > gpurun_out/r02_t_llama.log 2>&1; tail -8 gpurun_out/r02_t_llama.log
TCE_PK_DEBUG=1 python tools/pk_timeline.py --ctx 2048 --out gpurun_out/r02_pk_timeline_ctx2048.json > gpurun_out/r02_pk_timeline.txt 2>&1; cat gpurun_out/r02_pk_timeline.txt
TCE_PK_DEBUG=1 TCE_PK_STAGES=5 python tools/pk_timeline.py --ctx 2048 > gpurun_out/r02_pk_timeline_5st.txt 2>&1; cat gpurun_out/r02_pk_timeline_5st.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 3 -c 1 -o gpurun_out/r02_pk_full python bench.py --steps 1 --warmup 3 --ctx 2048 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu.log 2>&1; tail -3 gpurun_out/r02_ncu.log
ls -la gpurun_out/ | tail -8
