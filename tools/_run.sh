set -x
TCE_W4_GEMM=pair_fused timeout 120 python tools/gemm_pair_check.py > gpurun_out/r02_gemm_pair_fused.jsonl 2>&1; tail -12 gpurun_out/r02_gemm_pair_fused.jsonl
