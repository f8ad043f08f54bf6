# final round-2 validation call: smoke() of the shipped build, the C++ host tests, ncu --set full of the shipped prefill GEMM kernels
set -x
mkdir -p gpurun_out
timeout 170 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke_final.log
timeout 60 tests/cpp/test_host > gpurun_out/r02_test_host_final.log 2>&1; echo "test_host rc=$?"; tail -3 gpurun_out/r02_test_host_final.log
for v in pair pair_fused; do
  TCE_W4_GEMM=$v timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 2 -c 1 -f -o gpurun_out/r02_gemm_$v python tools/ncu_gemm_target.py > gpurun_out/r02_ncu_gemm_$v.log 2>&1; echo "ncu $v rc=$?"
  ncu -i gpurun_out/r02_gemm_$v.ncu-rep --page raw --csv > gpurun_out/r02_gemm_${v}_raw.csv 2>/dev/null
  python tools/ncu_raw_summary.py gpurun_out/r02_gemm_${v}_raw.csv "TCE_W4_GEMM=$v ncu --set full --clock-control none -k regex:gemm_pair -s 2 -c 1 python tools/ncu_gemm_target.py  (13B q|k|v shape: M 2048, N 15360, K 5120)" > gpurun_out/r02_ncu_gemm_$v.txt 2>&1
  head -12 gpurun_out/r02_ncu_gemm_$v.txt
done
ls -la gpurun_out
timeout 200 python -m pytest tests/test_gpu_llama.py -x -q --timeout 180 > gpurun_out/r02_t_llama_final.log 2>&1; echo "pytest llama rc=$?"; tail -3 gpurun_out/r02_t_llama_final.log
