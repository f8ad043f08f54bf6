# launch list of the decode step under ncu (duration + DRAM bytes per launch of the one kernel a step consists of)
mkdir -p gpurun_out
timeout 44 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:decode_persistent -c 10 --csv --log-file gpurun_out/r02_decode_step_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --ctx 2048 > gpurun_out/r02_decode_step_launches.log 2>&1
echo "rc=$?"; tail -c 600 gpurun_out/r02_decode_step_launches.csv
