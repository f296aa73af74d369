bash tools/collect_round.sh r02f "tests smoke bench b512 wmsa ncu"
timeout 100 python tools/outproj_probe.py > gpurun_out/r02f_outproj_probe.log 2>&1; tail -2 gpurun_out/r02f_outproj_probe.log
TAG=r02f bash tools/sanitizer_run.sh
WMSA_MB_ITERS=1 timeout 400 ncu --clock-control none --csv -k regex:wmsa --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct --log-file gpurun_out/r02f_wmsa_microbench_ncu.csv python tools/wmsa_microbench.py > /dev/null 2>&1; echo "ncu microbench rc=$?"
