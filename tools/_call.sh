bash tools/collect_round.sh r02f "tests smoke bench b512 wmsa ncu"
timeout 100 python tools/outproj_probe.py > gpurun_out/r02f_outproj_probe.log 2>&1; tail -2 gpurun_out/r02f_outproj_probe.log
TAG=r02f bash tools/sanitizer_run.sh
