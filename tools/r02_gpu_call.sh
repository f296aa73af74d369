mkdir -p gpurun_out
timeout 600 python tools/leff_fused_probe.py > gpurun_out/r02a_leff_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r02a_leff_probe.log
UFORMER_B200_LEFF=split timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_gpu_tests_split.log 2>&1; echo "pytest(split) rc=$?"; tail -5 gpurun_out/r02a_gpu_tests_split.log
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02a_gpu_tests_fused.log 2>&1; echo "pytest(fused) rc=$?"; tail -5 gpurun_out/r02a_gpu_tests_fused.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02a_bench_fused.json 2> gpurun_out/r02a_bench_fused.err; echo "bench fused rc=$?"; head -c 1500 gpurun_out/r02a_bench_fused.json; echo
UFORMER_B200_LEFF=split timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02a_bench_split.json 2> gpurun_out/r02a_bench_split.err; echo "bench split rc=$?"; head -c 600 gpurun_out/r02a_bench_split.json; echo
