timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ws16 or output_proj" -s 2>&1 | tail -60 > gpurun_out/c2_tests.log; tail -45 gpurun_out/c2_tests.log
timeout 100 python tools/outproj_probe.py 2>&1 | tail -3
timeout 300 python tools/wmsa_microbench.py > gpurun_out/c2_wmsa_microbench.json 2> gpurun_out/c2_wmsa_microbench.err; echo "microbench rc=$?"; tail -3 gpurun_out/c2_wmsa_microbench.err
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/c2_bench.json; echo
