mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02t_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02t_gpu_tests.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02t_bench.json')); print(d['value'], d['e2e']['value'], {k:v for k,v in d['roofline']['by_kernel_ms'].items() if '512' in k})"
