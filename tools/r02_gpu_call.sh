mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r02p_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02p_gpu_tests.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02p_bench.json')); print(d['value'], d['e2e']['value'], {k:v for k,v in d['roofline']['by_kernel_ms'].items() if 'wmsa' in k})"
timeout 200 python bench.py --steps 10 --no-cpu-baseline --size 512 > gpurun_out/r02p_bench512.json 2> gpurun_out/r02p_bench512.err; echo "bench512 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02p_bench512.json')); print(d['metric'], d['value'], d['e2e']['value'], d['config']['workload'])"
