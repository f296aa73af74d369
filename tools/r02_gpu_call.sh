mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02l_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02l_gpu_tests.log
grep -E "rel_l2|residual stream|droppath|train grads" gpurun_out/r02l_gpu_tests.log | grep -E "uformer|restore|residual stream|droppath|train" | head -40
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02l_bench.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
