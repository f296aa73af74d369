mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02n_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02n_gpu_tests.log
grep -E "rel_l2|residual stream|auto" gpurun_out/r02n_gpu_tests.log | grep -E "uformer_b_256:|uformer_b_256 |residual stream|auto|uformer_s2" | head -20
for r in auto fp32 bf16; do
timeout 200 python bench.py --steps 10 --no-cpu-baseline --residual $r > gpurun_out/r02n_bench_$r.json 2> gpurun_out/r02n_bench_$r.err; echo "bench $r rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_$r.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
done
