timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/c3_tests.log; cat gpurun_out/c3_tests.log
for pdl in 1 0 1 0; do
UFORMER_B200_PDL=$pdl timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c3_bench_pdl$pdl.json 2> gpurun_out/c3_bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/c3_bench_pdl$pdl.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']); r=d['roofline']['by_kernel_ms']; print({k:r[k] for k in r if k.startswith('down') or k.startswith('up') or 'proj' in k})"
done
