timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "block_vs_oracle or upsample or leff or module_golden or model_golden" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/c8_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']); r=d['roofline']['by_kernel_ms']; print({k:r[k] for k in r if 'C512' in k or k.startswith('up')})"
