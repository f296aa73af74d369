timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "leff or block_vs_oracle or module_golden or model_golden or graphed or concat" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/c12_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']); r=d['roofline']['by_kernel_ms']; print({k:r[k] for k in r if 'leff_C' in k})"
timeout 200 python bench.py --residual bf16 --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('bf16', d['value'],d['ms_per_step'])"
