timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "input_proj or output_proj or model_golden or graphed or downsample" 2>&1 | grep -E "input_proj|passed|failed|Error|error" | tail -12
timeout 100 python tools/outproj_probe.py 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/c6_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value']); r=d['roofline']['by_kernel_ms']; print({k:r[k] for k in r if k.startswith('down') or k.startswith('up') or 'proj' in k or 'T8192' in k})"
