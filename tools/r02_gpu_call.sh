mkdir -p gpurun_out
timeout 600 python tools/leff_fused_probe.py > gpurun_out/r02u_leff_probe.log 2>&1; echo "probe rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02u_leff_probe.log'):
    if ln.startswith('RESULT'):
        d=json.loads(ln[7:]); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k in ('C','H','B','rel_l2_vs_oracle','fused_us','split_us','strided_out_equal')})
    else: print(ln.strip()[:300])
PY
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02u_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02u_gpu_tests.log
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02u_bench.json')); print(d['value'], d['e2e']['value'])"
