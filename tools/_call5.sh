timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wmsa_tma or block_vs or module_golden" > gpurun_out/r02z5_tests.log 2>&1; tail -5 gpurun_out/r02z5_tests.log
PROBE_MODULATOR=1 timeout 600 python tools/wmsa_tma_probe.py > gpurun_out/r02z5_probe_mod.log 2>&1
