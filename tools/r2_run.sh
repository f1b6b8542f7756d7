timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2r_pytest.log
timeout 300 python tools/lml_total_probe.py 16384 2>&1 | tail -6
timeout 300 python tools/lml_total_probe.py 8192 2>&1 | tail -6
timeout 300 python tools/lml_total_probe.py 2500 2>&1 | tail -6
