timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "point_service or direct_maximiser" > gpurun_out/r2p_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/r2p_pytest.log
timeout 300 python tools/direct_probe.py > gpurun_out/r2p_direct.log 2>&1; echo probe rc=$?; tail -5 gpurun_out/r2p_direct.log
