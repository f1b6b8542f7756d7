timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo pytest rc=$?; tail -8 gpurun_out/r2q_pytest.log
{
echo "== TMA_MC=1"; timeout 300 python tools/lml_probe.py 16384 3
echo "== TMA_MC=0"; MFGP_GEMM_TMA_MC=0 timeout 300 python tools/lml_probe.py 16384 3
echo "== TMA_MC=1 n=8192"; timeout 300 python tools/lml_probe.py 8192 2
echo "== TMA_MC=0 n=8192"; MFGP_GEMM_TMA_MC=0 timeout 300 python tools/lml_probe.py 8192 2
} > gpurun_out/r2q_mc.log 2>&1
grep -v "rep 0" gpurun_out/r2q_mc.log
