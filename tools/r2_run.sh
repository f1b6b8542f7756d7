python -m pytest tests/test_gpu_models.py tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2n_pytest.log
{
MFGP_MC_SMALL=0 MFGP_PROBE_SAVE=/tmp/gen30.npy python tools/mc_small_probe.py 30 100
MFGP_TRACE_MC=1 MFGP_PROBE_COMPARE=/tmp/gen30.npy python tools/mc_small_probe.py 30 100
MFGP_MC_SMALL=0 MFGP_PROBE_SAVE=/tmp/gen60.npy python tools/mc_small_probe.py 60 100
MFGP_TRACE_MC=1 MFGP_PROBE_COMPARE=/tmp/gen60.npy python tools/mc_small_probe.py 60 100
MFGP_TRACE_MC=1 python tools/mc_small_probe.py 10 50
} > gpurun_out/r2n_small.log 2>&1
grep -v "rep 0" gpurun_out/r2n_small.log
