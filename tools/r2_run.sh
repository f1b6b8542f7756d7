python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2s_pytest.log
python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/r2s_smoke.log
