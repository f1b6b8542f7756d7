python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2o_pytest.log
python bench.py > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_bench_reference.json 2> gpurun_out/r2o_bench_reference.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/r2o_smoke.log
