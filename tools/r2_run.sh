timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "mc" > gpurun_out/r2t_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2t_pytest.log
for v in default sb1 sb4; do
  if [ $v = default ]; then unset MFGP_LIB; else export MFGP_LIB=tools/variants/libmfgp_$v.so; fi
  python bench.py --steps 2 --warmup 1 --no-extras --no-lml --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value %.4e'%d['value'], 'ms %.1f'%d['ms_per_step'], d['roofline']['other_kernels_ms_per_launch'], 'pce %.15g'%d['pce_mean'])"
done
