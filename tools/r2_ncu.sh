# round-2 ncu passes (one gpurun call): launch list of one LML+gradient evaluation, --set full captures of the dominant kernels
set -x
python tools/lml_probe.py 16384 1 > gpurun_out/r2k_plain_lml.log 2>&1 && \
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2k_launches_lml.csv python tools/lml_probe.py 16384 1 > gpurun_out/r2k_ncu2.log 2>&1
S="python bench.py --steps 1 --warmup 1 --points 65536 --no-lml --no-cpu --no-extras"
$S > gpurun_out/r2k_plain_s.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"trmm_sumsq|cross_gen_mc" --launch-skip 8 -c 4 -o gpurun_out/r2k_mc_hf $S > gpurun_out/r2k_ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma_kernel" --launch-skip 60 -c 2 -o gpurun_out/r2k_gemm_tma python tools/lml_probe.py 16384 1 > gpurun_out/r2k_ncu4.log 2>&1
ls -la gpurun_out/r2k*
