# round-2 ncu passes for the kernels added late in the round (one gpurun call)
set -x
P="python tools/mc_small_probe.py 30 100 131072 100"
$P > gpurun_out/r2u_plain_small.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mc_small" -c 2 -o gpurun_out/r2u_mc_small $P > gpurun_out/r2u_ncu1.log 2>&1
L="python tools/lml_probe.py 8192 1"
$L > gpurun_out/r2u_plain_lml.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tma_mc" -c 14 -o gpurun_out/r2u_gemm_tma_mc $L > gpurun_out/r2u_ncu2.log 2>&1
ls -la gpurun_out/r2u*
