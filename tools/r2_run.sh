python -m pytest tests -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2l_pytest.log
export MFGP_PROBE_CHECKSUM=1
{
for n in 16384 8192 32768; do
echo "== n=$n overlap=1"; python tools/lml_probe.py $n 3
echo "== n=$n overlap=0"; MFGP_OVERLAP_TRTRI=0 python tools/lml_probe.py $n 3
done
} > gpurun_out/r2l_overlap.log 2>&1
grep -v "rep 0" gpurun_out/r2l_overlap.log
