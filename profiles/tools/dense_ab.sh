# A/B of the dense iteration launch: programmatic dependent launch (default) against plain stream order
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "dense_alignment or multi_cta" 2>&1 | tail -2
for i in 1 2; do
echo "== PDL";     timeout 120 python profiles/tools/run_dense.py | tail -1
echo "== no PDL";  ICT_DENSE_NOPDL=1 timeout 120 python profiles/tools/run_dense.py | tail -1
done
echo "== PDL, 3 stages";     ICT_DENSE_NST=3 timeout 120 python profiles/tools/run_dense.py | tail -1
