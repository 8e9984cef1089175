# one --set full capture of a dense iteration launch (every level works on all 1 978 624 points)
set -e
timeout 300 python profiles/tools/run_dense.py > gpurun_out/dense_plain.log 2>&1
tail -1 gpurun_out/dense_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_iter_tma -s 40 -c 2 \
  -o gpurun_out/r01_dense_iter -f python profiles/tools/run_dense.py > gpurun_out/ncu_dense.log 2>&1
tail -3 gpurun_out/ncu_dense.log
