# end-of-round verification: GPU tests, smoke, bench, then one ncu pass (launch list of the dense workload + one full
# capture of the dense iteration kernel)
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense_iter_tma -s 40 -c 1 \
  -o gpurun_out/r01_dense_iter2 -f python profiles/tools/run_dense.py > gpurun_out/ncu_dense2.log 2>&1
tail -2 gpurun_out/ncu_dense2.log
