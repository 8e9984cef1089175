for d in 0 2 4; do echo "== dbg $d"; ICT_DBG_SKIP_SERIAL=$d timeout 120 python profiles/tools/run_dense.py | tail -1; done
ICT_DBG_SKIP_SERIAL=2 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none --csv --log-file gpurun_out/dense_dbg2_launches.csv python profiles/tools/run_dense.py > /dev/null 2>&1
