# quick GPU regression: debug rollouts vs oracle, the gpu test-suite, a short bench; everything under `timeout`
tag=${1:-x}
timeout 100 python tools/dbg_step.py cleanup 200 40 > gpurun_out/${tag}_d1.log 2>&1; echo rc=$? >> gpurun_out/${tag}_d1.log
timeout 100 python tools/dbg_step.py harvest 200 40 4 > gpurun_out/${tag}_d3.log 2>&1; echo rc=$? >> gpurun_out/${tag}_d3.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
timeout 250 python bench.py --steps 300 --warmup 300 --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 1500 -c 6 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 100 --warmup 500 --no-cpu --e2e-steps 2 --graph-steps 1 > gpurun_out/${tag}_l.log 2>&1
