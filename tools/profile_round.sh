# round evidence: tests, smoke, bench (both arms), launch list, one ncu --set full capture per hot kernel
tag=${1:-r1_v6}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_tests.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 100 --warmup 10 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 100 --warmup 500 --no-cpu --e2e-steps 2 --graph-steps 1 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:grid_obs -s 990 -c 1 -o gpurun_out/${tag}_obs python bench.py --steps 100 --warmup 500 --no-cpu --e2e-steps 2 --graph-steps 1 > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:grid_logic -s 990 -c 1 -o gpurun_out/${tag}_logic python bench.py --steps 100 --warmup 500 --no-cpu --e2e-steps 2 --graph-steps 1 > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.log
timeout 600 python tools/bench_configs.py --steps 300 > gpurun_out/${tag}_other_configs.jsonl 2> gpurun_out/${tag}_other_configs.err
