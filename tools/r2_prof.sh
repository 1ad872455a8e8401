# one ncu --set full capture per hot kernel of the headline step (steady state: after the burn-in), plus a launch list
tag=${1:-r2p}
cfg=${2:-cleanup8}
mkdir -p gpurun_out
B="python bench.py --config $cfg --steps 50 --warmup 5 --no-cpu --graph-steps 1 --e2e-steps 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_obs -s 1030 -c 1 -f -o gpurun_out/${tag}_obs $B > gpurun_out/${tag}_obs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_logic -s 1030 -c 1 -f -o gpurun_out/${tag}_logic $B > gpurun_out/${tag}_logic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $B > /dev/null 2>&1
