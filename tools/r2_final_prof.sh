# round-2 evidence: launch list of the headline step + one ncu --set full capture per kernel of every config (steady state).
# $2 selects the half (the reports of one call have to stay below gpurun's 64 MiB return limit).
tag=${1:-r2z}
half=${2:-a}
mkdir -p gpurun_out
B="python bench.py --steps 50 --warmup 5 --no-cpu --graph-steps 1 --e2e-steps 2"
cap() { # name regex config skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o gpurun_out/${tag}_$1 $B --config $3 > gpurun_out/${tag}_$1.log 2>&1
}
if [ $half = a ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 60 --csv --log-file gpurun_out/${tag}_launches_cleanup8.csv $B --config cleanup8 > /dev/null 2>&1
  cap obs grid_obs cleanup8 1030
  cap logic grid_logic cleanup8 1030
  cap reset grid_reset cleanup8 1030
  cap harvest_obs grid_obs harvest16k 1030
  cap harvest_logic grid_logic harvest16k 1030
else
  cap harvest_reward grid_reward harvest16k 1030
  cap feat_cleanup feat_kernel features1m 2020
  cap feat_harvest feat_kernel harvestfeat1m 2020
  cap car car_kernel selfdrive8 430
fi
