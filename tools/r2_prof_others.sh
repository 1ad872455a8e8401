# ncu --set full captures of the kernels of the other configs (steady state) + the masked reset kernel of the headline
tag=${1:-r2o}
mkdir -p gpurun_out
run() { # name regex config skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o gpurun_out/${tag}_$1 python bench.py --config $3 --steps 50 --warmup 5 --no-cpu --graph-steps 1 --e2e-steps 2 > gpurun_out/${tag}_$1.log 2>&1
}
run feat_step feat_step features1m 1030
run car_step car_step selfdrive8 430
run harvest_obs grid_obs harvest16k 1030
run harvest_logic grid_logic harvest16k 1030
run reset grid_reset cleanup8 1030
