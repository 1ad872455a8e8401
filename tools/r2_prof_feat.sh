# ncu --set full captures of the feature-env step kernel in steady state (cleanup and harvest)
tag=${1:-r2pf}
mkdir -p gpurun_out
run() { # name config skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:feat_kernel -s $3 -c 1 -f -o gpurun_out/${tag}_$1 python bench.py --config $2 --steps 50 --warmup 5 --no-cpu --graph-steps 1 --e2e-steps 2 > gpurun_out/${tag}_$1.log 2>&1
}
run feat_cleanup features1m 2020
run feat_harvest harvestfeat1m 2020
