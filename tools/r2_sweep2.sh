# observe kernel schedule sweep: share of statically scheduled envs
tag=${1:-r2s2}
mkdir -p gpurun_out
for pct in 100 85 62 40 0; do
  SSD_OBS_STATIC_PCT=$pct timeout 300 python bench.py --steps 300 --warmup 50 --no-cpu --e2e-steps 100 > gpurun_out/${tag}_p$pct.json 2> gpurun_out/${tag}_p$pct.err
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${tag}_tests.log
