# A/B of one env switch ($2=VAR) on the headline bench + the gpu test-suite
tag=${1:-r2ab}; var=${2:-SSD_NO_PDL}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
for v in 0 1 0 1; do
  env $var=$v timeout 300 python bench.py --steps 500 --warmup 50 --no-cpu --e2e-steps 100 > gpurun_out/${tag}_${var}_$v.json 2> gpurun_out/${tag}_${var}_$v.err
  cat gpurun_out/${tag}_${var}_$v.json >> gpurun_out/${tag}_all.jsonl
done
