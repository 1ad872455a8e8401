# compute-sanitizer over the eight-lanes-per-env kernels: the tools named after the tag (default all three), each under a timeout
tag=${1:-r2san}; shift
mkdir -p gpurun_out
for tool in ${@:-racecheck synccheck memcheck}; do
  SAN_STEPS=12 timeout 280 compute-sanitizer --tool $tool --print-limit 40 python tools/sanitize_octets.py > gpurun_out/${tag}_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/${tag}_$tool.log
done
