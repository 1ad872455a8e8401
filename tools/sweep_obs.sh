# observe-kernel launch shape sweep: rebuild with -DOBS_WARPS / -DOBS_MAXREG, short bench each (run on the GPU box)
for cfg in "8 80" "9 72" "5 80" "13 72" "9 64" "10 64"; do set -- $cfg
  python -c "
from contracts_b200 import build
build.build(force=True, extra_flags=['-DOBS_WARPS=$1', '-DOBS_MAXREG=$2'])" > /dev/null 2>&1 || { echo "warps $1 maxreg $2: build failed"; continue; }
  SSD_DEBUG=1 timeout 200 python bench.py --steps 200 --warmup 300 --no-cpu --e2e-steps 2 2> gpurun_out/sweep_$1_$2.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('warps $1 maxreg $2', '%.3e'%d['value'], d['roofline']['kernel_ms'], d['roofline']['kernels'])"
  grep -m1 "observe kernel" gpurun_out/sweep_$1_$2.err
done
python -c "
from contracts_b200 import build
build.build(force=True)"
