# logic-kernel launch shape sweep (run on the GPU box)
for cfg in "128 7" "64 14" "128 8" "64 16" "96 9"; do set -- $cfg
  python -c "
from contracts_b200 import build
build.build(force=True, extra_flags=['-DLOGIC_THREADS=$1', '-DLOGIC_MIN_BLOCKS=$2'])" > /dev/null 2>&1 || { echo "threads $1 min blocks $2: build failed"; continue; }
  cuobjdump -res-usage contracts_b200/libssd_b200.so 2>/dev/null | grep -A1 "grid_logic_kernelILi0" | grep -oE "REG:[0-9]+ STACK:[0-9]+"
  timeout 200 python bench.py --steps 200 --warmup 300 --no-cpu --e2e-steps 2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads $1 min blocks $2', d['roofline']['kernel_ms'], [round(k['ms'],4) for k in d['roofline']['kernels']])"
done
python -c "
from contracts_b200 import build
build.build(force=True)"
