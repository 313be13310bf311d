#!/bin/bash
# profiles/tools/wave_check.sh TAG -- long-pair wavefront: parity tests under a watchdog, then the C5 bench line
TAG=${1:-w}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_pair or wave" > $O/${TAG}_wave_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/${TAG}_wave_pytest.log | cut -c1-300
AADP_WAVE_DEBUG=1 timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_c5.json 2> $O/${TAG}_c5.err; echo "c5 rc=$?"
tail -4 $O/${TAG}_c5.err | cut -c1-1200
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_c5.json").read().strip().splitlines()[-1])
    print("C5 GCUPS", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), d.get("kernel_share"))
except Exception as e:
    print("c5 unreadable", e)
PY
