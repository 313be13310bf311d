#!/bin/bash
# profiles/tools/variants.sh TAG "name:lib:pad" ... -- C3 bench line per library variant (AADP_LIB) and shared-memory
# pad (AADP_PACKED_SMEM_PAD = "fwd,rev" bytes, lowers warps/SM), all in one gpurun call so they share a box.
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for spec in "$@"; do
  IFS=: read name lib pad <<< "$spec"
  if [ -n "$pad" ]; then export AADP_PACKED_SMEM_PAD=$pad; else unset AADP_PACKED_SMEM_PAD; fi
  AADP_LIB=alignment_algos_b200/$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-overlap > $O/${TAG}_${name}.json 2>$O/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_${name}.json").read().strip().splitlines()[-1])
    ks = d.get("kernel_share", {})
    ms = d["ms_per_step"]
    print("%-22s GCUPS %7.1f  step %6.3f ms  " % ("$name", d["value"], ms) + "  ".join("%s %.3f ms" % (k.split(">")[-1], v * ms) for k, v in ks.items()))
except Exception as e:
    print("$name unreadable:", e)
PY
done
