#!/bin/bash
# profiles/tools/ab.sh TAG [pytest -k expression] -- development run on the GPU box (through gpurun):
# parity tests of the current build, then the C3 bench line of the current build and of libaadp_r1.so
# (the library as round 1 left it, when present) back to back.
TAG=${1:-ab}; KEXPR=${2:-}
O=gpurun_out; mkdir -p $O
if [ -n "$KEXPR" ]; then python -m pytest tests -x -q -m gpu -k "$KEXPR" > $O/${TAG}_pytest.log 2>&1; else python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; fi
echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log
python bench.py --no-cpu-baseline --no-e2e-overlap > $O/${TAG}_c3.json 2>$O/${TAG}_c3.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("$O/${TAG}_c3.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), d.get("kernel_share"), "e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
if [ -f alignment_algos_b200/libaadp_r1.so ]; then
  AADP_LIB=alignment_algos_b200/libaadp_r1.so python bench.py --no-cpu-baseline --no-e2e-overlap > $O/${TAG}_c3_r1.json 2>$O/${TAG}_c3_r1.err
  python - <<PY
import json
d = json.loads(open("$O/${TAG}_c3_r1.json").read().strip().splitlines()[-1])
print("r1 lib: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), d.get("kernel_share"))
PY
fi
