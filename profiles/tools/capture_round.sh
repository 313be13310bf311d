#!/bin/bash
# profiles/tools/capture_round.sh TAG -- the evidence run of a session, executed on the GPU box through gpurun:
#   gpurun --timeout 2400 -- 'bash profiles/tools/capture_round.sh r01d'
# GPU tests, bench lines of every workload (own arm + reference arm), then -- each only after the same command exited
# 0 without ncu -- the ncu launch list of the default bench and one `--set full` capture of the packed kernels.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
if [ -z "$SKIP_PYTEST" ]; then python -m pytest tests -x -q -m gpu --durations=15 > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest_gpu.log; fi
python bench.py > $O/${TAG}_bench_default.json 2>$O/${TAG}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>$O/${TAG}_bench_reference.err; echo "ref rc=$?"
for w in c2 c4 c5 c3f; do
  python bench.py --workload $w --no-cpu-baseline > $O/${TAG}_bench_${w}_1gpu.json 2>$O/${TAG}_bench_${w}.err; echo "$w rc=$?"
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/${TAG}_b.json 2>$O/${TAG}_b.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${TAG}.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/${TAG}_ncu_l.log 2>&1
python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/${TAG}_b20k.json 2>$O/${TAG}_b.err && \
  ncu --set full --clock-control none --import-source on -k regex:packed_kernel --launch-skip 6 --launch-count 2 \
      -o $O/prof_${TAG} -f python bench.py --pairs 20000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/${TAG}_ncu_f.log 2>&1
for f in $O/${TAG}_bench_default.json $O/${TAG}_bench_reference.json $O/${TAG}_bench_c2_1gpu.json $O/${TAG}_bench_c4_1gpu.json $O/${TAG}_bench_c5_1gpu.json; do
  tail -1 $f | cut -c1-400
done
