"""profiles/tools/ncu_counts.py REP.ncu-rep --pairs N [--tag r02x] -- per-kernel counts of one `ncu --set full` capture of
`bench.py --pairs N --steps 1 ...` (the C3 stream, seed 1003), normalised per cell update, written to
profiles/r02_kernel_counts.json.  bench.py reads that file for `roofline.traffic` and `roofline.i_cell_sass`, so both are
measurements of the build that was profiled, scaled to the launch bench.py times.

  dram_bytes_per_cell_update   (dram__bytes_read.sum + dram__bytes_write.sum) / cell updates of the launch
  lane_ops_per_cell_update     smsp__inst_executed.sum * 32 / cell updates   (issue slots in lane units; every packed
                               instruction updates two cells)
Runs here (no GPU): it only reads the report."""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--pairs", type=int, required=True)
    ap.add_argument("--tag", default="r02")
    args = ap.parse_args()
    from alignment_algos_b200 import synth
    seqs, pq, pt = synth.pair_workload(1003, args.pairs, 100, 500)
    cu = float(sum(len(seqs[a]) * len(seqs[b]) for a, b in zip(pq, pt)))  # per direction
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"source": "%s (ncu --set full of bench.py --pairs %d, %s; summary committed as profiles/%s_ncu_full_packed_kernels.txt)"
                     % (os.path.basename(args.rep), args.pairs, args.tag, args.tag),
           "cell_updates_per_launch": cu, "kernels": {}}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"packed_kernel<\(?(?:int\))?(\d), \(?(?:int\))?(\d), \(?(?:int\))?(\d), \(?(?:int\))?(\d)", name) or \
            re.search(r"packed_kernel<(\d), (\d), (\d), (\d)>", name)
        if not m:
            continue
        tb, fst, msk, xm = (int(x) for x in m.groups())
        if xm:
            continue
        direction = "rev" if msk else "fwd"
        key = "packed_kernel<TB=%d,FST=%d,MSK=%d>%s" % (tb, fst, msk, direction)

        def f(metric):
            return float(r[col[metric]]) if metric in col and r[col[metric]] else float("nan")

        unit = {h: rows[1][i] for h, i in col.items()}
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        dram = sum(f(mn) * scale.get(unit[mn], 1.0) for mn in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        inst = f("smsp__inst_executed.sum")
        out["kernels"][key] = {
            "dram_bytes_per_cell_update": dram / cu,
            "lane_ops_per_cell_update": inst * 32.0 / cu,
            "warp_instructions": inst,
            "dram_bytes": dram,
            "duration_ms_under_ncu": f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit["gpu__time_duration.sum"], 1.0),
            "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "l1tex_lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": f("launch__registers_per_thread"),
            "warps_per_sm": f("sm__warps_active.avg.per_cycle_active"),
        }
    path = os.path.join(ROOT, "profiles", "r02_kernel_counts.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
