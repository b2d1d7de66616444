"""Per-launch table of the convolution kernels from a raw-page CSV (`ncu -i rep --page raw --csv`), run here (no GPU).

    python scripts/ncu_conv_table.py gpurun_out/prof_conv_fp32_raw.csv profiles/r2/conv_kernels_fp32_ncu.md

Next to the table it writes `<dst stem>.json`: per kernel family the launches, total time and the time-weighted tensor-pipe
activity, which `bench.py` reports as `roofline.tensor_pipe_pct_by_kernel`.
"""
import csv
import json
import sys
from pathlib import Path

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor(all) %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc smem wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lsu smem wavefronts %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__grid_size", "grid"),
    ("launch__shared_mem_per_block_dynamic", "smem/blk"),
    ("launch__registers_per_thread", "regs"),
]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(hdr):
        for w, _ in WANT:
            if h == w or h.endswith("." + w):
                col.setdefault(w, i)
    name_i = hdr.index("Kernel Name")
    fam = {}
    with open(dst, "w") as f:
        f.write(f"# convolution kernels, one row per launch of one step (`ncu --set full --clock-control none`, source {src})\n\n")
        f.write("| # | kernel | " + " | ".join(lbl for w, lbl in WANT if w in col) + " |\n")
        f.write("|---|---|" + "---:|" * sum(1 for w, _ in WANT if w in col) + "\n")
        for k, r in enumerate(rows[2:]):
            if len(r) != len(hdr):
                continue
            name = r[name_i].split("(")[0].replace("void ", "").replace("fsr::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
            vals = []
            for w, _ in WANT:
                if w not in col:
                    continue
                v = r[col[w]]
                if w == "gpu__time_duration.sum":
                    x = float(v.replace(",", ""))
                    x *= {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[col[w]], 1.0)
                    v = f"{x:.1f}"
                else:
                    try:
                        v = f"{float(v.replace(',', '')):.1f}"
                    except ValueError:
                        pass
                vals.append(v)
            f.write(f"| {k} | `{name}` | " + " | ".join(vals) + " |\n")
            try:
                us = float(vals[0])
                tp = float(r[col["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed"]].replace(",", ""))
            except (KeyError, ValueError):
                continue
            a = fam.setdefault(name.split("<")[0], {"launches": 0, "us": 0.0, "w": 0.0, "min": 1e9, "max": 0.0})
            a["launches"] += 1
            a["us"] += us
            a["w"] += us * tp
            a["min"] = min(a["min"], tp)
            a["max"] = max(a["max"], tp)
    out = {"source": src, "metric": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "kernels": {
        k: {"launches": a["launches"], "us_per_step_under_ncu": round(a["us"], 1), "tensor_pipe_pct_time_weighted": round(a["w"] / a["us"], 1),
            "tensor_pipe_pct_min": round(a["min"], 1), "tensor_pipe_pct_max": round(a["max"], 1)} for k, a in fam.items() if a["us"] > 0}}
    Path(dst).with_suffix(".json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", dst, "and", Path(dst).with_suffix(".json"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
