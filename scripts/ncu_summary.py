"""Summarise ncu captures into small tracked files under profiles/ (run here, no GPU needed).

    python scripts/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/launches_r1.md
    python scripts/ncu_summary.py kernel gpurun_out/prof_head_r1.ncu-rep profiles/head_kernel_ncu.json --tiles 32 --flops-per-tile 4.998e9
    (--index k picks the k-th captured launch of a multi-kernel report)
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    seq = []
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("fsr::<unnamed>::", "").replace("<unnamed>::", "")
        ns = float(r[ix["Metric Value"]].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
        seq.append((name, r[ix["Grid Size"]], ns))
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration per launch, --clock-control none (cold-cache, serialised: compare shares)\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % |\n")
        f.write(f"\ntotal {total / 1e6:.3f} ms over {len(seq)} launches\n\n## sequence\n\n```\n")
        for name, grid, ns in seq:
            f.write(f"{name:42s} {grid:16s} {ns / 1e3:9.1f} us\n")
        f.write("```\n")
    print("wrote", dst)


def kernel(src, dst, tiles, flops_per_tile, index=0):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + index]  # one row per captured launch
    want = [
        "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__sass_inst_executed_op_utcmma.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_a.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utcmma_matrix_b_scope_1cta.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    ]
    rec = OrderedDict()
    for h, u, v in zip(hdr, units, vals):
        for w in want:
            if h == w or h.endswith("." + w):
                rec[w] = {"value": v, "unit": u}
    def num(key):
        return float(rec[key]["value"].replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    rd = num("dram__bytes_read.sum") * scale[rec["dram__bytes_read.sum"]["unit"]]
    wr = num("dram__bytes_write.sum") * scale[rec["dram__bytes_write.sum"]["unit"]]
    us = num("gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[rec["gpu__time_duration.sum"]["unit"]]
    tensor_pct, tensor_metric = None, None
    for key in ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"):
        try:
            tensor_pct, tensor_metric = num(key), key
            break
        except (KeyError, ValueError):
            continue
    summary = {
        "source": src,
        "tensor_pipe_pct": tensor_pct,
        "tensor_pipe_metric": tensor_metric,
        "tiles_in_launch": tiles,
        "flops_per_launch": tiles * flops_per_tile,
        "dram_bytes_read": rd,
        "dram_bytes_write": wr,
        "duration_us_under_ncu": us,
        "tflops_under_ncu": tiles * flops_per_tile / (us * 1e-6) / 1e12,
        "metrics": rec,
    }
    json.dump(summary, open(dst, "w"), indent=1)
    print("wrote", dst, "traffic MB", (rd + wr) / 1e6, "us", us)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        tiles = int(sys.argv[sys.argv.index("--tiles") + 1])
        fpt = float(sys.argv[sys.argv.index("--flops-per-tile") + 1])
        idx = int(sys.argv[sys.argv.index("--index") + 1]) if "--index" in sys.argv else 0
        kernel(sys.argv[2], sys.argv[3], tiles, fpt, idx)
