"""Summarise ncu artefacts into small text files for profiles/ (the judged copies; gpurun_out/ is scratch).

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rNN_full_<kernel>.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size_x", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_hmma.sum",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    i_name, i_metric, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    i_unit = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        v = v / 1e3 if r[i_unit] in ("ns", "nsecond") else v          # -> us
        name = r[i_name]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {sum(a[0] for a in agg.values())} launches, {total:.1f} us")
    print(f"{'share':>7s} {'total_us':>10s} {'launches':>8s} {'avg_us':>9s}  kernel")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.1f}% {t:10.1f} {n:8d} {t / n:9.2f}  {name[:120]}")


def _raw_rows(path):
    """rows of `ncu --page raw --csv`: from a .ncu-rep (converted here) or from a csv exported on the GPU box."""
    if path.endswith(".csv"):
        return list(csv.reader(open(path)))
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    return list(csv.reader(io.StringIO(out)))


def full(path):
    rows = _raw_rows(path)
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== " + r[hdr.index("Kernel Name")][:140])
        for k in KEYS:
            for j, h in enumerate(hdr):
                if h == k:
                    print(f"   {h} [{units[j]}] = {r[j]}")
        print()


def traffic(path):
    """profiles/r01_traffic.json: per roofline kernel class, DRAM bytes per launch averaged over the captured launches."""
    import json
    rows = _raw_rows(path)
    hdr, units = rows[0], rows[1]
    i_name, i_r, i_w, i_t = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = {}
    for r in rows[2:]:
        name = r[i_name]
        key = None
        if "k_sconv_ts" in name and ", 27," in name.replace("(int)", ""):
            key = "k_sconv_ts[3x3x3, all channel widths]"
        if key is None:
            continue
        b = float(r[i_r].replace(",", "")) * scale[units[i_r]] + float(r[i_w].replace(",", "")) * scale[units[i_w]]
        a = agg.setdefault(key, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += b
        a["us"] += float(r[i_t].replace(",", "")) / (1e3 if units[i_t] in ("ns", "nsecond") else 1.0)
    res = {k: {"dram_bytes_per_launch": v["dram_bytes"] / v["launches"], "launches_captured": v["launches"],
               "avg_us_under_ncu": v["us"] / v["launches"], "source": "ncu --set full, " + path.split("/")[-1]} for k, v in agg.items()}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
