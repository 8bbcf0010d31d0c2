"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into the small CSV/JSON files kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/final_*.ncu-rep > profiles/r01_ncu_final_summary.json"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
    "smsp__inst_executed.sum": "warp_insts",
}
UNIT = {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"report": path.split("/")[-1], "kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        parts = d["report"].rsplit(".", 1)[0].split("__")     # <round>__<bench kernel>__<bench shape key>.ncu-rep
        if len(parts) == 3:
            d["bench_key"] = [parts[1], parts[2]]             # what bench.py's `ncu_traffic` looks up
        for k, name in KEYS.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT.get(units[i], 1.0)
                d[name + ("_us" if name == "duration" else ("_bytes" if name.startswith("dram_r") or name.startswith("dram_w") else ""))] = v
        if "dram_read_bytes" in d:
            d["dram_traffic_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
            d["dram_gbs"] = d["dram_traffic_bytes"] / d["duration_us"] / 1e3
        out.append(d)
    return out


if __name__ == "__main__":
    res = []
    for p in sys.argv[1:]:
        res += summarise(p)
    print(json.dumps(res, indent=1))
