"""Write profiles/roofline_traffic.json (the per-kernel counters bench.py attaches to its roofline objects) from ncu
--set full captures:   python tools/ncu_roofline_json.py role=rep[:kernel-index] ... [--note TEXT]
e.g. forward=gpurun_out/r02_fwd.ncu-rep backward_in=gpurun_out/r02_bwd.ncu-rep:0 backward_rel=gpurun_out/r02_bwd.ncu-rep:1"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(rep, idx):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, r = rows[0], rows[2 + idx]
    val = lambda k: float(r[hdr.index(k)].replace(",", ""))  # noqa: E731
    unit = lambda k: rows[1][hdr.index(k)]  # noqa: E731

    def nbytes(k):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit(k)]
        return int(val(k) * scale)
    dur = val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1e-6)
    return {
        "kernel": r[hdr.index("Kernel Name")],
        "ncu_duration_ms": dur,
        "dram_read_bytes": nbytes("dram__bytes_read.sum"), "dram_write_bytes": nbytes("dram__bytes_write.sum"),
        "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
        "dram_read_pct_of_peak": val("dram__bytes_read.sum.pct_of_peak_sustained_elapsed"), "dram_write_pct_of_peak": val("dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
        "l1tex_lsu_frac": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
        "l2_frac": val("lts__throughput.avg.pct_of_peak_sustained_elapsed") / 100.0,
        "l1_hit_rate": val("l1tex__t_sector_hit_rate.pct") / 100.0, "l2_hit_rate": val("lts__t_sector_hit_rate.pct") / 100.0,
        "l2_to_l1_bytes": int(val("lts__t_sectors_srcunit_tex_op_read.sum") * 32),
        "occupancy": val("sm__warps_active.avg.pct_of_peak_sustained_active") / 100.0,
        "registers": int(val("launch__registers_per_thread")),
        "source": "ncu --set full --clock-control none, %s kernel #%d" % (os.path.basename(rep), idx),
    }


def main():
    data = {}
    out_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    args = []
    for arg in sys.argv[1:]:
        if arg.startswith("--out="):
            out_path = arg[len("--out="):]
        else:
            args.append(arg)
    for arg in args:
        role, spec = arg.split("=", 1)
        rep, _, idx = spec.partition(":")
        data[role] = read(rep, int(idx or 0))
    if "backward_in" in data and "backward_rel" in data:
        a, b = data["backward_in"], data["backward_rel"]
        data["backward"] = {"passes": ["backward_in", "backward_rel"],
                            "dram_bytes_per_launch": a["dram_bytes_per_launch"] + b["dram_bytes_per_launch"],
                            "ncu_duration_ms": a["ncu_duration_ms"] + b["ncu_duration_ms"],
                            "l1tex_lsu_frac": (a["l1tex_lsu_frac"] * a["ncu_duration_ms"] + b["l1tex_lsu_frac"] * b["ncu_duration_ms"])
                            / (a["ncu_duration_ms"] + b["ncu_duration_ms"]),
                            "l2_frac": (a["l2_frac"] * a["ncu_duration_ms"] + b["l2_frac"] * b["ncu_duration_ms"])
                            / (a["ncu_duration_ms"] + b["ncu_duration_ms"]),
                            "source": "time-weighted over the two passes"}
    with open(out_path, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(data, indent=1))


if __name__ == "__main__":
    main()
