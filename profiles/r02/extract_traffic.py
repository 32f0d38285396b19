"""Reads the ncu --set full captures of profiles/r02/run_ncu.sh (gpurun_out/r02/prof_<workload>.ncu-rep) and writes
profiles/r02/kernel_traffic.json -- the per-launch DRAM traffic bench.py reports as roofline.traffic -- plus the raw metric
export and a SASS excerpt of the dominant kernel per workload.  Run in the build container: python profiles/r02/extract_traffic.py"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

OUT = os.path.join(ROOT, "profiles", "r02")
WANT = {"gpu__time_duration.sum": "gpu_time", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "dram__sectors_read.sum": "dram_sectors_read", "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct", "smsp__inst_executed.sum": "warp_instructions",
        "launch__registers_per_thread": "registers", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "lts__t_sectors_srcunit_tex_op_read.sum": "l2_sectors_read",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def main():
    out = {}
    for name, wl in bench.WORKLOADS.items():
        rep = os.path.join(ROOT, "gpurun_out", "r02", "prof_%s.ncu-rep" % name)
        if not os.path.exists(rep):
            continue
        head, units, rows = raw_rows(rep)
        open(os.path.join(OUT, "ncu_full_raw_%s.csv" % name), "w").write(
            subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
        row = rows[0]
        rec = {"kernel": row[head.index("Kernel Name")]}
        for m, key in WANT.items():
            if m in head:
                i = head.index(m)
                v = float(row[i].replace(",", ""))
                rec[key] = v * UNIT.get(units[i], 1.0)
        R = wl["reads_per_step"]
        kmers = R * (wl["read_len"] - bench.K + 1)
        rec["reads_per_launch"] = R
        rec["kmers_per_launch"] = kmers
        rec["dram_bytes_per_launch"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
        rec["dram_bytes_per_kmer"] = rec["dram_bytes_per_launch"] / kmers
        # DRAM lines = 128-byte rows touched; ncu counts 32-byte sectors, the L2 fetches what missed (fetch granularity 32 B here)
        rec["dram_lines_per_kmer"] = rec["dram_bytes_read"] / 128.0 / kmers
        rec["warp_instructions_per_32_kmers"] = rec.get("warp_instructions", 0) / (kmers / 32.0)
        rec["source"] = "ncu --set full --clock-control none of `python bench.py --workload %s --also none --steps 1 --warmup 3` (profiles/r02/run_ncu.sh), launch 4 of the kernel; raw export profiles/r02/ncu_full_raw_%s.csv" % (name, name)
        out[name] = rec
        # SASS of the kernel: the instruction mix that proves the wide loads / reductions
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        keep = [l for l in src.splitlines() if any(t in l for t in ("LDG", "REDG", "RED.", "ATOM", "STG", "LDS", "STS", "SHFL", "Kernel", "Source"))]
        open(os.path.join(OUT, "sass_memory_ops_%s.csv" % name), "w").write("\n".join(keep[:400]) + "\n")
    json.dump(out, open(os.path.join(OUT, "kernel_traffic.json"), "w"), indent=1)
    for k, v in out.items():
        print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != "source"})


if __name__ == "__main__":
    main()
