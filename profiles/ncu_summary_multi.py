"""Per-kernel table from an `ncu --set full` report exported with `--page raw --csv`: one line per captured launch.
usage: python profiles/ncu_summary_multi.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default


print(f"{'kernel':44s} {'us':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'GB/s':>7s} {'dram%':>6s} {'sm%':>6s} {'occ%':>6s} {'regs':>5s} {'issue%':>6s}")
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:44]
    t = f(r, "gpu__time_duration.sum")
    unit = rows[1][col["gpu__time_duration.sum"]]
    t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    ur, uw = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(ur, 1.0); wr *= scale.get(uw, 1.0)
    print(f"{name:44s} {t_us:8.1f} {rd / 1e6:10.1f} {wr / 1e6:10.1f} {(rd + wr) / t_us / 1e3:7.0f} "
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {f(r, 'launch__registers_per_thread'):5.0f} "
          f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f}")
