"""Per-launch table from an `ncu -i rep --page raw --csv` export: duration, tensor-pipe utilisation, DRAM / L2 / L2->SM bytes.
usage: python tools/ncu_tensor_table.py raw.csv [first_launch [count]]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = data[lo:lo + int(sys.argv[3])] if len(sys.argv) > 3 else data[lo:]
SC = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
def val(r, name):
    if name not in col or not r[col[name]]:
        return float("nan")
    return float(r[col[name]].replace(",", "")) * SC.get(units[col[name]], 1.0)
print(f"{'grid':>16} {'cluster':>8} {'us':>9} {'tc util %':>9} {'dram MB':>9} {'L2 MB':>9} {'L2->SM MB':>10} {'sm thr %':>8}")
tot = 0.0
for r in data:
    us = val(r, "gpu__time_duration.sum"); tot += us
    cl = r[col["launch__cluster_size"]] if "launch__cluster_size" in col else ""
    print(f"{r[col['Grid Size']]:>16} {cl:>8} {us:9.1f} {val(r, 'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active'):9.1f} "
          f"{val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum'):9.1f} {val(r, 'lts__t_bytes.sum'):9.1f} "
          f"{val(r, 'l1tex__m_xbar2l1tex_read_bytes_pipe_tma.sum'):10.1f} {val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):8.1f}")
print(f"sum {tot / 1e3:.1f} ms")
