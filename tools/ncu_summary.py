"""ncu_summary.py <file.ncu-rep> <out.csv> [traffic.json]: the handful of metrics DESIGN.md / profiles/README.md quote, one
column per kernel, from `ncu -i <rep> --page raw --csv`; optionally dram read+write bytes per launch as JSON (bench.py reads
profiles/ncu_traffic.json for roofline.traffic)."""
import csv, io, json, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
head, units, data = rows[0], rows[1], rows[2:]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_op_red.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
kcol = head.index("Kernel Name")
names = [re.sub(r"\(.*", "", r[kcol]).replace("aki::", "").replace("void ", "") for r in data]
with open(out, "w") as f:
    f.write("metric,unit," + ",".join(names) + "\n")
    for m in WANT:
        if m in head:
            c = head.index(m)
            f.write(f"{m},{units[c]}," + ",".join(r[c].replace(",", "") for r in data) + "\n")
if len(sys.argv) > 3:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cr, cw = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
    tr = {re.sub(r"<.*>", "", n): float(r[cr].replace(",", "")) * scale[units[cr]] + float(r[cw].replace(",", "")) * scale[units[cw]] for n, r in zip(names, data)}
    json.dump(tr, open(sys.argv[3], "w"), indent=1)
print(open(out).read())
