"""ncu report -> profiles/r01_traffic.json (per-kernel DRAM bytes and duration; bench.py reads it for roofline.traffic)."""
import subprocess, csv, json, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); H, U = rows[0], rows[1]
res = {}
for r in rows[2:]:
    k = r[H.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
    if k in res:
        continue
    def val(n):
        v = float(r[H.index(n)].replace(",", "")); u = U[H.index(n)]
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    res[k] = {"dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6, "traffic_bytes": rd + wr, "ncu_duration_ms": val("gpu__time_duration.sum"),
              "pipe_fma_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
              "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active")}
json.dump({"source": "ncu --set full --clock-control none, one launch per kernel, C2 bs=4096 (profiles/r01_c2_ncu_summary.md)", "kernels": res},
          open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
