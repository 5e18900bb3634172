"""DRAM traffic of one assembly step from an `ncu --set full --page raw --csv` export of its kernels (tools/make_profiles.sh):
sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches and writes profiles/<tag>_step_traffic.json.
    python tools/ncu_step_traffic.py gpurun_out/r02_step_raw.csv profiles/r02_step_traffic.json 70"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
H, U, data = rows[0], rows[1], rows[2:]
def col(name): return H.index(name)
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot_r = tot_w = 0.0
kern = []
for r in data:
    rd = float(r[col("dram__bytes_read.sum")].replace(",", "")) * scale[U[col("dram__bytes_read.sum")]]
    wr = float(r[col("dram__bytes_write.sum")].replace(",", "")) * scale[U[col("dram__bytes_write.sum")]]
    t = float(r[col("gpu__time_duration.sum")].replace(",", ""))
    tot_r += rd; tot_w += wr
    kern.append({"kernel": r[col("Kernel Name")][:60], "grid": r[col("Grid Size")], "time": t, "time_unit": U[col("gpu__time_duration.sum")],
                 "dram_read_bytes": rd, "dram_write_bytes": wr})
out = {"M": int(sys.argv[3]), "dram_bytes_per_step": tot_r + tot_w, "dram_read_bytes": tot_r, "dram_write_bytes": tot_w,
       "launches": len(kern), "kernels": kern, "source": sys.argv[1]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "kernels"}))
