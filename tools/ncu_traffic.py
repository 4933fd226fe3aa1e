"""DRAM traffic of the step kernels from an `ncu --set full` report -> profiles/traffic_r02.json entry.
usage: python tools/ncu_traffic.py report.ncu-rep game mode envs_per_launch "<command that produced the report>" """
import csv
import io
import json
import os
import subprocess
import sys

rep, game, mode, envs, cmd = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]


def val(r, key):
    v = float(r[hdr.index(key)].replace(",", ""))
    u = units[hdr.index(key)].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1)


path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic_r02.json")
db = json.load(open(path)) if os.path.exists(path) else {}
entry = {"source": f"ncu --set full --clock-control none, one serialised launch of {envs} envs: {cmd}; report {os.path.basename(rep)}", "kernels": {}}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    short = "render_kernel" if "render_kernel" in name else "setup_kernel" if "setup_kernel" in name else "logic_kernel" if "logic_kernel" in name else name[:30]
    entry["kernels"][short] = {"dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                               "duration_ms_under_ncu": val(r, "gpu__time_duration.sum"), "grid": int(float(r[hdr.index("launch__grid_size")]))}
k = entry["kernels"].get("render_kernel")
if k:
    entry["dram_bytes_per_launch"] = k["dram_read_bytes"] + k["dram_write_bytes"]
    entry["algorithmic_bytes_per_launch"] = 12288 * envs
db[f"{game}:{mode}:{envs}"] = entry
json.dump(db, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))
