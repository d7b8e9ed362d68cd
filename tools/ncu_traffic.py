"""Reads an `ncu --set full` report of `tools/ncu_step.py` (here, no GPU needed) and writes the per-launch DRAM traffic of
the MLP kernels of ONE training step to profiles/r2_traffic.json - the file bench.py's `roofline.traffic` comes from.

    python tools/ncu_traffic.py gpurun_out/r2_step.ncu-rep [profiles/r2_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r2_traffic.json"
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}


def num(r, name):
    v = r[col[name]].replace(",", "")
    u = units[col[name]]
    f = float(v)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1}.get(u, 1)
    return f * scale


launches = []
for r in body:
    launches.append({
        "kernel": r[col["Kernel Name"]].split("(")[0],
        "grid": r[col["Grid Size"]] if "Grid Size" in col else None,
        "ms_under_ncu": num(r, "gpu__time_duration.sum"),
        "dram_read_bytes": num(r, "dram__bytes_read.sum"),
        "dram_write_bytes": num(r, "dram__bytes_write.sum"),
        "tensor_pipe_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None,
        "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in col else None,
    })
fwd = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches if "mlp_fwd" in l["kernel"])
bwd = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in launches if "mlp_bwd" in l["kernel"])
res = {"src": "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of the MLP launches of one 4096-ray step, "
              "`ncu --set full --clock-control none -k regex:mlp_ -s 16 -c 8 python tools/ncu_step.py` on this tree (" + rep + ")",
       "fwd_train_bytes_per_step": fwd, "bwd_bytes_per_step": bwd, "launches": launches}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
