"""Picks the headline metrics out of an `ncu -i rep --page raw --csv` dump (first captured kernel) and prints JSON.
usage: python tools/ncu_raw_pick.py gpurun_out/prof_X_raw.csv [key=value ...]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
d = dict(zip(h, rows[2])) if len(rows) > 2 else {}
num = lambda k: float(d[k].replace(",", "")) if k in d and d[k] not in ("", "n/a") else None  # noqa: E731
units = dict(zip(h, rows[1])) if len(rows) > 1 else {}


def to_bytes(k):
    v = num(k)
    if v is None:
        return None
    u = units.get(k, "")
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


out = {
    "kernel": d.get("Kernel Name"),
    "grid": d.get("Grid Size"), "block": d.get("Block Size"),
    "duration_us": (num("gpu__time_duration.sum") or 0) * {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units.get("gpu__time_duration.sum", "us"), 1.0),
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "dram_throughput_pct": num("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    "lts_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    "l1tex_throughput_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    "sm_throughput_pct": num("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    "issue_active_pct": num("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "tensor_pipe_active_pct": num("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active") or num("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": num("launch__registers_per_thread"),
    "local_spill_requests": num("l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
}
for kv in sys.argv[2:]:
    k, v = kv.split("=", 1)
    try:
        out[k] = json.loads(v)
    except ValueError:
        out[k] = v
print(json.dumps(out, indent=1))
