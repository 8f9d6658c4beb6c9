#!/usr/bin/env python3
"""Write profiles/r2_ncu_metrics.json: per-launch hardware counters of the dominant kernels from `ncu --set full`
captures of bench.py (read here, without a GPU). bench.py reports them as roofline.traffic / executed_flop_frac.
usage: tools/ncu_metrics_json.py <workload> <ris.ncu-rep> <winner.ncu-rep> <trace.ncu-rep> [note]"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def raw(path):
    rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    r = rows[2]

    def get(key):
        if key not in hdr:
            return None
        v = float(r[hdr.index(key)].replace(",", ""))
        u = units[hdr.index(key)].lower()
        scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        return v * scale
    return r[hdr.index("Kernel Name")], get


def main():
    workload, ris, winner, trace = sys.argv[1:5]
    note = sys.argv[5] if len(sys.argv) > 5 else ""
    out_path = ROOT / "profiles" / "r2_ncu_metrics.json"
    data = json.loads(out_path.read_text()) if out_path.exists() else {}
    entry = dict(source=f"ncu --set full --clock-control none, one launch per kernel of `bench.py --workload {workload} --steps 1 --warmup 1` {note}".strip())
    shading_dram, shading_flop = 0.0, 0.0
    for tag, path in (("ris", ris), ("winner", winner)):
        name, get = raw(path)
        dram = (get("dram__bytes_read.sum") or 0.0) + (get("dram__bytes_write.sum") or 0.0)
        # the full set reports these counters as rates per elapsed cycle
        cycles = get("smsp__cycles_elapsed.max") or get("sm__cycles_elapsed.max") or 0.0
        rate = lambda op: get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") or 0.0
        flop = (2.0 * rate("ffma") + rate("fmul") + rate("fadd")) * cycles
        entry[f"{tag}_kernel"] = name[:80]; entry[f"{tag}_dram_bytes_per_launch"] = dram; entry[f"{tag}_executed_flop_per_launch"] = flop
        entry[f"{tag}_us"] = (get("gpu__time_duration.sum") or 0.0)
        shading_dram += dram; shading_flop += flop
    entry["shading_dram_bytes_per_launch"] = shading_dram
    entry["shading_executed_flop_per_launch"] = shading_flop
    name, get = raw(trace)
    entry["trace_kernel"] = name[:80]
    entry["trace_dram_bytes_per_launch"] = (get("dram__bytes_read.sum") or 0.0) + (get("dram__bytes_write.sum") or 0.0)
    sectors = get("lts__t_sectors.sum")
    entry["trace_l2_bytes_per_launch"] = get("lts__t_bytes.sum") or (32.0 * sectors if sectors else None)
    data[workload] = entry
    out_path.write_text(json.dumps(data, indent=1) + "\n")
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
