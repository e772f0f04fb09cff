#!/usr/bin/env python
"""Summarise `ncu --set full` raw pages (tools/gpu_traffic.sh) into profiles/<tag>_traffic.json + a readable table.

    python tools/traffic_json.py profiles/r02 gpurun_out/r02_full_a_raw.csv gpurun_out/r02_full_b_raw.csv
Per kernel (first launch of each distinct name+grid): duration, DRAM bytes read+written, L2->SM bytes, achieved
occupancy, tensor-pipe and issue utilisation, registers, shared memory."""
import csv, json, re, sys

tag, paths = sys.argv[1], sys.argv[2:]
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
want = {
    "us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum",
    "dram_wr": "dram__bytes_write.sum",
    "l2_to_l1": "lts__t_sectors_srcunit_tex_op_read.sum",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_pct_alt": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lds_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "lds_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "lds_pipe_pct": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex_lsu_pct": "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "smem_dyn": "launch__shared_mem_per_block_dynamic",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "sm_mhz": "sm__cycles_elapsed.avg.per_second",
}
kernels, table = {}, []
for path in paths:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("gr::", "")
        key = f"{name} grid={r[col['Grid Size']]}"
        if key in kernels:
            continue
        e = {}
        for k, m in want.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                e[k] = v * U.get(units[col[m]], 1.0)
        if "l2_to_l1" in e:
            e["l2_to_l1"] *= 32.0  # sectors -> bytes
        e["dram_bytes_per_launch"] = e.get("dram_rd", 0.0) + e.get("dram_wr", 0.0)
        if e.get("us"):
            e["dram_gbs"] = e["dram_bytes_per_launch"] / e["us"] * 1e-3
        kernels[key] = e
        table.append((key, e))
out = {"source": "ncu --set full --clock-control none (default cache control: caches flushed before each kernel), "
                 "first launch of each kernel+grid in the warm-up step of `bench.py --steps 1 --warmup 1`; " + ", ".join(paths),
       "kernels": {}}
for key, e in table:
    base = key.split(" grid=")[0]
    out["kernels"].setdefault(re.sub(r"<.*", "", base).replace("tc::", ""), e)  # plain name, first grid: bench.py's lookup key
    out["kernels"].setdefault(base, e)
    out["kernels"][key] = e
json.dump(out, open(tag + "_traffic.json", "w"), indent=1)
with open(tag + "_ncu_full_summary.txt", "w") as f:
    f.write("ncu --set full, one launch per kernel+grid (caches flushed before each launch: cold numbers)\n")
    f.write("%-62s %8s %9s %9s %7s %6s %6s %6s %5s\n" % ("kernel grid", "us", "dram MB", "dram GB/s", "dram%", "warp%", "issue%", "tens%", "regs"))
    for key, e in table:
        f.write("%-62s %8.1f %9.2f %9.0f %7.1f %6.1f %6.1f %6.1f %5d\n" % (
            key[:62], e.get("us", 0), e["dram_bytes_per_launch"] / 1e6, e.get("dram_gbs", 0), e.get("dram_pct", 0),
            e.get("warps_active_pct", 0), e.get("issue_active_pct", 0), e.get("tensor_pipe_pct", e.get("tensor_pipe_pct_alt", 0)), int(e.get("regs", 0))))
print(open(tag + "_ncu_full_summary.txt").read())
