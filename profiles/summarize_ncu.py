#!/usr/bin/env python
"""Turns an `ncu --set full --import-source on` report into the text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep <total hops in the profiled launch> > profiles/<name>.txt
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    # round 2: the memory side of the table look-ups
    "sm__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
]
STALLS = "smsp__average_warps_issue_stalled_"


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, hops = sys.argv[1], float(sys.argv[2])
    json_out = sys.argv[3] if len(sys.argv) > 3 else None
    members = float(sys.argv[4]) if len(sys.argv) > 4 else None
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"kernel: {kname}")
    print(f"hops in this launch: {hops:.6g}")
    d = dict(zip(hdr, zip(units, vals)))
    for k in WANT:
        if k in d:
            print(f"{k:90s} {d[k][1]:>16s} {d[k][0]}")
    stalls = sorted(((float(v[1].replace(",", "")), k) for k, v in d.items() if k.startswith(STALLS) and k.endswith("_per_warp_active.pct")
                     ), reverse=True)[:8]
    for v, k in stalls:
        print(f"stall {k[len(STALLS):-len('_per_warp_active.pct')]:40s} {v:8.2f} %")
    # cycles a warp waits per instruction it issues, by reason (their sum = cycles between two issues of a warp)
    per_issue = sorted(((float(v[1].replace(",", "") or 0), k) for k, v in d.items()
                        if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and "not_issued" not in k), reverse=True)[:9]
    for v, k in per_issue:
        print(f"stall cycles per issued instruction: {k[len(STALLS):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
    if "smsp__inst_executed.sum" in d:
        inst = float(d["smsp__inst_executed.sum"][1].replace(",", ""))
        print(f"warp-instructions per hop: {inst / hops:.1f}")
    if "gpu__time_duration.sum" in d:
        t = float(d["gpu__time_duration.sum"][1].replace(",", ""))
        u = d["gpu__time_duration.sum"][0]
        t *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
        print(f"hops/s under the profiler (serialised, cold): {hops / t:.4g}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    h2 = src[1]
    isrc, iex = h2.index("Source"), h2.index("Instructions Executed")
    hist = collections.Counter()
    for r in src[2:]:
        t = r[isrc].strip().split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        hist[op.split(".")[0]] += int(r[iex])
    if json_out:
        import json

        def num(k):
            return float(d[k][1].replace(",", "")) if k in d else None

        def byt(k):
            if k not in d:
                return None
            return num(k) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d[k][0], 1.0)
        out = {"source": f"{rep} (ncu --set full --clock-control none)", "kernel": kname, "hops_in_capture": hops,
               "members_in_capture": members, "warp_inst_per_hop": inst / hops,
               "ipc_per_sm": num("sm__inst_executed.avg.per_cycle_active"),
               "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "xu_pipe_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
               "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
               "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
               "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
               "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "smem_wavefront_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
               "smem_bank_conflict_wavefronts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
               "smem_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
               "registers_per_thread": num("launch__registers_per_thread"),
               "achieved_occupancy_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
               "l1tex_data_pipe_wavefront_pct": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
               "l1_sector_hit_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_sector_hit_pct": num("lts__t_sector_hit_rate.pct"),
               "l2_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "l2_sectors_per_hop": (num("lts__t_sectors.sum") or 0) / hops,
               "sm_active_frac": (num("sm__cycles_active.avg") or 0) / (num("sm__cycles_elapsed.max") or 1)}
        if len(sys.argv) > 7:
            out.update({"hops_per_member": int(sys.argv[5]), "N": int(sys.argv[6]), "P": int(sys.argv[7])})
        out["dram_bytes_per_hop"] = ((byt("dram__bytes_read.sum") or 0) + (byt("dram__bytes_write.sum") or 0)) / hops
        if members:
            out["dram_bytes_read_per_member"] = byt("dram__bytes_read.sum") / members
            out["dram_bytes_written_per_member"] = byt("dram__bytes_write.sum") / members
        json.dump(out, open(json_out, "w"), indent=1)
    print("SASS opcode mix, warp-instructions per hop:")
    print("  " + ", ".join(f"{op} {c / hops:.1f}" for op, c in hist.most_common(30)))


if __name__ == "__main__":
    main()
