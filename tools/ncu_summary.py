#!/usr/bin/env python
"""Text summary of an ncu capture for profiles/: per kernel launch the headline counters (duration, DRAM bytes, L1 / L2
sector traffic, sectors per request, shared-memory wavefronts and bank conflicts, occupancy, issue rate), the warp-stall
breakdown and the most-stalled SASS instructions (the capture must have been taken with --import-source on).
usage: python tools/ncu_summary.py <file.ncu-rep> [kernel regex] > profiles/<name>.txt"""
import csv
import io
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__average_warp_latency_per_inst_issued.ratio"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else "."
    print(f"# {rep}  (ncu --set full --clock-control none --import-source on; summary by tools/ncu_summary.py)")
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[ci["Kernel Name"]]
        if not re.search(pat, name):
            continue
        print(f"\n== {name}   grid {r[ci['Grid Size']]} x block {r[ci['Block Size']]}")
        vals = {}
        for m in RAW:
            if m in ci:
                vals[m] = r[ci[m]]
                print(f"  {m:82s} {r[ci[m]]:>16s} {units[ci[m]]}")
        try:
            req = float(vals["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
            sec = float(vals["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
            print(f"  {'=> sectors per global-load request':82s} {sec / max(req, 1):16.2f}")
            w = float(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", ""))
            c = float(vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"].replace(",", ""))
            if w:
                print(f"  {'=> share of shared-memory wavefronts that are bank conflicts':82s} {c / w:16.2f}")
        except Exception:
            pass
        stalls = [(h, r[ci[h]]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        st = sorted(((float(v.replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0]) for h, v in stalls if v), reverse=True)[:7]
        print("  stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st))
    # source page: most-stalled SASS instructions
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    kname, block = None, []
    def flush():
        if not block or kname is None or not re.search(pat, kname):
            return
        h = block[0]
        c = {x: i for i, x in enumerate(h)}
        data = []
        for q in block[1:]:
            try:
                data.append((float(q[c["# Samples"]]), q))
            except Exception:
                pass
        tot = sum(v for v, _ in data) or 1.0
        print(f"\n-- most-stalled SASS instructions of {kname[:90]} ({int(tot)} samples)")
        names = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        for v, q in sorted(data, key=lambda t: -t[0])[:10]:
            top = max(((float(q[c[s]] or 0), s) for s in names))
            print(f"  {100 * v / tot:5.1f} %  {q[c['Source']].strip()[:70]:70s} {top[1]}")
    for q in src:
        if q and q[0] == "Kernel Name":
            flush()
            kname, block = q[1], []
        elif q:
            block.append(q)
    flush()


if __name__ == "__main__":
    main()
