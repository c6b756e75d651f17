"""Text summary of an .ncu-rep (read on the CPU box: `ncu -i ... --page raw/source --csv`): the metrics the DESIGN / profiles
tables quote, and the warp-stall samples by SASS region.  Usage:

    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep "<heading line>" > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__icc_request_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio", "sm__warps_active.avg.per_cycle_active"]
STALLS = {"wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
          "math_pipe_throttle": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "selected": "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
          "barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "dispatch_stall": "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
          "no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "branch_resolving": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
          "mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
          "sleeping": "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "membar": "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"}


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    print(sys.argv[2] if len(sys.argv) > 2 else rep)
    print()
    raw = page(rep, "raw")
    h, units = raw[0], raw[1]
    for row in raw[2:]:
        kname = row[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print("kernel:", kname[:140])
        for m in WANT:
            if m in h:
                i = h.index(m)
                print(f"{m:75s} {row[i]} {units[i]}")
        st = []
        for k, m in STALLS.items():
            if m in h:
                try:
                    st.append((float(row[h.index(m)]), k))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("stall reasons (warps per issue-active cycle): " + ", ".join(f"{k}={v:.2f}" for v, k in st[:10]))
        print()
    src = page(rep, "source")
    heads = [i for i, r in enumerate(src) if "Address" in r and "Source" in r]
    for n, h0 in enumerate(heads):
        hh = src[h0]
        data = [r for r in src[h0 + 1:(heads[n + 1] if n + 1 < len(heads) else len(src))] if len(r) == len(hh)]
        kn = src[h0 - 1][1][:100] if h0 > 0 and len(src[h0 - 1]) > 1 else ""
        region_table(hh, data, kn)


def region_table(hh, data, kernel_name):
    isrc, isamp, iex = hh.index("Source"), hh.index("Warp Stall Sampling (All Samples)"), hh.index("Instructions Executed")

    def num(v):
        try:
            return int(v or 0)
        except ValueError:
            return 0
    stc = [i for i, n in enumerate(hh) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(num(r[isamp]) for r in data) or 1
    totex = sum(num(r[iex]) for r in data) or 1
    step = 240
    print(f"warp-stall sampling by SASS region, {kernel_name} (blocks of {step} instructions in address order; {tot} samples, {totex} warp-instructions):")
    print("  first-instr  samples  executed  dominant opcodes | dominant stall reasons")
    for a in range(0, len(data), step):
        blk = data[a:a + step]
        s = sum(num(r[isamp]) for r in blk)
        e = sum(num(r[iex]) for r in blk)
        if s < tot * 0.005 and e < totex * 0.005:
            continue
        ops, d = {}, {}
        for r in blk:
            t = r[isrc].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
            for i in stc:
                v = num(r[i])
                if v:
                    d[hh[i][6:]] = d.get(hh[i][6:], 0) + v
        topo = ", ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
        ds = sum(d.values()) or 1
        tops = ", ".join(f"{k}:{100 * v // ds}%" for k, v in sorted(d.items(), key=lambda x: -x[1])[:5])
        print(f"  {a:7d}  {100 * s / tot:5.1f}%  {100 * e / totex:5.1f}%  {topo} | {tops}")
    print()


if __name__ == "__main__":
    main()
