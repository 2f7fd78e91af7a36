"""Summarise an .ncu-rep: key raw metrics, opcode mix, stall reasons, hottest SASS lines."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    print("==", d.get("Kernel Name", ("?",))[0][:100])
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
              "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
              "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
              "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
              "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
              "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]:
        if k in d:
            print(f"  {k:70s} {d[k][0]} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, hi in enumerate(starts):
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    print("== source:", rows[hi - 1][1][:80] if hi > 0 and len(rows[hi - 1]) > 1 else "")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    op, stall, lines = collections.Counter(), collections.Counter(), []
    tot = 0
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr):
            continue
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        o = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        o = o.split(".")[0]
        try:
            n = int(float(r[ix["Instructions Executed"]] or 0))
        except ValueError:
            continue
        op[o] += n
        tot += n
        smp = int(float(r[ix["# Samples"]] or 0))
        st = {c: int(float(r[ix[c]] or 0)) for c in stall_cols if r[ix[c]]}
        for c, v in st.items():
            stall[c] += v
        lines.append((smp, r[ix["Source"]][:90], max(st, key=st.get) if st else ""))
    tot = tot or 1
    print("total warp instructions", tot, "samples", sum(l[0] for l in lines))
    print("  " + "  ".join(f"{o}:{100*n/tot:.1f}%" for o, n in op.most_common(16)))
    ts = sum(stall.values()) or 1
    print("  stalls: " + "  ".join(f"{c[6:]}:{100*n/ts:.1f}%" for c, n in stall.most_common(9)))
    for smp, s_, why in sorted(lines, reverse=True)[:topn]:
        print(f"  {smp:6d} {why:22s} {s_}")
