"""Summarise an ncu report's source page: top stalled SASS instructions and stall-reason totals."""
import collections, csv, subprocess, sys

def main(path, top=28, skip=0):
    sel = ["--launch-skip", str(skip), "--launch-count", "1"]     # one launch of a multi-launch report
    raw = subprocess.run(["ncu", "-i", path] + sel + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keep = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
            "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
    print(f"# {path} (launch {skip}): {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
    for h, u, v in zip(hdr, units, vals):
        if h in keep:
            print(f"{h} [{u}] = {v}")
    src = subprocess.run(["ncu", "-i", path] + sel + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ci, cs, cx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    st = {h: hdr.index(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    data, agg = [], collections.Counter()
    for r in rows[2:]:
        try:
            sm = float(r[cs] or 0)
        except Exception:
            continue
        data.append((sm, r[ci].strip(), r[cx], r))
        for h, i in st.items():
            try:
                agg[h] += float(r[i])
            except Exception:
                pass
    tot = sum(d[0] for d in data) or 1
    print("stall reasons (all warps):", ", ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in agg.most_common(9)))
    print("top sampled instructions:")
    for sm, s, ex, r in sorted(data, key=lambda d: -d[0])[:top]:
        why = max(st, key=lambda h: float(r[st[h]] or 0))
        print(f"{100*sm/tot:5.1f}% exec={ex:>9} {why[6:]:<14} {s[:96]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
