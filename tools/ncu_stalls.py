"""Warp-stall breakdown of every launch in an .ncu-rep (`--set full` capture, raw page): the cycles an average warp spends
in each stall reason per instruction it issues (smsp__average_warps_issue_stalled_*_per_issue_active), largest first, with
the issue-slot utilisation and warp latency per instruction beside it.  Says what a latency-bound kernel waits on."""
import csv, subprocess, sys
PRE, SUF = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
HEAD = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    print("==", rep)
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:80])
        for w in HEAD:
            if w in hdr:
                print(f"  {w:60s} {r[hdr.index(w)]:>14s} {rows[1][hdr.index(w)]}")
        st = sorted(((float(r[i] or 0), n[len(PRE):-len(SUF)]) for i, n in enumerate(hdr)
                     if n.startswith(PRE) and n.endswith(SUF) and "_not_issued" not in n), reverse=True)
        tot = sum(v for v, _ in st)
        print("  stall cycles per issued instruction (share of the warp's latency):")
        for v, n in st[:7]:
            print(f"    {n:28s} {v:8.3f}  {100 * v / tot:5.1f} %")
