"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_source_hot.py report.ncu-rep [top [all_lines.csv]]"""
import csv, subprocess, sys
def main(rep, top=45):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur, hdr, agg = None, None, []
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] in ("Function Name",) or hdr is None: continue
        if r[0] != "":
            d = dict(zip(hdr[4:], r[4:]))
            try:
                agg.append((cur, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
            except ValueError:
                pass
    ti = sum(a[4] for a in agg); ts = sum(a[3] for a in agg); tt = sum(a[5] for a in agg)
    print(f"total warp-instr {ti:.4g}  thread-instr {tt:.4g}  avg lanes {tt/ti:.2f}  samples {ts}")
    print(f"{'file:line':24s} {'%inst':>6s} {'%smpl':>6s} {'lanes':>5s}  source")
    for a in sorted(agg, key=lambda a: -a[4])[:int(top)]:
        print(f"{a[0][:16]+':'+str(a[1]):24s} {100*a[4]/ti:6.2f} {100*a[3]/ts:6.2f} {a[5]/max(a[4],1):5.1f}  {a[2]}")
    if len(sys.argv) > 3:
        # every line, for offline aggregation by phase (tools/ncu_phase_share.py)
        with open(sys.argv[3], "w") as f:
            f.write("file,line,samples,warp_inst,thread_inst\n")
            for a in sorted(agg, key=lambda a: (a[0], a[1])):
                f.write(f"{a[0]},{a[1]},{a[3]},{a[4]},{a[5]}\n")
if __name__ == "__main__":
    main(*sys.argv[1:3])
