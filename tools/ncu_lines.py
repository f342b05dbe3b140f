"""Per-source-line instruction and stall-sample totals of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on):   python tools/ncu_lines.py report.ncu-rep [rows] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, hdr, out = None, None, []
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == "File Path": fname = row[1].split("/")[-1]; hdr = None; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr and row[0] not in ("", "..."):
        d = dict(zip(hdr, row))
        try:
            out.append((fname, int(row[0]), row[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"])))
        except ValueError:
            pass
tot_i = sum(o[4] for o in out); tot_s = sum(o[3] for o in out)
print(f"total warp-instructions {tot_i}" + (f" = {tot_i/units:.0f}/row" if units else "") + f", samples {tot_s}")
for f, ln, src, s, n in sorted(out, key=lambda o: -o[4])[:top]:
    per = f"{n/units:7.1f}/row" if units else f"{n:12d}"
    print(f"{per} {100*n/tot_i:5.1f}% instr {100*s/max(tot_s,1):5.1f}% stall  {f}:{ln}  {src}")
