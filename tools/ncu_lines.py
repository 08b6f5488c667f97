"""dev tool: per-source-line summary of an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top-N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ksub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = None; hdr = None; data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] not in ("", "Line No") and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            data.append((cur_fn, cur_file, int(r[0]), r[1].strip()[:90], int(r[7]), int(r[8]), int(r[4])))
        except ValueError:
            pass
data = [d for d in data if ksub in d[0]]
tot_i = sum(d[4] for d in data); tot_t = sum(d[5] for d in data); tot_s = sum(d[6] for d in data)
print(f"total warp-instr {tot_i:.3e} thread-instr {tot_t:.3e} avg threads {tot_t/max(1,tot_i):.2f} samples {tot_s}")
print(f"{'file:line':28s} {'%inst':>6s} {'%smpl':>6s} {'thr':>5s}  source")
for d in sorted(data, key=lambda d: -d[6])[:top]:
    print(f"{d[1]+':'+str(d[2]):28s} {100*d[4]/tot_i:6.2f} {100*d[6]/max(1,tot_s):6.2f} {d[5]/max(1,d[4]):5.1f}  {d[3]}")
