"""dev tool: SASS-level profile of one kernel in an ncu report, grouped into code regions of equal
execution count (loop bodies), with instruction share, stall-sample share and active lanes.
usage: python tools/ncu_regions.py report.ncu-rep [min-share-%]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; floor = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:120])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < 10: continue
    data.append((int(r[0], 16), r[1].strip(), int(r[ix['# Samples']]), int(r[ix['Instructions Executed']]), int(r[ix['Thread Instructions Executed']])))
base = data[0][0]
tot_s = sum(d[2] for d in data); tot_i = sum(d[3] for d in data); tot_t = sum(d[4] for d in data)
print(f"samples {tot_s}  warp-instr {tot_i:.4e}  thread-instr {tot_t:.4e}  lanes/instr {tot_t / tot_i:.2f}")
cur = None; blocks = []
for d in data:
    off = d[0] - base
    if cur is None or abs(d[3] - cur['i0']) > 0.25 * max(cur['i0'], 1):
        cur = {'start': off, 'i0': d[3], 'n': 0, 's': 0, 'i': 0, 't': 0, 'first': d[1]}
        blocks.append(cur)
    cur['n'] += 1; cur['s'] += d[2]; cur['i'] += d[3]; cur['t'] += d[4]; cur['end'] = off; cur['last'] = d[1]
print(f"{'sass range':>13s} {'n':>4s} {'runs':>10s} {'%inst':>6s} {'%smpl':>6s} {'lanes':>5s}  first .. last instruction")
for b in blocks:
    if 100 * b['i'] / tot_i > floor or 100 * b['s'] / tot_s > floor:
        print(f"{b['start']:6x}-{b['end']:6x} {b['n']:4d} {b['i0']:10.3e} {100 * b['i'] / tot_i:6.2f} {100 * b['s'] / tot_s:6.2f} {b['t'] / max(1, b['i']):5.1f}  {b['first'][:34]} .. {b['last'][:30]}")
