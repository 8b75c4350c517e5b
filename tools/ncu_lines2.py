"""Per-source-line instruction counts / stall samples of one kernel from an ncu report.
   python tools/ncu_lines2.py report.ncu-rep kernel_regex [min_share]"""
import csv
import subprocess
import sys
rep, rx = sys.argv[1], sys.argv[2]
share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      rx, '--launch-count', '1', '--launch-skip', sys.argv[4] if len(sys.argv) > 4 else '0'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
def I(x):
    try:
        return int(x)
    except Exception:
        return 0
tot = 0; st_tot = 0; agg = {}
for r in rows:
    if len(r) > 8 and r[0].isdigit():
        v = I(r[7]); tot += v; st_tot += I(r[4])
        a = agg.setdefault(int(r[0]), [r[1][:110], 0, 0]); a[1] += v; a[2] += I(r[4])
print('total warp-instr', tot, 'stall samples', st_tot)
for ln in sorted(agg):
    src, v, st = agg[ln]
    if v > tot * share or st > st_tot * share:
        print(f'{ln:5d} {v:11d} {v / max(tot,1):6.3f}  st {st / max(st_tot,1):6.3f}  {src}')
