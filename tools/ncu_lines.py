"""Join an ncu SASS-level source page (csv) with nvdisasm --print-line-info of the cubin, so
that executed-instruction counts and stall samples are attributed to CUDA source lines.
    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all libupcp_cuda.so ; nvdisasm --print-line-info knn.sm_100a.cubin > all.sass
    python tools/ncu_lines.py src.csv all.sass <mangled kernel name> [top]"""
import collections
import csv
import re
import sys

src_csv, sass, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur = None
inside = False
for l in open(sass):
    if l.startswith('.text.'):
        inside = l.strip() == f'.text.{kname}:'
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
allrows = list(csv.reader(open(src_csv)))
# the page holds one section per profiled kernel ("Kernel Name", <demangled>): keep the one asked for
short = re.sub(r'^_ZN\d+[a-z]+(\d+)', '', kname)
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name']
sel = [i for i in starts if any(part in allrows[i][1] for part in re.findall(r'[a-z_]{6,}', kname))] or starts[:1]
import os
beg = sel[int(os.environ.get('SECTION', '-1'))]
end = min([i for i in starts if i > beg] + [len(allrows)])
rows = allrows[beg:end]
hdr = rows[1]
ia, iinst, ithr, isamp = (hdr.index(h) for h in ('Address', 'Instructions Executed', 'Thread Instructions Executed', '# Samples'))
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    key = line_of.get(int(r[ia], 16) - base, ('?', 0))
    a = agg[key]
    for j, c in enumerate((iinst, ithr, isamp)):
        a[j] += int(r[c]); tot[j] += int(r[c])
print(f'total warp-instructions {tot[0]:.4g}, avg active lanes {tot[1] / tot[0]:.1f}, samples {tot[2]}')
cache = {}
for (f, ln), (ins, thr, smp) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    text = ''
    try:
        if f not in cache:
            cache[f] = open(f'urban_pointcloud_processing_b200/csrc/{f}').read().splitlines()
        text = cache[f][ln - 1].strip()[:80]
    except Exception:
        pass
    print(f'{100 * smp / tot[2]:5.1f}% stall  {100 * ins / tot[0]:5.1f}% inst  lanes {thr / max(ins, 1):5.1f}  {f}:{ln}  {text}')
