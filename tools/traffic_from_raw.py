"""Per-bracket DRAM traffic / time / instruction counts from `ncu --nvtx ... --page raw --csv`, in the format of
profiles/r2_traffic.json (what bench.py reports as roofline.traffic).  A kernel belongs to every bracket whose NVTX
range (api.cu pushes one per cudaEvent bracket, named like the bracket) is on its range stack; only the FIRST pipeline
pass of the capture is counted (one launch of the S4 bracket).
    python tools/traffic_from_raw.py raw.csv "<source description>" > profiles/r2_traffic.json"""
import collections
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
        'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'inst': 1.0}
BRACKETS = ['raster_kernel', 'pip_query_kernel', 'radix_sort (hist+scan+scatter)', 'cc_link_kernel',
            'knn_stage (build+search+epilogue)', 'grow_kernel', 'knn_build', 'knn_search (block tiers + per-query tier)',
            'knn_epilogue_kernel']

SINGLE = ('raster_kernel', 'pip_query_kernel', 'knn_epilogue_kernel')     # one kernel launch per bracket


def sources_sha():
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'urban_pointcloud_processing_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        if f.endswith(('.cu', '.cuh')):
            h.update(f.encode()); h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:16]


def short(name):
    name = name.replace('void ', '').replace('upcp::', '')
    depth, out = 0, ''
    for ch in name:                       # cut at the first '(' outside template brackets
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        elif ch == '(' and depth == 0:
            break
        out += ch
    return out


def main(path, source):
    rows = list(csv.reader(open(path, errors='replace')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    nvtx_cols = [i for i, h in enumerate(hdr) if 'Push/Pop' in h]

    def val(r, name):
        return float(r[col[name]].replace(',', '')) * UNIT.get(units[col[name]], 1.0)
    per = collections.OrderedDict()
    br = {b: {'kernels': 0, 'dram_bytes': 0.0, 'time_ms': 0.0, 'calls': 0} for b in BRACKETS}
    prev_inside = set()
    grown = False
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        stack = ' '.join(r[i] for i in nvtx_cols)
        name = short(r[col['Kernel Name']])
        inside = set(b for b in BRACKETS if b in stack)
        grown = grown or 'grow_kernel' in inside
        if grown and 'raster_kernel' in inside:
            break                          # the second pipeline pass begins (AHNFuser's raster launch)
        traffic = val(r, 'dram__bytes_read.sum') + val(r, 'dram__bytes_write.sum')
        t = val(r, 'gpu__time_duration.sum')
        k = per.setdefault(name, {'launches': 0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0, 'time_ms': 0.0,
                                  'warp_inst': 0.0})
        k['launches'] += 1
        k['dram_read_bytes'] += val(r, 'dram__bytes_read.sum')
        k['dram_write_bytes'] += val(r, 'dram__bytes_write.sum')
        k['time_ms'] += t
        k['warp_inst'] += val(r, 'smsp__inst_executed.sum')
        for b in inside:
            br[b]['kernels'] += 1; br[b]['dram_bytes'] += traffic; br[b]['time_ms'] += t
            # a bracket = one launch of the kernel it is named after, else one contiguous group of launches
            # carrying its range
            if name.split('<')[0] == b or (b not in SINGLE and b not in prev_inside):
                br[b]['calls'] += 1
        prev_inside = inside
    out = {b: {'calls': v['calls'], 'kernels': v['kernels'], 'bytes': v['dram_bytes'], 'time_ms': v['time_ms']}
           for b, v in br.items() if v['kernels']}
    json.dump({'source': source, 'csrc_sha': sources_sha(), 'per_kernel': per, 'brackets_first_pass': out,
               'bracket_traffic_bytes_per_launch': {b: v['bytes'] / max(v['calls'], 1) for b, v in out.items()},
               'note': 'cudaMemsetAsync of the 4 B x G cell table (~270 MB per S4 launch) is not a kernel and is not in these sums'},
              sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
