"""Per-kernel DRAM traffic / time / instruction counts from `ncu -i rep --page raw --csv`, in the format of
profiles/r*_traffic.json (what bench.py reports as roofline.traffic).
    python tools/traffic_from_raw.py raw.csv "<source description>" > profiles/rN_traffic.json"""
import collections
import csv
import json
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
        'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'inst': 1.0}
# cudaEvent brackets of bench.py -> the kernels inside them
BRACKETS = {'knn_search_kernel': ('knn_block_kernel', 'knn_select_kernel'),
            'cc_link_kernel': ('cc_link_kernel',), 'raster_kernel': ('raster_kernel',),
            'pip_query_kernel': ('pip_query_kernel',), 'grow_kernel': ('grow_kernel', 'grow_seed_pass_kernel')}


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

    def val(r, name):
        return float(r[col[name]].replace(',', '')) * UNIT.get(units[col[name]], 1.0)
    per = collections.OrderedDict()
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = per.setdefault(short(r[col['Kernel Name']]), {'launches': 0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0,
                                                            'time_ms': 0.0, 'warp_inst': 0.0})
        k['launches'] += 1
        k['dram_read_bytes'] += val(r, 'dram__bytes_read.sum')
        k['dram_write_bytes'] += val(r, 'dram__bytes_write.sum')
        k['time_ms'] += val(r, 'gpu__time_duration.sum')
        k['warp_inst'] += val(r, 'smsp__inst_executed.sum')
    brackets = {}
    for name, parts in BRACKETS.items():
        ks = [v for k, v in per.items() if any(k.startswith(p) for p in parts)]
        if not ks:
            continue
        total = sum(v['dram_read_bytes'] + v['dram_write_bytes'] for v in ks)
        # one bracket per pipeline pass for the search / growth, one per launch otherwise
        launches = max(v['launches'] for v in ks) if len(parts) == 1 else min(v['launches'] for v in ks)
        brackets[name] = total / max(launches, 1)
    json.dump({'source': source, 'per_kernel': per, 'bracket_traffic_bytes_per_launch': brackets}, sys.stdout, indent=1)
    print()


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
