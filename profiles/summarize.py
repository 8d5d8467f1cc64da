#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv "header text"  > profiles/xx_launches_summary.txt
    python profiles/summarize.py full gpurun_out/prof_k.ncu-rep [...]             > profiles/xx_ncu_full_summary.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'launch__shared_mem_per_block_dynamic']


def launches(path, header):
    print("#", header)
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mn, mu = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Name', 'Metric Unit'))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv or r[mn] != 'gpu__time_duration.sum':
            continue
        v = float(r[mv].replace(',', ''))
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(r[mu], 1)
        name = r[kn].split('(')[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} n={v[0]:5d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e3:9.2f} us share={v[1] / tot:.3f}")


def full(paths):
    for p in paths:
        out = subprocess.run(['ncu', '-i', p, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr = rows[0]
        kn = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
        print("==", p, "| kernel:", rows[2][kn].split('(')[0] if kn is not None and len(rows) > 2 else '?',
              "| one column per captured launch (first row = unit)")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  ", w, [r[i] for r in rows[1:]])


def _raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def _num(v):
    try:
        return float(v.replace(',', ''))
    except ValueError:
        return None


def facts(out_json, specs):
    """specs: name=path[:launch] ...  -> JSON of the per-launch facts bench.py quotes (launch = index of the captured launch,
    default the last one).  Units are converted to bytes / microseconds."""
    import json
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6, 'usecond': 1, 'msecond': 1e3, 'nsecond': 1e-3}
    res = {}
    for spec in specs:
        name, rest = spec.split('=', 1)
        path, _, idx = rest.partition(':')
        hdr, units, rows = _raw(path)
        r = rows[int(idx) if idx else -1]
        def get(metric):
            if metric not in hdr:
                return None
            i = hdr.index(metric)
            v = _num(r[i])
            return None if v is None else v * scale.get(units[i], 1)
        rd, wr = get('dram__bytes_read.sum'), get('dram__bytes_write.sum')
        res[name] = {
            "source": f"profiles/{path.split('/')[-1].replace('.ncu-rep', '.txt')} (ncu --set full, launch {idx or 'last'} of the capture)",
            "kernel": r[hdr.index('Kernel Name')].split('(')[0],
            "dram_bytes_per_launch": None if rd is None else rd + wr,
            "dram_bytes_read": rd, "dram_bytes_write": wr,
            "duration_us_under_ncu": get('gpu__time_duration.sum'),
            "issue_slots_busy_pct": get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
            "sm_busy_pct": get('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
            "dram_throughput_pct": get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            "warps_active_pct": get('sm__warps_active.avg.pct_of_peak_sustained_active'),
            "grid": get('launch__grid_size'), "block": get('launch__block_size'), "registers": get('launch__registers_per_thread'),
            "waves_per_sm": get('launch__waves_per_multiprocessor'),
        }
    with open(out_json, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "facts":
        facts(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[2:])
