"""DRAM traffic per launch of the hot kernels from an `ncu --set full` report, tied to the sources it was captured from:
    python tools/ncu_traffic.py <config> <report.ncu-rep> [<report2> ...]   -> updates profiles/r02_traffic.json
Kernel families as bench.py names them (dht, gather_push, deposit_J, deposit_rho); the figure is the mean over the
captured launches of dram__bytes_read.sum + dram__bytes_write.sum.  bench.py reports it as roofline.traffic only while
the hash of fbpic_b200/csrc/ still equals the one recorded here."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FAMILIES = [('dht', r'k_dht'), ('gather_push', r'k_gather_push'), ('deposit_J', r'k_deposit_mma<\d+, *(1|true)\b'),
            ('deposit_rho', r'k_deposit_mma<\d+, *(0|false)\b')]
UNIT = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    from bench import csrc_hash
    config, reports = sys.argv[1], sys.argv[2:]
    out_path = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    sha = csrc_hash()
    try:
        out = json.load(open(out_path))
    except Exception:
        out = {}
    if out.get('csrc_sha256') != sha:
        out = {'csrc_sha256': sha, 'kernels': {}}
    acc = {}
    for rep in reports:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[ix['Kernel Name']]
            fam = next((f for f, pat in FAMILIES if re.search(pat, name)), None)
            if fam is None:
                continue
            b = sum(float(r[ix[m]]) * UNIT[units[ix[m]]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            t = float(r[ix['gpu__time_duration.sum']])
            acc.setdefault(fam, []).append((b, t, name[:60], os.path.basename(rep)))
    for fam, v in acc.items():
        # the launches of a family differ in size (set-up launches, the J+rho and the E+B Hankel batches): the
        # figure is that of the LONGEST captured launch, the steady-state one
        b, t, name, rep = max(v, key=lambda x: x[1])
        out['kernels']['%s:%s' % (config, fam)] = {
            'dram_bytes_per_launch': b, 'launches_captured': len(v), 'duration_under_ncu_us': t,
            'source': 'ncu --set full --clock-control none, %s: longest of %d captured launches of %s (%.0f us under ncu)'
                      % (rep, len(v), name, t)}
    json.dump(out, open(out_path, 'w'), indent=1)
    print(json.dumps(out['kernels'], indent=1))


if __name__ == '__main__':
    main()
