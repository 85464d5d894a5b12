"""Summarise `ncu --page source --csv` and `--page details --csv` exports (no GPU needed)."""
import csv, sys, collections, re

def source_summary(fn, top=25):
    with open(fn) as f:
        first = f.readline()
        rows = list(csv.DictReader(f))
    print("kernel:", first.strip()[:160])
    stall_cols = [c for c in rows[0].keys() if c and c.startswith("stall_") and "Not Issued" not in c]
    tot = collections.Counter()
    nsamp = 0
    ops = collections.Counter()
    opsamp = collections.Counter()
    conf = exc = ideal = 0
    for r in rows:
        if r.get("# Samples") in (None, ""):
            continue
        s = int(r["# Samples"]); nsamp += s
        for c in stall_cols:
            tot[c] += int(r[c] or 0)
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r["Source"])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += int(r["Instructions Executed"] or 0)
        opsamp[op] += s
        exc += int(r["L1 Wavefronts Shared Excessive"] or 0); ideal += int(r["L1 Wavefronts Shared Ideal"] or 0)
    print(f"samples={nsamp}  instr rows={len(rows)}")
    print("stalls: " + "  ".join(f"{k[6:]}={v*100/max(1,sum(tot.values())):.1f}%" for k, v in tot.most_common(9)))
    tot_i = sum(ops.values())
    print("instr mix (warp-level): " + "  ".join(f"{k}={v*100/tot_i:.1f}%" for k, v in ops.most_common(14)))
    print("samples by op: " + "  ".join(f"{k}={v*100/max(1,nsamp):.1f}%" for k, v in opsamp.most_common(10)))
    print(f"shared wavefronts: ideal={ideal} excessive={exc} ({exc*100/max(1,ideal):.1f}% extra)")
    rows.sort(key=lambda r: -int(r["# Samples"] or 0))
    print("top instructions by samples:")
    for r in rows[:top]:
        st = {c[6:]: int(r[c] or 0) for c in stall_cols if int(r[c] or 0)}
        st = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"  {int(r['# Samples']):6d}  {r['Source'].strip()[:70]:70s} {st}")

def details_summary(fn, keys=None):
    keys = keys or ['Duration', 'DRAM Throughput', 'Memory Throughput', 'Registers Per', 'Achieved Occupancy', 'Theoretical Occupancy',
                    'Compute (SM) Throughput', 'L1/TEX Hit', 'L2 Hit', 'Executed Ipc Active', 'Issue Slots Busy', 'Block Limit', 'Dynamic Shared',
                    'Mem Busy', 'Max Bandwidth', 'No Eligible', 'Eligible Warps', 'Local', 'Grid Size', 'Block Size']
    rows = list(csv.DictReader(open(fn)))
    ids = []
    for r in rows:
        if r['ID'] not in ids:
            ids.append(r['ID'])
    for kid in ids:
        name = next(r['Kernel Name'] for r in rows if r['ID'] == kid)
        print('==', kid, name[:110])
        for r in rows:
            if r['ID'] == kid and any(k in r['Metric Name'] for k in keys):
                print(f"   {r['Metric Name'][:44]:44s} {r['Metric Value']:>14s} {r['Metric Unit']}")

def raw_pick(fn, pats):
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '')[:100])
        for h, u in zip(hdr, units):
            if any(re.search(p, h) for p in pats):
                print(f"   {h:60s} {d[h]:>16s} {u}")

if __name__ == "__main__":
    mode, fn = sys.argv[1], sys.argv[2]
    if mode == "source":
        source_summary(fn, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    elif mode == "details":
        details_summary(fn)
    else:
        raw_pick(fn, sys.argv[3:] or [r"dram__bytes_(read|write)\.sum$", r"gpu__time_duration\.sum", r"launch__registers_per_thread", r"sm__warps_active\.avg\.pct", r"dram__throughput.avg.pct_of_peak_sustained_elapsed"])
