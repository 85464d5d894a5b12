#!/usr/bin/env python
"""CPU tool: condense build/ptxas<_tag>.log (written by mhdflows_jl_b200/build.py, nvcc -Xptxas -v) into one line per
kernel instantiation: registers, stack frame, spill stores / loads, static shared memory.

    python tools/ptxas_summary.py [build/ptxas.log] > profiles/rNN_ptxas_summary.txt
    python tools/ptxas_summary.py --diff profiles/r01_ptxas_summary.txt    # instantiations whose numbers changed

Static evidence only (no GPU): run it after every kernel change, before spending GPU time.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.split("\n")
        return out[:len(names)]
    except Exception:
        return list(names)


def short(sig):
    """'void mhdf::k_pass<float, 256, ...>(args)' -> 'k_pass<float, 256, ...>'"""
    s = re.sub(r"^void ", "", sig).replace("mhdf::", "")
    s = re.sub(r"\btrue\b", "1", re.sub(r"\bfalse\b", "0", s))   # same spelling as the round-1 summary (bools as 0 / 1)
    s = re.sub(r"\((?:mhdf::)?\w+\)(\d+)", r"\1", s)                # enum template arguments: (Phys)1 -> 1
    depth = 0
    for i, c in enumerate(s):
        if c == "<":
            depth += 1
        elif c == ">":
            depth -= 1
        elif c == "(" and depth == 0:
            return s[:i]
    return s


def parse(path):
    rows, cur = {}, None
    with open(path) as f:
        for line in f:
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = {"name": m.group(1), "stack": 0, "st": 0, "ld": 0, "regs": 0, "smem": 0}
                rows[cur["name"]] = cur
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                cur["stack"], cur["st"], cur["ld"] = map(int, m.groups())
            m = re.search(r"Used (\d+) registers", line)
            if m:
                cur["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(s.group(1)) if s else 0
    names = list(rows)
    out = {}
    for mangled, sig in zip(names, demangle(names)):
        out[short(sig)] = rows[mangled]
    return out


def read_summary(path):
    """Rows of an earlier summary file: name -> (regs, stack, 'st/ld')."""
    old = {}
    with open(path) as f:
        for line in f:
            m = re.match(r"^(k_\S.*?)\s+(\d+)\s+(\d+)\s+(\d+/\d+)", line)
            if m:
                old[m.group(1).strip()] = (int(m.group(2)), int(m.group(3)), m.group(4))
    return old


def main(argv):
    diff = None
    if "--diff" in argv:
        i = argv.index("--diff")
        diff = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    path = argv[0] if argv else os.path.join(ROOT, "build", "ptxas.log")
    rows = parse(path)
    if diff:
        old = read_summary(diff)
        new = {k: (v["regs"], v["stack"], f"{v['st']}/{v['ld']}") for k, v in rows.items()}
        changed = [(k, old[k], new[k]) for k in sorted(new) if k in old and old[k] != new[k]]
        added = [k for k in sorted(new) if k not in old]
        gone = [k for k in sorted(old) if k not in new]
        print(f"# {len(new)} instantiations now, {len(old)} in {os.path.basename(diff)}: {len(changed)} changed, {len(added)} new, {len(gone)} gone")
        for k, o, n in changed:
            print(f"changed  {k:<92} regs {o[0]} -> {n[0]}, stack {o[1]} -> {n[1]}, spills {o[2]} -> {n[2]}")
        for k in added:
            print(f"new      {k:<92} regs {new[k][0]}, stack {new[k][1]}, spills {new[k][2]}")
        for k in gone:
            print(f"gone     {k}")
        return 0
    print("# nvcc -gencode arch=compute_100a,code=sm_100a -Xptxas -v : registers / stack / spills / static smem per kernel instantiation")
    print(f"{'kernel':<96}{'regs':>6}{'stack':>7}{'spill st/ld':>13}{'smem':>8}")
    for k in sorted(rows):
        v = rows[k]
        print(f"{k:<96}{v['regs']:>6}{v['stack']:>7}{(str(v['st']) + '/' + str(v['ld'])):>13}{v['smem']:>8}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
