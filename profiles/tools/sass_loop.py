"""Opcode histogram of the hot loop of a kernel from `cuobjdump -sass` output (no GPU needed).

usage: python profiles/tools/sass_loop.py <lib.so> <kernel-name-substring> [--dump]
The hot loop is taken as the innermost backward branch span that holds the most VIMNMX/FMNMX instructions.
"""
import collections
import re
import subprocess
import sys


def kernels(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, out = None, {}
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    t = ins.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


def hot_loop(ins, key=None):
    if key is None:  # the packed kernels are recognised by their int16x2 min/max, the others by any min/max
        key = (".S16x2", ".16x2") if any(".S16x2" in s for _, s in ins) else ("VIMNMX", "FMNMX", "VIADDMNMX")
    addr = [a for a, _ in ins]
    best = None
    for i, (a, s) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)* (?:`\(\.L_x_\d+\)|0x([0-9a-f]+))", s)
        if not m or not m.group(1):
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        j = addr.index(tgt) if tgt in addr else None
        if j is None:
            continue
        body = ins[j:i + 1]
        n = sum(1 for _, x in body if any(k in x for k in key))
        # innermost loop that still holds (nearly) all of the min/max instructions
        if n and (best is None or n > best[0] * 1.5 or (n >= best[0] * 0.6 and len(body) < len(best[1]))):
            best = (n, body)
    return best[1] if best else []


if __name__ == "__main__":
    so, pat = sys.argv[1], sys.argv[2]
    for name, ins in kernels(so).items():
        if pat not in name:
            continue
        body = hot_loop(ins)
        h = collections.Counter(opcode(s) for _, s in body)
        print(f"== {name}: {len(ins)} instr, hot loop {len(body)} instr [{body[0][0]:#x}..{body[-1][0]:#x}]" if body else f"== {name}: no loop")
        for k, v in h.most_common():
            print(f"   {v:5d} {k}")
        if "--dump" in sys.argv:
            for a, s in body:
                print(f"  {a:06x} {s}")
