"""Static instruction mix of the hottest loop of a kernel from `cuobjdump -sass` (no GPU needed): the innermost-largest backward
branch region of each function matching a pattern, opcodes grouped into classes.  For the rjl kernels one loop trip handles two
list slots.  Counts are static: a branch inside the loop body (switch zone, wrapped pair) is counted once although it is rarely taken,
so the table also gives the straight-line path (instructions outside any forward-branch shadow).
  python profiles/static_loop_mix.py pfmds_b200/csrc/forces.o k_rjl_force k_rjl_density"""
import re
import subprocess
import sys
from collections import Counter


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    name, rows, res = None, [], {}
    for l in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", l)
        if m:
            if name:
                res[name] = rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        res[name] = rows
    return res


def klass(op):
    op = op.split(".")[0]
    if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
        return "fp64"
    if op == "MUFU":
        return "mufu"
    if op in ("LDG", "LD", "LDC", "LDCU", "ULDC", "LDS", "STG", "ST", "STS"):
        return "memory/const"
    if op in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "CALL", "RET", "BRX"):
        return "control"
    if op in ("MOV", "IMAD", "UMOV", "IADD3", "IADD", "LEA", "LOP3", "SHF", "ISETP", "SEL", "FSEL", "VIMNMX3", "VIMNMX", "UIADD3", "ULOP3", "PLOP3", "PRMT", "R2UR", "S2R", "UISETP", "USEL", "IMNMX", "HFMA2", "FMUL", "FADD", "FFMA", "I2F", "F2F", "CS2R", "UIMAD", "ULEA", "USHF"):
        return "int/move"
    return "other"


def hottest_loop(rows):
    best = None
    for k, (addr, text) in enumerate(rows):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:`\(\.\w+\)|0x([0-9a-f]+))", text)
        if m and m.group(1):
            tgt = int(m.group(1), 16)
            if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    return best


if __name__ == "__main__":
    obj, pats = sys.argv[1], sys.argv[2:]
    demangle = lambda s: subprocess.run(["c++filt", s], stdout=subprocess.PIPE, text=True).stdout.strip()
    for name, rows in functions(obj).items():
        dn = demangle(name)
        if not any(p in dn for p in pats):
            continue
        loop = hottest_loop(rows)
        if not loop:
            continue
        body = [(a, t) for a, t in rows if loop[0] <= a <= loop[1]]
        ops = [re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for _, t in body]
        mix = Counter(klass(o) for o in ops)
        print("%s\n  loop 0x%x..0x%x: %d instructions per trip (%s)" % (dn[:150], loop[0], loop[1], len(body), ", ".join("%s %d" % kv for kv in sorted(mix.items(), key=lambda kv: -kv[1]))))
        top = Counter(o.split(".")[0] for o in ops).most_common(12)
        print("  top opcodes: " + ", ".join("%s %d" % kv for kv in top))
