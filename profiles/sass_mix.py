"""Summarise an `ncu --page source --csv` export: opcode mix, stall samples, divergence. Usage: sass_mix.py file.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
ia, ie, isamp, ith = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
f = lambda s: float(s) if s not in ("", "-") else 0.0
tot = sum(f(r[ie]) for r in data)
tth = sum(f(r[ith]) for r in data)
print("warp instructions %.4g  thread instructions %.4g  avg active lanes %.2f  sass lines %d" % (tot, tth, tth / tot, len(data)))
cls, samp = collections.Counter(), collections.Counter()
for r in data:
    t = r[ia].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    cls[op] += f(r[ie])
    samp[op] += f(r[isamp])
ts = sum(samp.values()) or 1
for op, c in cls.most_common(24):
    print("%-10s %6.2f%% of instructions  %6.2f%% of stall samples" % (op, 100 * c / tot, 100 * samp[op] / ts))
print("stall reasons (all samples):")
for name in hdr:
    if name.startswith("stall_") and "Not Issued" not in name:
        v = sum(f(r[hdr.index(name)]) for r in data)
        if v > 0.01 * ts:
            print("  %-24s %6.2f%%" % (name, 100 * v / ts))
