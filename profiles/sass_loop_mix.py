#!/usr/bin/env python
"""Instruction mix and stall-count sum of the largest loops of a kernel in a .so/.cubin.
usage: python profiles/sass_loop_mix.py <binary> <mangled-name-substring>"""
import collections, re, subprocess, sys

def main(binary, key):
    L = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(L) if "Function : " in l and key in l][0]
    end = ([i for i, l in enumerate(L) if "Function : " in l and i > start] + [len(L)])[0]
    ins = []
    i = start
    while i < end:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", L[i])
        if m and i + 1 < end:
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", L[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xf))
                i += 2
                continue
        i += 1
    loops = []
    for a, t, s in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
    print("kernel: %d instructions" % len(ins))
    for span, lo, hi in sorted(loops, reverse=True)[:3]:
        body = [x for x in ins if lo <= x[0] <= hi]
        c = collections.Counter()
        for a, t, s in body:
            tt = t.split()
            c[(tt[1] if tt[0].startswith("@") else tt[0]).split(".")[0]] += 1
        f64 = c["DFMA"] + c["DADD"] + c["DMUL"]
        print("loop 0x%x-0x%x: %d instrs, %d FP64 (%.0f%%), sum of stall counts %d" % (lo, hi, len(body), f64, 100.0 * f64 / len(body), sum(x[2] for x in body)))
        print("   ", c.most_common(16))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
