"""SASS opcode histogram of one kernel in an object / shared library (static; cuobjdump -sass).
usage: sass_hist.py <file> <substring of the mangled or demangled kernel name> [--loop]
--loop restricts the count to the hottest backward-branch loop body (largest one)."""
import re, subprocess, sys, collections
f, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", f], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for fn in funcs[1:]:
    name = fn.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat not in name and pat not in dem:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)[^;]*;", fn)
    addrs = [int(a, 16) for a, _, _ in ins]
    ops = [o for _, _, o in ins]
    lo, hi = 0, len(ops)
    if "--loop" in sys.argv:
        best = None
        for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?BRA[A-Z.]*\s+(?:U?!?U?P\d,?\s*)?`?\(?\.?L?_?x?_?\d*\)?[^;]*0x([0-9a-f]+)", fn):
            src, dst = int(m.group(1), 16), int(m.group(3), 16)
            if dst < src and (best is None or src - dst > best[1] - best[0]):
                best = (dst, src)
        if best:
            lo = addrs.index(best[0]) if best[0] in addrs else 0
            hi = addrs.index(best[1]) + 1
    h = collections.Counter(o.split(".")[0] for o in ops[lo:hi])
    print(dem[:110])
    print("  instructions: %d" % (hi - lo), " ".join("%s:%d" % kv for kv in h.most_common()))
