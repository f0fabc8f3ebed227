"""Static view of a kernel's traversal loop in SASS (no GPU needed): finds the innermost loop around the 8-byte node
fetch (LDG.E.64) and prints its instructions with the pipe each one issues on (B300_MICROARCH.md: FFMA/FMUL/FADD/IMAD/
HFMA2 = FMA pipe, full rate; integer add / logic / shift / compare / select / min-max / MOV = ALU pipe, half rate).
   python tools/sass_loop.py <object or .so> <kernel-name-substring> [-q] [-a]"""
import re
import subprocess
import sys

FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2")
ALU = ("IADD3", "IADD", "VIADD", "LOP3", "PLOP3", "SHF", "LEA", "MOV", "SEL", "FSEL", "FSETP", "ISETP", "FMNMX", "FMNMX3", "IMNMX",
       "VIMNMX", "PRMT", "FLO", "POPC", "IABS", "VOTE", "P2R", "R2P", "BMSK", "SGXT", "LOP", "FCHK", "I2F", "F2I", "I2FP", "F2FP")
CTRL = ("BRA", "BSSY", "BSYNC", "BREAK", "EXIT", "RET", "WARPSYNC", "NOP", "BAR", "YIELD", "CALL")
MEM = ("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "ATOMG", "ATOMS", "RED", "LDCU")


def pipe(op):
    base = op.split(".")[0]
    if base in ("FLO", "POPC", "MUFU", "I2F", "F2I"):
        return "XU"
    if base in FMA:
        return "FMA"
    if base in ALU:
        return "ALU"
    if base in CTRL:
        return "CTL"
    if base in MEM:
        return "MEM"
    return "?" + base


def main():
    obj, name = sys.argv[1], sys.argv[2]
    quiet = "-q" in sys.argv
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    for f in funcs[1:]:
        fname = f.split("\n", 1)[0].strip()
        if name not in fname:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                text = m.group(2).strip()
                pred = ""
                pm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
                if pm:
                    pred, text = pm.group(1), pm.group(2)
                ins.append((int(m.group(1), 16), pred, text))
        addr = {a: i for i, (a, _, _) in enumerate(ins)}
        # innermost loop containing an LDG.E.64: smallest backward branch span around it
        ldg = [i for i, (_, _, t) in enumerate(ins) if t.startswith("LDG.E.64")]
        best = None
        loops = []
        for i, (a, p, t) in enumerate(ins):
            m = re.match(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) in addr and addr[int(m.group(1), 16)] <= i:
                lo = addr[int(m.group(1), 16)]
                if any(lo <= k <= i for k in ldg):
                    loops.append((lo, i))
                    if best is None or i - lo < best[1] - best[0]:
                        best = (lo, i)
        if best is None:
            print(fname, ": no loop found")
            continue
        if "-a" in sys.argv:      # every innermost loop around a node fetch (a frame kernel has one per ray class)
            inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
            for lo, hi in inner:
                c2 = {}
                for a, p, t in ins[lo:hi + 1]:
                    k = pipe(t.split()[0])
                    c2[k] = c2.get(k, 0) + 1
                print("  loop %04x..%04x: %d instructions %s" % (ins[lo][0], ins[hi][0], hi - lo + 1, dict(sorted(c2.items()))))
        lo, hi = best
        cnt = {}
        for a, p, t in ins[lo:hi + 1]:
            k = pipe(t.split()[0])
            cnt[k] = cnt.get(k, 0) + 1
            if not quiet:
                print("  %04x %-6s %-4s %s" % (a, p, k, t))
        print("%s\n  loop %04x..%04x: %d instructions %s" % (fname[:110], ins[lo][0], ins[hi][0], hi - lo + 1, dict(sorted(cnt.items()))))


if __name__ == "__main__":
    main()
