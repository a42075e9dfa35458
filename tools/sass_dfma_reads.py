"""Operand-read cost of the DFMA instructions in a SASS listing (cuobjdump -sass): a DFMA that has to read three
source operands from the register file takes 3 cycles of the FP64 pipe instead of 2 (tools/probe_rf.cu); an operand the
previous instruction flagged .reuse in the same slot comes from the reuse cache; the same register in two slots is one read.
usage: sass_dfma_reads.py file.sass [first_line last_line]   -> per function (or line range): DFMAs, reads, pipe-cycle ceiling"""
import re
import sys


def analyse(lines):
    cache, n, cost, reads_tot, hist = {}, 0, 0, 0, {1: 0, 2: 0, 3: 0, 0: 0}
    for l in lines:
        m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Z][\w.]*)\s+(.*?);", l)
        if not m:
            continue
        op, args = m.group(2), [a.strip() for a in m.group(3).split(",")]
        if not op.startswith("DFMA"):
            continue          # other instructions use their own operand slots only when they carry .reuse themselves; ignored
        n += 1
        fresh, newcache = set(), {}
        for slot, s in enumerate(args[1:4]):
            reg = s.replace(".reuse", "").lstrip("-|").rstrip("|")
            if cache.get(slot) != reg and reg.startswith("R") and reg != "RZ":
                fresh.add(reg)
            if s.endswith(".reuse"):
                newcache[slot] = reg
        cache = newcache
        r = len(fresh)
        hist[r] = hist.get(r, 0) + 1
        reads_tot += r
        cost += max(2, r)
    return n, reads_tot, cost, hist


def hot_loop(lines):
    """The branch-free stretch with the most DFMAs of a function's SASS lines."""
    best, cur = [], []
    for l in lines + ["BRA"]:
        if re.search(r"\b(BRA|EXIT|RET|CALL|BSYNC|WARPSYNC)\b", l):
            if sum("DFMA" in x for x in cur) > sum("DFMA" in x for x in best):
                best = cur
            cur = []
        else:
            cur.append(l)
    return best


def kernel_ceiling(so_path, mangled):
    """Issue ceiling (fraction of the DFMA peak) of the hot loop of one kernel of a built library, or None."""
    import subprocess
    try:
        out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, so_path], capture_output=True, text=True, timeout=60).stdout
    except Exception:
        return None
    n, r, c, h = analyse(hot_loop(out.splitlines()))
    return {"dfma_in_hot_loop": n, "three_operand_reads": h.get(3, 0), "frac_of_dfma_peak": 2.0 * n / c} if n >= 16 else None


if __name__ == "__main__":
    src = open(sys.argv[1]).read().splitlines()
    if len(sys.argv) > 3:
        blocks = [("lines %s-%s" % (sys.argv[2], sys.argv[3]), src[int(sys.argv[2]):int(sys.argv[3])])]
    else:
        idx = [i for i, l in enumerate(src) if "Function :" in l] + [len(src)]
        blocks = [(src[a].split(":")[1].strip(), src[a:b]) for a, b in zip(idx[:-1], idx[1:])]
    for name, lines in blocks:
        # the hot loop = the branch-free stretch with the most DFMAs
        best, cur = [], []
        for l in lines + ["BRA"]:
            if re.search(r"\b(BRA|EXIT|RET|CALL|BSYNC|WARPSYNC)\b", l):
                if sum("DFMA" in x for x in cur) > sum("DFMA" in x for x in best):
                    best = cur
                cur = []
            else:
                cur.append(l)
        n, r, c, h = analyse(best)
        if n >= 16:
            print("%-60s hot loop: DFMA %4d  reads/DFMA %.2f  3-read %4d  ceiling %.1f%%" % (name[:60], n, r / n, h.get(3, 0), 200.0 * n / c))
