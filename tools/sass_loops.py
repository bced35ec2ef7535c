"""Lists the loops of one kernel's SASS (cuobjdump -sass -fun NAME lib.so) with their instruction mix.

    python tools/sass_loops.py lib.so MANGLED_NAME [min_instructions] [--dump OP:N]

--dump OP:N prints the body of the innermost loops that hold exactly N instructions with mnemonic OP
(e.g. FFMA2:128 = the unrolled screen loop of prefixn_kernel).

A loop = the address range of a backward branch [target, branch]; only the innermost ranges are of interest.
"""
import collections
import re
import subprocess
import sys


def main():
    lib, fun = sys.argv[1], sys.argv[2]
    min_ins = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 16
    dump = None
    if "--dump" in sys.argv:
        op, n = sys.argv[sys.argv.index("--dump") + 1].split(":")
        dump = (op, int(n))
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
    ins = []
    for line in txt.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                loops.append((addr_idx[tgt], i))
    print(f"{fun}: {len(ins)} instructions, {len(loops)} backward branches")
    for lo, hi in loops:
        if hi - lo + 1 < min_ins:
            continue
        inner = [l for l in loops if l != (lo, hi) and l[0] >= lo and l[1] <= hi and l[1] - l[0] + 1 >= min_ins]
        ops = collections.Counter()
        for _, t in ins[lo:hi + 1]:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            ops[t.split()[0]] += 1
        top = ", ".join(f"{k} {v}" for k, v in ops.most_common(14))
        print(f"  loop {ins[lo][0]:#06x}..{ins[hi][0]:#06x} ({hi - lo + 1} instr{', has inner loops' if inner else ''}): {top}")
        if dump and not inner and ops.get(dump[0]) == dump[1]:
            for a_, t_ in ins[lo:hi + 1]:
                print(f"      /*{a_:04x}*/  {t_}")
            dump = None                      # the first match only


if __name__ == "__main__":
    main()
