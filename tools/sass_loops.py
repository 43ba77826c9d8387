"""Instruction mix of the innermost loops (those containing MUFU.EX2) of the blend kernels, from cuobjdump -sass.
Developer tool: python tools/sass_loops.py [object file] [kernel substring]"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1] if len(sys.argv) > 1 else "street-gaussians-ns_b200/build/blend.o"
want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ALU = ("FSETP", "ISETP", "SEL", "FSEL", "LOP3", "PLOP3", "FMNMX", "IMNMX", "IADD3", "SHF", "I2F", "F2I", "IABS", "LEA", "VIADD", "IMAD.MOV", "P2R", "R2P", "FCHK", "IADD", "VIMNMX", "I2FP")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if want not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", f, re.M)]
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
                body = ins[addr[int(m.group(1), 16)]: i + 1]
                nex = sum("MUFU.EX2" in x for _, x in body)
                if nex:
                    loops.append((nex, len(body), body))
    print(name)
    for nex, n, body in sorted(loops, key=lambda x: (x[0], x[1])):
        ops = [re.sub(r"^@!?U?P\d+\s+", "", x).split()[0] for _, x in body]
        c = collections.Counter(o.split(".")[0] for o in ops)
        print(f"  ex2={nex} len={n}: " + " ".join(f"{k}:{v}" for k, v in c.most_common()))
