"""SASS instruction mix per kernel of the built cubin: python tools/sass_mix.py [cubin] > profiles/rN_sass_mix.txt
Static counts (every instruction of the function once, whatever its trip count); the ncu source page gives the executed mix."""
import collections
import re
import subprocess
import sys

cubin = sys.argv[1] if len(sys.argv) > 1 else "swiftvideo_b200/svb200_kernels.cubin"
txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
fn, mix = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = m.group(1)
        mix[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and fn:
        op, mods = m.group(1), m.group(2)
        key = op
        if op in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "SYNCS", "UTMALDG", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "BAR", "ELECT", "VOTE", "REDUX"):
            key = op + "".join("." + x for x in mods.split(".")[1:] if x in ("128", "64", "U8", "U16", "ARRIVE", "TRYWAIT", "EXCH", "2D", "SYNC", "ANY", "ALL"))
        mix[fn][key] += 1
groups = {"packed fp32x2": ("FFMA2", "FMUL2", "FADD2"), "scalar fp32": ("FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSETP", "FRND", "F2I", "I2F", "F2F", "MUFU", "I2FP", "F2FP"),
          "integer/logic": ("IADD3", "IADD", "IMAD", "LOP3", "SHF", "LEA", "PRMT", "ISETP", "SEL", "IMNMX", "VIADD", "VIMNMX", "IABS", "BREV", "FLO", "POPC", "SGXT"),
          "shared loads": ("LDS",), "shared stores": ("STS",), "global loads": ("LDG",), "global stores": ("STG",), "local (spill/stack)": ("LDL", "STL"),
          "TMA / bulk copy": ("UTMALDG", "UBLKCP", "UTMAPF", "UTMACCTL"), "mbarrier": ("SYNCS",), "barrier": ("BAR", "WARPSYNC"), "control": ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "NOP", "BRX", "JMP")}
for fn, c in mix.items():
    total = sum(c.values())
    print(f"== {fn}: {total} instructions")
    for g, ops in groups.items():
        n = sum(v for k, v in c.items() if k.split(".")[0] in ops)
        if n:
            detail = ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1]) if k.split(".")[0] in ops)
            print(f"   {g:22s} {n:6d} ({100.0 * n / total:4.1f} %)  {detail}")
    rest = {k: v for k, v in c.items() if not any(k.split(".")[0] in ops for ops in groups.values())}
    if rest:
        print(f"   {'other':22s} {sum(rest.values()):6d} ({100.0 * sum(rest.values()) / total:4.1f} %)  " + ", ".join(f"{k} {v}" for k, v in sorted(rest.items(), key=lambda kv: -kv[1])))
