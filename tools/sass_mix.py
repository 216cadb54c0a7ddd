#!/usr/bin/env python3
"""Static SASS instruction mix per function of a cubin / object (nvdisasm -c output), and a dynamic estimate for the
compiled pairing kernel from the known call counts of one pairing.  Usage: tools/sass_mix.py build/obj/pairing_st.o [kernel-substring]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disassemble(obj):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    return subprocess.check_output(["nvdisasm", "-c", os.path.join(d, cub)], text=True)


KEYS = ("IMAD.WIDE", "IMAD", "IADD3", "LOP3", "SEL", "SHF", "MOV", "LDS", "STS", "LDG", "STG", "LDL", "STL", "CALL", "BRA", "ISETP", "RET")


def mix(text):
    stats, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.match(r"^(\$?_Z[\w$]+):", line)
        if m:
            cur = m.group(1)
            stats.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(2)
            c = stats[cur]
            c["n"] += 1
            for k in KEYS:
                if op.startswith(k):
                    c[k] += 1
                    break
    return stats


if __name__ == "__main__":
    st = mix(disassemble(sys.argv[1]))
    sel = sys.argv[2] if len(sys.argv) > 2 else ""
    for f, c in st.items():
        if sel in f and c["n"] > 8:
            name = f.split("$")[-1] if "$" in f[1:] else f
            print("%-60s n=%5d wide=%4d " % (name[:60], c["n"], c["IMAD.WIDE"]) + " ".join("%s=%d" % (k, c[k]) for k in KEYS[1:] if c[k]))
