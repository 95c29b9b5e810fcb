#!/usr/bin/env python
"""Static SASS opcode mix of the kernels in an object / shared library.
    python tools/sass_mix.py montecarlocuda_b200/build/kernels_vanilla.o accumulate.*VanillaId [top]
"""
import collections
import re
import subprocess
import sys


def main():
    path, pattern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, mixes = None, collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            op = m.group(1)
            base = op.split(".")[0]
            if base in ("IMAD", "MUFU") and "." in op:
                base = ".".join(op.split(".")[:2])
            mixes[name][base] += 1
    for fn, mix in mixes.items():
        if re.search(pattern, fn):
            total = sum(mix.values())
            print(f"== {fn}: {total} instructions")
            print("   " + ", ".join(f"{k} {v}" for k, v in mix.most_common(top)))


if __name__ == "__main__":
    main()
