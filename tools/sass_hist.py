"""Opcode histogram per kernel from `cuobjdump -sass` (profiles/r02_sass_opcodes.txt is its output).

usage: python tools/sass_hist.py rotationnormflow_b200/librnf_b200.so [--top N] [--range FUNC_SUBSTR START_HEX END_HEX]
"""
from __future__ import annotations

import collections
import re
import subprocess
import sys

KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "MUFU", "HMMA")


def parse(path: str):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs: dict[str, list[tuple[int, str]]] = {}
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur is not None:
            ins = m.group(2).strip()
            ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
            funcs[cur].append((int(m.group(1), 16), ins))
    return funcs


def demangle(n: str) -> str:
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main() -> None:
    path = sys.argv[1]
    top = 14
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    rng = None
    if "--range" in sys.argv:
        i = sys.argv.index("--range")
        rng = (sys.argv[i + 1], int(sys.argv[i + 2], 16), int(sys.argv[i + 3], 16))
    for name, ins in parse(path).items():
        if rng is not None:
            if rng[0] not in name:
                continue
            ins = [(a, s) for a, s in ins if rng[1] <= a < rng[2]]
        ops = collections.Counter(s.split()[0].split(".")[0] for _, s in ins)
        d = re.sub(r"\(anonymous namespace\)::", "", demangle(name))
        d = re.sub(r"\(.*", "", d)
        print(f"{d}: {len(ins)} instructions")
        keys = "  ".join(f"{k}={ops[k]}" for k in KEY if ops.get(k))
        if keys:
            print(f"    key: {keys}")
        print("    top: " + "  ".join(f"{k}={v}" for k, v in ops.most_common(top)))


if __name__ == "__main__":
    main()
