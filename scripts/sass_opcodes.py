#!/usr/bin/env python
"""SASS evidence that the kernels are Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"):
opcode counts of `cuobjdump -sass vggt_qwen3_b200/libvq_b200.so`, per kernel family and in total.
usage: python scripts/sass_opcodes.py > profiles/rNN_sass_opcodes.txt   (no GPU needed)"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "vggt_qwen3_b200" / "libvq_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "HGMMA", "SYNCS", "MUFU.EX2",
         "MUFU.TANH", "FFMA2", "FADD2", "RED", "ATOM", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    fam = collections.defaultdict(collections.Counter)
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = ("gemm_kernel" if "gemm_kernel" in name else "perceiver_attn (one-tile fwd / bwd)" if "perceiver_attn" in name
                   else "attention_fwd_bounded" if "attention_fwd_bounded" in name
                   else "attention_fwd_wide" if "attention_fwd_wide" in name else "attention_fwd (split-KV)" if "attention_fwd" in name
                   else "row / elementwise kernels")
            fam[cur]["<kernels>"] += 1
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    fam[cur][w] += 1
            fam[cur]["<instructions>"] += 1
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  (sm_100a; counts of instructions, static)")
    print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, legacy mma.sync -> HMMA (must be 0)")
    total = collections.Counter()
    for f, c in sorted(fam.items()):
        print(f"\n## {f}: {c['<kernels>']} kernel instantiations, {c['<instructions>']} instructions")
        for w in WATCH:
            if c[w]:
                print(f"{w:10s} {c[w]}")
        total.update(c)
    print(f"\n## total: {total['<kernels>']} kernels, {total['<instructions>']} instructions")
    for w in WATCH:
        print(f"{w:10s} {total[w]}")
    if total["HMMA"] or total["HGMMA"]:
        print("WARNING: legacy tensor-core instructions present", file=sys.stderr)


if __name__ == "__main__":
    main()
