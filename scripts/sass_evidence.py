#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the Blackwell async/tensor instructions in the in-tree library
(cuobjdump -sass kaito_b200/libkaito_rag.so): UTCHMMA (tcgen05.mma), UTMALDG (TMA tile loads), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), LDGSTS (cp.async), SYNCS (mbarrier), plus the kernel's instruction count."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "kaito_b200/libkaito_rag.so"
out = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
pats = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "LDGSTS", "SYNCS", "ATOMS", "F2FP", "MATCH", "VOTE", "HMMA", "LDSM"]
cnt, n_ins, name = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("krag::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
    if m and name:
        n_ins[name] += 1
        ins = m.group(1)
        for p in pats:
            if re.search(r"\b" + p, ins):
                variant = re.search(r"\b(" + p + r"[\w.]*)", ins).group(1)
                cnt[name][variant] += 1
print(f"# SASS evidence: {so} (sm_100a), per kernel: instruction count, then async/tensor instruction variants and their counts")
for k in sorted(n_ins, key=lambda k: -n_ins[k]):
    if not cnt[k] or "cub::" in k:
        continue
    print(f"{k[:70]:70s} {n_ins[k]:6d}  " + "  ".join(f"{v}x{c}" for v, c in sorted(cnt[k].items())))
