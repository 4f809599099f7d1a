"""cuobjdump -sass kurosiwo_b200/libkurosiwo_b200.so | python scripts/sass_tally.py > profiles/rN_sass_mnemonics.txt
Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md, 'What proves a Blackwell-native kernel')."""
import collections
import re
import subprocess
import sys
cur = None
cnt = collections.OrderedDict()
keys = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "LDSM", "ELECT"]
for l in sys.stdin:
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            op = m.group(1)
            for k in keys:
                if op.startswith(k):
                    cnt[cur][k] += 1
print("SASS mnemonic counts per kernel of kurosiwo_b200/libkurosiwo_b200.so (cuobjdump -sass, sm_100a build); kernels with none of them are omitted")
print("tcgen05 = UTCHMMA (mma) / LDTM (tcgen05.ld) / UTCBAR (commit); TMA = UTMALDG (tensor load) / UBLKCP (cp.async.bulk); mbarrier = SYNCS;")
print("warp-level tensor cores = HMMA + LDSM (mma.sync + ldmatrix); ELECT = elect.sync")
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
for fn, d in zip(names, dem):
    c = cnt[fn]
    if sum(c.values()) == 0:
        continue
    d = re.sub(r"\(.*", "", d)
    print(f"{d[:100]:102s} " + "  ".join(f"{k}={c[k]}" for k in keys if c[k]))
