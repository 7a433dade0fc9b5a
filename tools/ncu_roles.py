"""Break an `ncu --set full --import-source on` capture of the tensor-core kernel down by warp role
(producer / epilogue wait / epilogue body / MMA issuers) and stall reason, from the SASS-correlated source page.
    python tools/ncu_roles.py gpurun_out/prof.ncu-rep
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = idx = cur = None
seen = {}
for r in rows:
    if len(r) == 2:
        continue
    if r[0] == "Line No":
        hdr = r
        idx = {h: j for j, h in enumerate(hdr)}
        continue
    if r[0] != "":
        cur = int(r[0])
        continue
    if r[2] in ("...", ""):
        continue
    try:
        n = int(r[idx["# Samples"]])
    except ValueError:
        continue
    a = int(r[2], 16)
    if a in seen:
        seen[a][3].append(cur)
        continue
    st = {h: int(r[j] or 0) for h, j in idx.items() if h.startswith("stall_") and "Not Issued" not in h}
    seen[a] = [r[3].strip(), n, st, [cur], int(r[idx["Instructions Executed"]] or 0)]
addrs = sorted(seen)
base = addrs[0]
tot = sum(seen[a][1] for a in addrs)
# role boundaries: the producer's bulk copy, the epilogue's first LDTM, the first MMA
def first(pred):
    for a in addrs:
        if pred(seen[a][0]):
            return a
    return None
a_prod = first(lambda s: "UBLKCP" in s)
a_ldtm = first(lambda s: s.startswith("LDTM"))
a_mma = first(lambda s: "UTCHMMA" in s or "UTCQMMA" in s)
# the epilogue's chunk wait is the hottest TRYWAIT before the first LDTM
waits = [a for a in addrs if "TRYWAIT" in seen[a][0] and a_prod < a < a_ldtm]
a_wait = max(waits, key=lambda a: seen[a][1] + seen.get(a + 16, ["", 0])[1])
# issuer region starts at the first TRYWAIT after the last STG/epilogue code before the MMAs: approximate with the last
# SYNCS.ARRIVE of the epilogue tail
mma_start = max(a for a in addrs if a < a_mma and ("BAR.SYNC" in seen[a][0] or "ATOMG" in seen[a][0] or "STG" in seen[a][0]))
regions = [("setup", base, a_prod - 0x400), ("producer", a_prod - 0x400, a_prod + 0x200), ("epilogue: chunk wait", a_wait - 0x40, a_wait + 0x60),
           ("epilogue: body", a_wait + 0x60, mma_start + 0x10), ("MMA issuers / forwarders", mma_start + 0x10, addrs[-1] + 16)]
print(f"total samples {tot}")
for name, lo, hi in regions:
    sel = [a for a in addrs if lo <= a < hi]
    n = sum(seen[a][1] for a in sel)
    ie = sum(seen[a][4] for a in sel)
    st = {}
    for a in sel:
        for k, v in seen[a][2].items():
            if v:
                st[k] = st.get(k, 0) + v
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7])
    print(f"{name:28s} {n:7d} {100 * n / tot:5.1f}%  inst {ie:11d}  | {top}")
if "-v" in sys.argv:
    for a in addrs:
        if seen[a][1] >= 200:
            print(f"{a - base:6x} L{max(seen[a][3]):4d} {seen[a][1]:6d} {seen[a][4]:9d}  {seen[a][0][:90]}")
