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
# Role of an instruction = where its CUDA source line lies between the role markers of csrc/tc_kernel.cuh (inlined helpers -- waits,
# fences, split arithmetic -- are attributed to the instruction's LAST source line, i.e. the call site inside the role).
import os
import re

src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "21cmvae_b200", "csrc", "tc_kernel.cuh")).read().splitlines()
marks = []
for i, line in enumerate(src, 1):
    m = re.search(r"// =+ (producer|follower of a pair|MMA issuers|prologue warp|epilogue warps)", line)
    if m:
        marks.append((i, m.group(1)))
    if "---- teardown" in line:
        marks.append((i, "teardown"))
    if "mbar_wait(bar_chunk_full" in line:
        chunk_wait_line = i
first_role = marks[0][0]


def role_of(lines):
    # an instruction may carry several source lines (inlining): the largest one inside the kernel body is the call site
    inside = [l for l in lines if first_role <= l < marks[-1][0]]
    if not inside:
        return None  # an inlined helper (barrier wait, split arithmetic ...): the caller decides from the neighbouring instructions
    l = max(inside)
    if l == chunk_wait_line:
        return "epilogue: wait for an accumulator chunk"
    name = "setup / teardown"
    for ln, nm in marks:
        if l >= ln:
            name = nm
    return {"producer": "weight producer", "follower of a pair": "ring forwarder (follower CTA)", "MMA issuers": "MMA issuers",
            "prologue warp": "prologue warp", "epilogue warps": "epilogue: body", "teardown": "setup / teardown"}[name]


print(f"total samples {tot}")
agg = {}
prev_role = "setup / teardown"
prev_txt = None
for a_ in addrs:
    txt_, n, st, lines, ie = seen[a_]
    r_ = role_of(lines)
    if r_ is None:  # inlined helper code sits inside its caller's address range
        r_ = prev_role if prev_role != "epilogue: wait for an accumulator chunk" or "SYNCS" in txt_ or "BRA" in txt_ else "epilogue: body"
    prev_role = r_
    # the epilogue's mbarrier waits (for the next accumulator chunk; the try_wait and the branch that closes its loop)
    if r_.startswith("epilogue"):
        is_wait = "SYNCS.PHASECHK" in txt_ or (prev_txt is not None and "SYNCS.PHASECHK" in prev_txt and "BRA" in txt_)
        r_ = "epilogue: wait for an accumulator chunk" if is_wait else "epilogue: body"
    prev_txt = txt_
    g = agg.setdefault(r_, [0, 0, {}])
    g[0] += n
    g[1] += ie
    for k, v in st.items():
        if v:
            g[2][k] = g[2].get(k, 0) + v
for name, (n, ie, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7])
    print(f"{name:42s} {n:7d} {100 * n / tot:5.1f}%  inst {ie:11d}  | {top}")
if "-v" in sys.argv:
    for a_ in addrs:
        if seen[a_][1] >= 200:
            print(f"{a_ - base:6x} L{max(seen[a_][3]):4d} {seen[a_][1]:6d} {seen[a_][4]:9d}  {seen[a_][0][:90]}")
