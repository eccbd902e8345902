"""Per-device-function shares of the warp-stall samples of one `ncu --set full --import-source on` capture.
Input: `ncu -i <rep> --page source --csv` (SASS rows with addresses) and the library that was profiled: the device
functions k_integrate calls are not inlined, so `cuobjdump -sass` of the library's entry function lists them as
function symbols of the cubin's symbol table (`cuobjdump -elf`: value = offset inside the entry function's text);
the CALL targets found in the SASS rows give the same boundaries.
usage: ncu_function_shares.py <source.csv> <lib.so> [entry-function-substring]"""
import csv, re, subprocess, sys, collections
src, lib = sys.argv[1], sys.argv[2]
entry = sys.argv[3] if len(sys.argv) > 3 else "k_integrate"
rows = list(csv.reader(open(src)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
col = {h: i for i, h in enumerate(H)}
stall_cols = [h for h in H if h.startswith("stall_") and "(Not Issued)" not in h]
data = []
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    try:
        addr = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
    except ValueError:
        continue
    n = float(r[col["# Samples"]] or 0)
    st = {h: float(r[col[h]] or 0) for h in stall_cols}
    data.append((addr, r[col["Source"]], n, st, float(r[col["Instructions Executed"]] or 0),
                 float(r[col["L1 Wavefronts Shared"]] or 0), float(r[col["L1 Wavefronts Shared Ideal"]] or 0)))
data.sort()
base = data[0][0]
# function boundaries: targets of CALL instructions + RET positions (a function ends at a RET followed by a call target)
targets = set()
for a, s, *_ in data:
    m = re.search(r"CALL\.\w+(?:\.\w+)*\s+(0x[0-9a-f]+)", s)
    if m:
        targets.add(int(m.group(1), 16))
# names from the ELF symbol table (cuobjdump -elf): STT_FUNC symbols with their values relative to the entry
names = {}
try:
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    for line in elf.splitlines():
        m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+\d+\s+0x[0-9a-f]+\s+\$\S*" + entry + r"\S*?\$(\S+)", line)
        if m:
            names[int(m.group(1), 16)] = (m.group(3), int(m.group(2), 16))
except Exception:
    pass
print("base address", hex(base), "instructions", len(data), "call targets", len(targets), "func symbols", len(names))
tot = sum(d[2] for d in data)
# assign instructions to the nearest preceding call target (entry function = base)
starts = sorted({base} | {t for t in targets if t >= base})
def owner(a):
    import bisect
    i = bisect.bisect_right(starts, a) - 1
    return starts[i]
agg = collections.defaultdict(lambda: [0.0, collections.Counter(), 0.0, 0.0, 0.0, 0])
for a, s, n, st, ie, wf, wfi in data:
    o = owner(a)
    g = agg[o]
    g[0] += n
    for k, v in st.items():
        g[1][k] += v
    g[2] += ie; g[3] += wf; g[4] += wfi; g[5] += 1
print(f"{'start':>10} {'instrs':>6} {'share':>7}  {'inst_exec':>12} {'smem wf/ideal':>13}  top stalls")
for o, g in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if g[0] < 0.004 * tot:
        continue
    ts = sum(g[1].values()) or 1
    top = ", ".join(f"{k[6:]} {100 * v / ts:.0f}%" for k, v in g[1].most_common(4))
    nm = names.get(o - base, ("", 0))[0]
    print(f"{hex(o):>10} {g[5]:6d} {100 * g[0] / tot:6.1f}%  {g[2]:12.0f} {g[3] / max(g[4], 1):13.2f}  {top}  {nm}")
