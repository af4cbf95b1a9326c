"""Join an `ncu --page source --csv` dump with `nvdisasm --print-line-info` of the same cubin: executed warp
instructions and stall samples per source line (the .ncu-rep of a gpurun call carries no CUDA source view).
    python tools/ncu_by_line.py <src.csv> <nvdisasm.txt> <mangled kernel name> [top]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
base = int(data[0][ix["Address"]], 16)
ex = {int(r[ix["Address"]], 16) - base: (int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)) for r in data}
kern = sys.argv[3]
line = None; on = False
by = collections.defaultdict(lambda: [0, 0, 0])
for l in open(sys.argv[2]):
    if l.startswith("//---") and ".text." in l:
        on = (".text." + kern + " ") in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m and line:
        off = int(m.group(1), 16)
        if off in ex:
            by[line][0] += ex[off][0]; by[line][1] += ex[off][1]; by[line][2] += 1
ti = sum(v[0] for v in by.values()); ts = sum(v[1] for v in by.values())
print(f"matched warp instructions {ti}, samples {ts}")
src = {}
for (f, n), v in sorted(by.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    if f not in src:
        try: src[f] = open("mansy_immersivevideostreaming_b200/csrc/" + f).read().split("\n")
        except OSError: src[f] = []
    text = src[f][n - 1].strip()[:80] if n - 1 < len(src[f]) else ""
    print(f"{100*v[0]/ti:5.1f}% inst {100*v[1]/ts:5.1f}% stall  {v[2]:4d} sass  {f}:{n:<4d} {text}")
