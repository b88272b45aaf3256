"""Aggregate ncu warp-stall samples per CUDA source line (no GPU needed).
    python scripts/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
Uses `ncu --page source --csv` (SASS addresses + samples) and `nvdisasm -g` (SASS offset -> file:line)."""
import csv, io, re, subprocess, sys

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infunc = {}, None, False
for ln in dis.splitlines():
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infunc = kname in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and infunc:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if len(r) > 2 and r[0] == "Address"][0]
hdr = rows[hi]
si = hdr.index("Warp Stall Sampling (All Samples)")
base = None
agg, sass = {}, {}
for r in rows[hi + 1:]:
    if len(r) <= si or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    v = float(r[si] or 0)
    key = line_of.get(addr - base)
    agg[key] = agg.get(key, 0) + v
    if v:
        sass.setdefault(key, []).append((v, r[1].strip()))
tot = sum(agg.values()) or 1
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    text = ""
    if k:
        try:
            if k[0] not in src:
                import glob
                src[k[0]] = open(glob.glob(f"/root/repo/tnsp_b200/csrc/{k[0]}")[0]).read().splitlines()
            text = src[k[0]][k[1] - 1].strip()[:110]
        except Exception:
            pass
    hot = max(sass.get(k, [(0, "")]))[1][:50]
    print(f"{100 * v / tot:5.1f}%  {k}  {text}   [{hot}]")
