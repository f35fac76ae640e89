"""Per-source-line profile of a kernel from an .ncu-rep: joins ncu's per-SASS-instruction counters with nvdisasm's inline
line info of the same cubin.  Every instruction is charged to the deepest inlining frame that lies in `focus` (default
stacb_fast.cuh), so a line that calls rotq() / qmul() carries the instructions of the callee.

    python tools/ncu_lines.py rep.ncu-rep object.o kernel-substring [focus-file] [rounds]
"""
import collections, csv, io, re, subprocess, sys, tempfile, os

rep, obj, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
focus = sys.argv[4] if len(sys.argv) > 4 else "stacb_fast.cuh"
rounds = float(sys.argv[5]) if len(sys.argv) > 5 else None

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and ksub in l)
frames, cur, insts = [], [], []
pending = []
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m:
        if pending:
            cur, pending = pending, []
        insts.append((int(m.group(1), 16), m.group(2).strip(), list(cur)))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
body = rows[2:]
iA, iS, iX = hdr.index("Address"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
assert len(body) == len(insts), (len(body), len(insts))
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "stac_mjx_b200", "csrc", focus)).read().splitlines()
ex, st = collections.Counter(), collections.Counter()
for r, (off, text, fr) in zip(body, insts):
    line = next((ln for f, ln in fr if f == focus), None)  # frames are innermost first
    key = line if line is not None else (fr[0] if fr else ("?", 0))
    ex[key] += int(r[iX]); st[key] += int(r[iS])
tot, tots = sum(ex.values()), max(1, sum(st.values()))
per = (lambda n: f"{n / rounds:8.1f}") if rounds else (lambda n: f"{n:10d}")
print(f"kernel {ksub}: {len(insts)} SASS instructions, {tot} executed warp instructions, {tots} stall samples")
print(f"{'line':>6} {'exec%':>6} {'stall%':>6} {'exec' + ('/round/warp' if rounds else ''):>12}  source")
for key, n in sorted(ex.items(), key=lambda kv: -kv[1])[:70]:
    text = src[key - 1].strip()[:110] if isinstance(key, int) else str(key)
    print(f"{str(key) if isinstance(key, int) else '-':>6} {100 * n / tot:6.2f} {100 * st[key] / tots:6.2f} {per(n):>12}  {text}")
