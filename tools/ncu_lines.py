"""Development: join the per-instruction stall samples of an .ncu-rep (source page) with the line table of the
cubin inside libmpcb200.so and print the samples per source line / per function region.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel-name-substring> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
lib = os.environ.get("MPCB_LIB", os.path.join(ROOT, "safe-autonomous-driving-mpc_b200", "libmpcb200.so"))
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("mpcb_api.") and f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
starts = [(i, l) for i, l in enumerate(sass) if l.startswith(".text.")]
sel = [k for k, (i, l) in enumerate(starts) if kname in l][0]
lo = starts[sel][0]
hi = starts[sel + 1][0] if sel + 1 < len(starts) else len(sass)
cur, seq = None, []
for l in sass[lo:hi]:
    m = re.search(r'File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), m.group(2).strip(), cur))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + os.environ["KREGEX"]] if "KREGEX" in os.environ else []), capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
want = [k for k in range(len(secs) - 1) if kname.split("ILb")[0].split("_kernel")[0] in rows[secs[k]][1]
        and ("(bool)" + kname.split("ILb")[1][0] in rows[secs[k]][1] if "ILb" in kname else True)]
k0 = want[int(os.environ.get("SECTION", "0"))]
rows = rows[secs[k0]:secs[k0 + 1]]
hdr, data = rows[1], rows[2:]
assert len(data) == len(seq), (len(data), len(seq))
iS, iE, iSel = hdr.index(os.environ.get("COL", "# Samples")), hdr.index("Instructions Executed"), hdr.index("stall_selected")
by, ex, n, issued = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for k, r in enumerate(data):
    ln = seq[k][2]
    by[ln] += int(r[iS]); ex[ln] += int(r[iE]); n[ln] += 1; issued[ln] += int(r[iSel])
tot = sum(by.values())
print("total samples", tot, "instructions", len(seq), "warp-instructions executed", sum(ex.values()))
byfile = collections.Counter()
for ln, s in by.items():
    byfile[ln[0] if ln else None] += s
print("by file:", dict(byfile))
for ln, s in by.most_common(top):
    print(f"{str(ln):34s} {s:6d} {100 * s / tot:5.1f}%  execs {ex[ln]:9d}  ninstr {n[ln]:4d}  issued {issued[ln]}")
