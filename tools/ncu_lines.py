"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_lines.py REPORT [TOPN]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr = None
lines = []
for raw in out.split("\n"):
    if raw.startswith('"Line No"'):
        hdr = next(csv.reader([raw]))
        continue
    m = re.match(r'^"(\d+)","', raw)
    if hdr is None or not m:
        continue
    # the source text may hold quotes and commas: take the numeric columns from the right
    parts = raw.rstrip().rstrip('"').split('","')
    nfix = len(hdr) - 2
    if len(parts) < nfix + 2:
        continue
    tailf = parts[-nfix:]
    text = '","'.join(parts[1:len(parts) - nfix])
    lines.append([m.group(1), text] + tailf)
ix = {}
for i, n in enumerate(hdr):
    ix.setdefault(n, i)
I, S = ix["Instructions Executed"], ix["# Samples"]
tot_i = sum(int(r[I]) for r in lines)
tot_s = sum(int(r[S]) for r in lines)
print("total inst %d samples %d" % (tot_i, tot_s))
src = open("/root/repo/handposeestimation-with-3d-cnns_b200/csrc/dtsdf_kernels.cu").read().split("\n")
# regions by function
regions = []
cur = "?"
names = {}
for n, text in enumerate(src, 1):
    m = re.match(r"^(?:__device__|__global__|template|static).*", text)
    m2 = re.search(r"\b([a-zA-Z_0-9]+)\s*\((?:const|float|int|void|unsigned|FrameShared|Frame|BatchArgs)", text)
    if (text.startswith("__device__") or text.startswith("__global__")) and m2:
        cur = m2.group(1)
    names[n] = cur
agg = {}
for r in lines:
    k = names.get(int(r[0]), "?")
    a = agg.setdefault(k, [0, 0])
    a[0] += int(r[I]); a[1] += int(r[S])
print("by function (inst%, samples%):")
for k, (a, b) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("  %-22s %6.1f%% %6.1f%%" % (k, 100.0 * a / tot_i, 100.0 * b / max(tot_s, 1)))
print("top lines by instructions:")
for r in sorted(lines, key=lambda r: -int(r[I]))[:topn]:
    print("  %5s %6.2f%% %6.2f%%  %s" % (r[0], 100.0 * int(r[I]) / tot_i, 100.0 * int(r[S]) / max(tot_s, 1), r[1].strip()[:110]))
