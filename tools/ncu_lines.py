"""Attribute ncu per-instruction counters (SASS source page) to CUDA source lines via nvdisasm -g line info.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled-name-prefix> [top]
env NCU_LINES_CUBIN=<cubin> NCU_LINES_SRC=<dir>: use an already built cubin / source directory (a report captured from an older tree)"""
import csv, re, subprocess, sys, collections, os
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cub = os.environ.get("NCU_LINES_CUBIN", "/tmp/liinit_lines.cubin")
if "NCU_LINES_CUBIN" not in os.environ:
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false", "-cubin", "-o", cub,
                       os.path.join(root, "lidar_imu_init_b200/csrc/liinit_gpu.cu")])
sass = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
lines = []  # per instruction: (file, line)
inside = False
cur = ("?", 0)
for l in sass:
    if l.startswith(".text."):
        inside = l.startswith(".text." + mangled)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ie, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ins = [r for r in rows[hi + 1:] if len(r) > ie and r[0].startswith("0x")]
# several kernel instances may be concatenated: take the first len(lines)
n = min(len(ins), len(lines))
if len(ins) != len(lines):
    print(f"warning: {len(ins)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
for k in range(n):
    a = agg[lines[k]]
    a[0] += int(ins[k][ie]); a[1] += int(ins[k][it]); a[2] += int(ins[k][isamp])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
src_cache = {}
def src(f, ln):
    if f not in src_cache:
        p = os.path.join(os.environ.get("NCU_LINES_SRC", os.path.join(root, "lidar_imu_init_b200/csrc")), f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    s = src_cache[f]
    return s[ln - 1].strip()[:90] if 0 < ln <= len(s) else ""
print(f"total warp instructions {tot}, samples {tots}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst  {100*a[2]/max(tots,1):5.1f}% samp  lanes {a[1]/max(a[0],1):4.1f}  {f}:{ln:<4d} {src(f, ln)}")
