"""Which kernels changed since a commit? Builds the library of <commit> and of the working tree (same nvcc flags as _build.py) and compares the
SASS of every kernel function by function. Used to show that host-side / test-infrastructure edits left the GPU-validated device code untouched.
usage: python tools/sass_diff.py <commit>"""
import hashlib, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_imu_init_b200 import _build


def funcs(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d, cur = {}, None
    for l in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip())
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in d.items()}


def build(src_root, out):
    subprocess.check_call([_build._nvcc()] + _build.NVCC_FLAGS + ["-o", out, os.path.join(src_root, "lidar_imu_init_b200", "csrc", "liinit_gpu.cu")])


commit = sys.argv[1]
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(f"git -C {ROOT} archive {commit} lidar_imu_init_b200/csrc include | tar -x -C {td}", shell=True)
    build(td, os.path.join(td, "old.so"))
    build(ROOT, os.path.join(td, "new.so"))
    a, b = funcs(os.path.join(td, "old.so")), funcs(os.path.join(td, "new.so"))
same = [k for k in a if k in b and a[k] == b[k]]
print(f"{commit}: {len(a)} kernels; working tree: {len(b)} kernels; identical SASS: {len(same)}; changed: {len([k for k in a if k in b and a[k] != b[k]])}; "
      f"new: {len([k for k in b if k not in a])}; removed: {len([k for k in a if k not in b])}")
for k in a:
    if k in b and a[k] != b[k]:
        print("  CHANGED", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip())
for k in b:
    if k not in a:
        print("  new    ", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip())
for k in a:
    if k not in b:
        print("  removed", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip())
