"""Offline view of the NVRTC-specialised sweep kernel (no GPU needed): generate the JIT source of
a workload, compile it with nvcc for sm_100a with the JIT's flags, and report registers plus the
static SASS instruction count per source-line range (phase 4 = the column loop).
    python tools/jit_sass.py [workload] [--dump]"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import OpenGoddard.optimize as api
from opengoddard_b200 import capi, tape, workloads, build

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2_goddard50"
build.build_lib()
wl = workloads.build(name, api)
ir = tape.build_ir(wl.prob, wl.obj)
nbytes, src = capi.jit_check(ir)
maxN = max(wl.prob.nodes)
nr = 0 if maxN > 128 else (maxN + 31) // 32
out = "/tmp/ogbjit"
os.makedirs(out, exist_ok=True)
packed = 1 if "--packed" in sys.argv else 0          # which of the two kernels (dense / packed output) to look at
src += ("\ntemplate __global__ void ogb_sweep_kernel<%d, %d>(OgbProb, OgbPlan, const double*, const double*, const double*, "
        "const double*, double, int, double*, double*, int, int, int, int, int, unsigned long long*, unsigned long long, int);\n" % (nr, packed))
open(out + "/ogb_jit.cu", "w").write(src)
cmd = ["nvcc", "-cubin", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-lineinfo",
       "-Xptxas", "-v", "-I", ROOT + "/opengoddard_b200/csrc", "-I", ROOT + "/include", "-o", out + "/k.cubin", out + "/ogb_jit.cu"]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    print(r.stderr); sys.exit(1)
for line in r.stderr.splitlines():
    if "registers" in line or "spill" in line:
        print(line.strip())
dis = subprocess.run(["nvdisasm", "-g", "-c", out + "/k.cubin"], capture_output=True, text=True).stdout
if "--dump" in sys.argv:
    open(out + "/k.sass", "w").write(dis)
cur = None
cnt = collections.Counter()
ops = collections.Counter()
lines = open(ROOT + "/opengoddard_b200/csrc/ogb_sweep.cuh").read().splitlines()
p4 = next(i for i, l in enumerate(lines) if "---- phase 4" in l) + 1
p4end = next(i for i, l in enumerate(lines) if "all warps are done reading this item" in l) + 1
total = 0
for l in dis.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f, ln, rest = os.path.basename(m.group(1)), int(m.group(2)), m.group(3)
        # an inlined function: attribute to the outermost call site in ogb_sweep.cuh if present
        cur = (f, ln)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', l)
    if m and cur:
        total += 1
        ins = m.group(1)
        cnt[cur] += 1
        if cur[0] == "ogb_sweep.cuh" and p4 <= cur[1] <= p4end:
            op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
            ops[op.split(".")[0]] += 1
print("total SASS instructions", total)
in4 = sum(v for (f, ln), v in cnt.items() if f == "ogb_sweep.cuh" and p4 <= ln <= p4end)
print("attributed to ogb_sweep.cuh phase-4 lines %d-%d: %d" % (p4, p4end, in4))
print("  mix:", ", ".join("%s %d" % kv for kv in ops.most_common(14)))
core = sum(v for (f, ln), v in cnt.items() if f == "ogb_core.h")
print("attributed to ogb_core.h: %d" % core)
