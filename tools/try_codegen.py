import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, os, subprocess, tempfile
from gaast_b200 import workloads as W, _lib as L
from gaast_b200.device import Plan
name, ept, variant = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
w = W.WORKLOADS[name]
d = tempfile.mkdtemp()
os.environ["GAAST_KERNEL_CACHE"] = d
os.environ["GAAST_TEST_HOOKS"] = "1"
plan = Plan(None, W.specialize(w))
plan.set_tuning(ept, variant)
info = plan.precompile(w.broadcast_mask(), L.ARITH_FMA, len(sys.argv) > 4, True)
cub = info.split("key=")[1].split()[0] + ".cubin"
out = subprocess.run(["cuobjdump", "--dump-resource-usage", os.path.join(d, cub)], capture_output=True, text=True).stdout
src = open(os.path.join(d, cub[:-6] + ".cu")).read().split("\n")[1]
print(name, "ept", ept, "variant", variant, [l.strip().split(" SHARED")[0] for l in out.split("\n") if "REG" in l], src[:200])
