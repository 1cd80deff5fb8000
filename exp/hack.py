"""Timing experiments: compile a hand-edited copy of a generated kernel into an alternative kernel
cache under the SAME key, so that the runtime loads it (GAAST_TEST_HOOKS=1 GAAST_KERNEL_CACHE=exp/<name>;
such kernels report origin=override-unverified in gaast_plan_last_kernel).  Results
of such kernels may be wrong: timing only.
    python exp/hack.py <key> <name> <transform>"""
import os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key, name, tf = sys.argv[1:4]
src = open(os.path.join(ROOT, "gaast_b200/kernel_cache", key + ".cu")).read()
lines = src.split("\n")
def body_only(pred):
    out, inside = [], False
    for l in lines:
        if "gaast_eval(" in l and "__global__" in l: inside = True
        if inside and pred(l): continue
        out.append(l)
    return out
if tf == "none":
    out = lines
elif tf == "no_tmem":      # drop the stash writes, the direct updates and the tile-end absorption
    out = body_only(lambda l: re.match(r"\s+(tm_put\(tb \+ \d+u, d_hsum|tm_add\(|tm_acc\d+<)", l))
elif tf == "no_tmem_no_guard":
    out = [l.replace("if (active) ", "") for l in body_only(lambda l: re.match(r"\s+(tm_put\(tb \+ \d+u, d_hsum|tm_add\(|tm_acc\d+<)", l))]
elif tf == "no_stash_absorb":  # keep stash writes, drop tile-end absorption
    out = body_only(lambda l: re.match(r"\s+(tm_acc\d+<)", l))
elif tf == "c3_no_loads":   # cfg3: no TMA copies, no waits, no global loads (compute + stores only)
    out = []
    for l in lines:
        if re.match(r"\s+tma_row\(", l) or "mbar_expect_tx(stage_bar" in l or "mbar_wait(stage_bar" in l:
            continue
        m = re.match(r"(\s+const D v(\d+) = )d_load\(.*\);", l)
        if m:
            l = f"{m.group(1)}__longlong_as_double(e + {m.group(2)}LL) * 1e-300;"
        out.append(l)
elif tf == "c3_no_stores":  # cfg3: one store per output coset instead of 16
    out = []
    for l in lines:
        m = re.match(r"(\s+)ro\[dense_out\[\(oh << 4\) \+ (\d+)\]\] = q(\d+);", l)
        if m:
            if m.group(2) == "0":
                l = m.group(1) + "ro[dense_out[(oh << 4)]] = " + " + ".join(f"q{i}" for i in range(16)) + ";"
            else:
                continue
        out.append(l)
elif tf == "c3_no_sign":    # cfg3: right operand used as loaded (no per-tile sign flip)
    out = [re.sub(r"flip_sign\((xs_ldd\([^)]*\)\)), sg\)", r"\1", l) for l in lines]
elif tf == "c3_compute_only":
    out = []
    for l in lines:
        if re.match(r"\s+tma_row\(", l) or "mbar_expect_tx(stage_bar" in l or "mbar_wait(stage_bar" in l:
            continue
        m = re.match(r"(\s+const D v(\d+) = )d_load\(.*\);", l)
        if m:
            l = f"{m.group(1)}__longlong_as_double(e + {m.group(2)}LL) * 1e-300;"
        m = re.match(r"(\s+)ro\[dense_out\[\(oh << 4\) \+ (\d+)\]\] = q(\d+);", l)
        if m:
            if m.group(2) == "0":
                l = m.group(1) + "ro[dense_out[(oh << 4)]] = " + " + ".join(f"q{i}" for i in range(16)) + ";"
            else:
                continue
        out.append(l)
else:
    raise SystemExit("unknown transform")
d = os.path.join(ROOT, "exp", name)
if not os.path.isdir(d):
    shutil.copytree(os.path.join(ROOT, "gaast_b200/kernel_cache"), d)
p = os.path.join(d, key + ".hack.cu")
open(p, "w").write("\n".join(out))
r = subprocess.run(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-std=c++17", "-Xptxas", "-v",
                    "-o", os.path.join(d, key + ".cubin"), p], capture_output=True, text=True)
print(name, tf, r.returncode, [l for l in r.stderr.split("\n") if "registers" in l or "spill" in l][-2:])
