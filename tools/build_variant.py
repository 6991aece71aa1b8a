"""Developer tool: build a VARIANT of the library with extra nvcc flags on some translation units (experiments with
build-time switches such as -DZK_AFF_MINB=4), next to the shipped one.  The objects of the other units are reused.
usage: python tools/build_variant.py NAME "unit1,unit2" -DFLAG[=v] ...   ->  zikkurat_algebra_b200/lib/variant_NAME.so
(to try it on a GPU box: cp zikkurat_algebra_b200/lib/variant_NAME.so zikkurat_algebra_b200/lib/libzkmsm_b200.so there)"""
import concurrent.futures, os, subprocess, sys
sys.path.insert(0, ".")
from zikkurat_algebra_b200 import build as B
name, units, flags = sys.argv[1], sys.argv[2].split(","), sys.argv[3:]
B.build()
nvcc = B._nvcc()
vdir = os.path.join(B.OBJDIR, "variant_" + name)
os.makedirs(vdir, exist_ok=True)
def cc(u):
    obj = os.path.join(vdir, u + ".o")
    r = subprocess.run([nvcc, *B.NVCC_FLAGS, *flags, "-Xptxas", "-v", "-c", os.path.join(B.CSRC, u + ".cu"), "-o", obj], capture_output=True, text=True)
    if r.returncode: raise SystemExit(r.stderr)
    open(os.path.join(vdir, u + ".ptxas.txt"), "w").write(r.stderr)
    return obj
with concurrent.futures.ThreadPoolExecutor(8) as ex:
    vobjs = dict(zip(units, ex.map(cc, units)))
objs = [vobjs.get(u, os.path.join(B.OBJDIR, u + ".o")) for u in B.UNITS]
out = os.path.join(B.LIBDIR, f"variant_{name}.so")
r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs, "-cudart", "static", "-Xcompiler", "-pthread"], capture_output=True, text=True)
if r.returncode: raise SystemExit(r.stderr)
print(out)
