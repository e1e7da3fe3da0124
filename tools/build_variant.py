"""Build a variant of the product library with extra nvcc flags for ONE source (A/B of a compile-time kernel variant):

    python tools/build_variant.py <name> <source.cu> -DFLAG=1 ...   ->  tools/_variants/lib_<name>.so

The other objects are the product build's (cryovit_b200/lib/obj). Load it with CRYOVIT_B200_LIB=<path>."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cryovit_b200 import build  # noqa: E402

name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
build.build()
out = ROOT / "tools" / "_variants"
out.mkdir(exist_ok=True)
obj = out / f"{name}_{src}.o"
subprocess.check_call([build._nvcc(), *build.NVCC_FLAGS, *flags, "-c", str(build.CSRC / src), "-o", str(obj)])
others = [str(build.LIBDIR / "obj" / (s + ".o")) for s in build.SOURCES if s != src]
lib = out / f"lib_{name}.so"
subprocess.check_call([build._nvcc(), "-shared", "-o", str(lib), str(obj), *others, "-gencode", "arch=compute_100a,code=sm_100a"])
obj.unlink()
print(lib)
