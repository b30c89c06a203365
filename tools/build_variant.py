"""Experiment builds: tools/build_variant.py <name> <file.cu> [nvcc flags ...]

Recompiles ONE translation unit with extra flags and links it with the product objects into
mvuld_b200/csrc/build/variants/lib_<name>.so (git-ignored, travels with gpurun).  Use with MVULD_LIB=<that path>.
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvuld_b200 import _build as b  # noqa: E402


def main():
    name, unit, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    b.build()
    vdir = os.path.join(b.OBJDIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    src = os.path.join(b.CSRC, unit)
    obj = os.path.join(vdir, f"{name}_{unit[:-3]}.o")
    subprocess.run([b._nvcc(), *b.NVCC_FLAGS, *flags, "-c", src, "-o", obj], check=True)
    objs = [os.path.join(b.OBJDIR, f) for f in sorted(os.listdir(b.OBJDIR))
            if f.endswith(".o") and f != unit[:-3] + ".o"] + [obj]
    lib = os.path.join(vdir, f"lib_{name}.so")
    subprocess.run([b._nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
