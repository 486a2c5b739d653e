"""Build an experimental variant of the library for A/B measurements on the GPU box (one gpurun call, several libraries):

    python tools/build_variant.py TAG -DTC4_PROD2=0 -DTC4_PUB=1 ...   ->  fn_ssl_b200/variants/libfnssl_b200_TAG.so

Only the sources named in VARIANT_SRC (default lstm_tc4.cu) are recompiled with the extra defines; every other object is the regular build's.  Select a variant at run
time with FNSSL_B200_LIB=<path> (fn_ssl_b200/_lib.py; development / profiling only -- the product loads libfnssl_b200.so).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fn_ssl_b200 import build as B  # noqa: E402


def main():
    tag, defines = sys.argv[1], sys.argv[2:]
    B.build()
    vdir = os.path.join(B.HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    srcs = os.environ.get("VARIANT_SRC", "lstm_tc4.cu").split(",")      # the sources the defines apply to
    regs, vobjs = [], []
    for src in srcs:
        obj = os.path.join(B.BUILD, f"variant_{src[:-3]}_{tag}.o")
        cmd = [B._nvcc()] + B.NVCC_FLAGS + defines + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise SystemExit(res.stdout + res.stderr)
        regs += [l for l in res.stderr.splitlines() if "registers" in l or "spill" in l]
        vobjs.append(obj)
    objs = [os.path.join(B.BUILD, s[:-3] + ".o") for s in B._sources() if s not in srcs] + vobjs
    lib = os.path.join(vdir, f"libfnssl_b200_{tag}.so")
    subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
    worst = max((int(l.split("Used ")[1].split()[0]) for l in regs if "Used " in l), default=0)
    spills = sorted({l.strip() for l in regs if "spill" in l and "0 bytes spill stores" not in l})
    print(f"{lib}: defines {defines}, max registers {worst}, spills: {spills or 'none'}")


if __name__ == "__main__":
    main()
