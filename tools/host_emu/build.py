"""Build the HOST EMULATION of the CUDA-core training kernels (test infrastructure; see tools/host_emu/common.cuh).

    python tools/host_emu/build.py  ->  tools/host_emu/_build/libfnssl_emu.so

The .cu sources are taken from fn_ssl_b200/csrc as they are, with two textual rewrites: CUDA launch syntax -> emu_launch(...),
and `extern __shared__` arrays -> fixed-size static arrays.  `#include "common.cuh"` resolves to the shim in this directory."""
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "fn_ssl_b200", "csrc")
OUT = os.path.join(HERE, "_build")
SOURCES = ["lstm_simt.cu", "lstm_train.cu", "conv_train.cu", "train.cu", "head.cu", "spatialnet.cu"]
LIB = os.path.join(OUT, "libfnssl_emu.so")

LAUNCH = re.compile(r"^(\s*)([\w:]+(?:<[^<>;]*>)?)<<<(.+?),\s*([^,]+?),\s*([^,]+?),\s*([^,]+?)>>>\((.*?)\);", re.M | re.S)
DYN_SHARED = re.compile(r"extern __shared__ (?:__align__\(16\) )?(\w+) (\w+)\[\];")


def transform(text: str) -> str:
    text, n = LAUNCH.subn(lambda m: f'{m.group(1)}emu_launch("{m.group(2)}", [&]() {{ {m.group(2)}({m.group(7)}); }}, dim3({m.group(3)}), dim3({m.group(4)}));', text)
    assert "<<<" not in text, "an unconverted kernel launch is left"
    text = DYN_SHARED.sub(lambda m: f"static __attribute__((aligned(16))) {m.group(1)} {m.group(2)}[57344];", text)
    assert "extern __shared__" not in text
    text = re.sub(r'asm\("ex2\.approx\.ftz\.f32 %0, %1;" : "=f"\((\w+)\) : "f"\((\w+)\)\);', r"\1 = exp2f(\2);", text)   # the one PTX line
    assert "asm(" not in text and "asm volatile" not in text, "inline PTX cannot be emulated"
    return text


def build() -> str:
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    srcs = []
    for name in SOURCES:
        text = transform(open(os.path.join(CSRC, name)).read())
        h.update(text.encode())
        dst = os.path.join(OUT, name[:-3] + "_emu.cpp")
        srcs.append((dst, text))
    for dep in (os.path.join(HERE, "common.cuh"), os.path.join(ROOT, "include", "fnssl_b200.h")):
        h.update(open(dep, "rb").read())
    stamp = os.path.join(OUT, "stamp")
    if os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return LIB
    for dst, text in srcs:
        with open(dst, "w") as fh:
            fh.write(text)
    cmd = ["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", "-w", "-I", HERE, "-o", LIB] + [d for d, _ in srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host emulation build failed:\n" + res.stderr[-4000:])
    with open(stamp, "w") as fh:
        fh.write(h.hexdigest())
    return LIB


if __name__ == "__main__":
    print(build())
