"""TEST INFRASTRUCTURE ONLY: build tests/host_twin/libctk_host_twin.so with g++ (no CUDA needed)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libctk_host_twin.so")


def build():
    src = os.path.join(HERE, "twin.cpp")
    deps = [src] + [os.path.join(HERE, "..", "..", "control_toolkit_b200", "csrc", f) for f in ("ctk_math.cuh", "ctk_derive.h", "ctk_ode_scaled.cuh", "ctk_args.cuh")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", OUT], check=True)
    return OUT


if __name__ == "__main__":
    build()
