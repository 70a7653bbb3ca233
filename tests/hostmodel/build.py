"""Compile the CPU model of the device codec (test infrastructure; see host_model.cpp)."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "libhostmodel.so"


def build():
    src = HERE / "host_model.cpp"
    deps = [src] + list((HERE.parent.parent / "datacompressionfloat_b200" / "csrc").glob("*.cuh"))
    if OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    subprocess.run(["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", str(OUT), str(src)],
                   check=True)
    return OUT


if __name__ == "__main__":
    build()
