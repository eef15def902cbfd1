"""Build the native pieces in-tree (so the .so files travel with the repo snapshot).

    python -m parallelparsing_b200.build            # product library + corpus tools
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parallelparsing_b200", "lib", "libppb200.so")
TOOLS = os.path.join(ROOT, "tools", "_build")


def build_library():
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> parallelparsing_b200/lib/libppb200.so"""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "parallelparsing_b200", "csrc")])
    return LIB


def build_tools():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")])
    return TOOLS


if __name__ == "__main__":
    print(build_library())
    print(build_tools())
