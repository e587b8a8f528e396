"""Builds the test-only host simulation of the kernel core (see hostsim.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libss_hostsim.so")


def build():
    src = os.path.join(HERE, "hostsim.cpp")
    core = os.path.join(HERE, "..", "..", "skillshot_learning_b200", "csrc", "ss_env_core.cuh")
    newest = max(os.path.getmtime(src), os.path.getmtime(core))
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma",
                        "-Wno-unknown-pragmas", "-o", SO, src, "-lm"], check=True)
    return SO
