"""Builds the PyTorch extension `_ggq_torch.so` in-tree with one g++ command (torch headers, pybind11; links libggq.so
through an $ORIGIN rpath).  Called by the Makefile; no JIT cache is involved, so the built file travels with the tree."""
import os
import subprocess
import sys
import sysconfig

import torch
from torch.utils import cpp_extension as ce

here = os.path.dirname(os.path.abspath(__file__))
out = os.path.join(here, sys.argv[1] if len(sys.argv) > 1 else "_ggq_torch.so")
inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}", "-I/usr/local/cuda/include"]
libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
       "-DTORCH_EXTENSION_NAME=_ggq_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
       f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
       *inc, os.path.join(here, "csrc", "torch_binding.cpp"), "-o", out,
       f"-L{here}", "-lggq", f"-L{libdir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
       "-L/usr/local/cuda/lib64", "-lcudart",
       "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{libdir}"]
print(" ".join(cmd))
subprocess.check_call(cmd)
