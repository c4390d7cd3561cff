"""Experimental A/B build of libdkb.so (2 filter bits only: a third of the kernels).
Usage: python scripts/ab_build.py NAME [-DFLAG ...]  ->  ab/libdkb_NAME.so ; run with DKB_LIBRARY=ab/libdkb_NAME.so"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from denovo_kmer_b200 import build  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "ab"), exist_ok=True)
out = os.path.join(ROOT, "ab", f"libdkb_{name}.so")
print(build.build(out=out, extra=["-DDKB_AB_BUILD"] + flags, obj_dir=os.path.join(ROOT, "build", f"ab_{name}")))
