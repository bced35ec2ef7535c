"""Builds compile-time variants of the library into build/variants/ (git-ignored, travels with gpurun).
usage: python tools/build_variants.py name:-DFLAG=V,-DFLAG2=V2 ..."""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diplomjourney_b200 import build

def one(spec):
    name, _, flags = spec.partition(":")
    out = os.path.join("build", "variants", f"lib_{name}.so")
    build.build_library(force=True, extra_flags=[f for f in flags.split(",") if f], out=os.path.abspath(out))
    return out

with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
