"""Scratch: short run of a generated stencil for ncu.  usage: prof_gen.py <stencil> <n> [key=value engine options ...]"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from bench_stencils import run
print(run(sys.argv[1], int(sys.argv[2]), 2, 1, 2, [a for a in sys.argv[3:] if "=" in a]))
