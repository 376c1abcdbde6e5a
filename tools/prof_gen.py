"""Scratch: short run of a generated stencil for ncu."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from bench_stencils import run
print(run(sys.argv[1], int(sys.argv[2]), 2, 1, 2))
