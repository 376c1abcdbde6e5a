"""Scratch: short iso3dfd run at a given radius for ncu (not part of the product).  usage: prof_tt.py N steps radius [opt=val ...]"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, '..')
from yask_b200 import capi
from yask_b200.synth import var_salt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
R = int(sys.argv[3]) if len(sys.argv) > 3 else 2
opts = dict(a.split("=") for a in sys.argv[4:])
s = capi.Solution("iso3dfd", radius=R)
s.set_overall_domain_size_vec((N, N, N))
for k, v in opts.items():
    s.set_option(k, v)
s.prepare_solution(0)
p, v = s.get_var("p"), s.get_var("v")
for t in (0, 1):
    p.fill_hash(t, 1, var_salt("p", t), -1.0, 1.0)
v.fill_hash(0, 1, var_salt("v", 0), 0.05, 0.3)
s.run_solution(0, steps - 1)
st = s.get_stats()
print("gpts", N**3 * steps / st.elapsed_secs / 1e9, "launches", st.kernel_launches)
s.close()
