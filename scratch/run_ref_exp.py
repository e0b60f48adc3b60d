"""Runs one of the reference's experiment files unchanged (but for T) on the engine through abm_b200.compat.
usage: EXPERIMENT_NAME=dropin python run_ref_exp.py <experiment file> <T>   (cwd: a folder holding dropin.env)"""
import sys, time
sys.path.insert(0, '/root/repo')
import abm_b200.compat
path, T = sys.argv[1], int(sys.argv[2])
import re
src = open(path).read()
src, n = re.subn(r'Constant\("T",\s*[0-9]+\)', f'Constant("T", {T})', src)
assert n >= 1, "no Constant(\"T\", ...) in this file"
t0 = time.time()
exec(compile(src, path, "exec"))
print("EXPERIMENT_DONE in %.1f s" % (time.time() - t0))
