"""Runs one of the reference's experiment files unchanged (but for T) on the engine through abm_b200.compat.
usage: EXPERIMENT_NAME=dropin python run_ref_exp.py <experiment file> <T>   (cwd: a folder holding dropin.env)"""
import sys, time
sys.path.insert(0, '/root/repo')
import abm_b200.compat
path, T = sys.argv[1], int(sys.argv[2])
src = open(path).read()
assert src.count('Constant("T", 25000)') == 1
t0 = time.time()
exec(compile(src.replace('Constant("T", 25000)', f'Constant("T", {T})'), path, "exec"))
print("EXPERIMENT_DONE in %.1f s" % (time.time() - t0))
