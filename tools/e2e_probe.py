"""debug: why is the device time of an evaluation through the model API larger than a bare engine evaluation?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpy_b200
from gpy_b200 import _ffi
from bench import synthetic, theta_for_step

N, D = 16384, 8
X, Y = synthetic(N, D)
eng = _ffi.Engine(0)
eng.set_data(X, Y)

def show(tag):
    st = eng.stats()
    print("%-34s total %.2f kbuild %.3f sweep %.2f update %.2f lauum %.2f solve %.3f tries %s" % (
        tag, st["total_ms"], st["kbuild_ms"], st["sweep_ms"], st["update_ms"], st["lauum_ms"], st["solve_ms"], st.get("tries")), flush=True)

for s in range(4):
    eng.exact_eval("rbf", True, *theta_for_step(D, s)); show("bare eval %d" % s)
for s in range(3):
    eng.set_data(X.copy(), Y.copy()); eng.exact_eval("rbf", True, *theta_for_step(D, s)); show("set_data + eval %d" % s)
for s in range(3):
    time.sleep(0.2); eng.exact_eval("rbf", True, *theta_for_step(D, s)); show("sleep 0.2 s + eval %d" % s)
m = gpy_b200.GPRegression(X, Y, gpy_b200.RBF(D, ARD=True), noise_var=0.01, device=0, engine=eng)
show("model constructor (default theta)")
m.update_model(False)
for s in range(3):
    m.set_XY(X.copy(), Y.copy()); m.set_theta(*theta_for_step(D, s)); show("model set_XY + set_theta %d" % s)
for s in range(3):
    m.set_theta(*theta_for_step(D, s)); show("model set_theta only %d" % s)
for s in range(3):
    eng.exact_eval("rbf", True, *theta_for_step(D, s)); show("bare eval again %d" % s)
print("lml model %.9f" % m.log_likelihood())
