import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, ".")
os.environ["LRB_TC_TIMELINE"] = "1"
import logreg_b200 as lr
from logreg_b200 import _native as N
prob = lr.Problem(); bt = prob.gen_synthetic(1_000_000, 64, mode="fp32")
B = bt + 0.1 * np.random.RandomState(1).randn(512, 64)
prob.eval_many(B)
eta = np.zeros((512, prob._lib.lrb_tc_tile_rows()), dtype=np.float32)
prob._ck(prob._lib.lrb_debug_tc_eta(prob._h, N.as_dp(np.ascontiguousarray(B)), 512, eta.ctypes.data_as(C.POINTER(C.c_float))))
