import time, numpy as np, sys
sys.path.insert(0, ".")
import oracle
from clustassess_b200 import ecs, synth, _capi
L, _ = synth.config("C4")
m, n = L.shape
t = time.time(); S = ecs.element_sim_matrix(list(L)); t1 = time.time() - t
ts = []
for _ in range(3):
    t = time.time(); S = ecs.element_sim_matrix(list(L)); ts.append(time.time() - t)
pairs = m * (m - 1) // 2
print("C4 element_sim_matrix: first %.3f s, then %s s; pair*cells/s = %.4g (e2e incl. Python list handling)" % (t1, ["%.3f" % x for x in ts], pairs * n / min(ts)))
print("profile", _capi.profile())
rng = np.random.default_rng(0)
err = 0
for _ in range(40):
    i, j = sorted(rng.choice(m, 2, replace=False))
    err = max(err, abs(S[i, j] - oracle.disjoint_ecs(L[i], L[j]).mean()))
print("max |err| over 40 sampled entries", err, "symmetric", np.array_equal(S, S.T), "diag", np.all(np.diag(S) == 1))
# weighted consistency with matrix on C3 full
L3, w3 = synth.config("C3")
t = time.time(); r = ecs.weighted_element_consistency(list(L3), weights=w3, calculate_sim_matrix=True); t3 = time.time() - t
t = time.time(); r = ecs.weighted_element_consistency(list(L3), weights=w3, calculate_sim_matrix=True); t3 = time.time() - t
ecc, sim = oracle.weighted_element_consistency(L3, w3, calculate_sim_matrix=True)
print("C3 weighted+matrix: %.4f s; err ecc %.3g sim %.3g" % (t3, np.abs(r["ecc"] - ecc).max(), np.abs(r["ecs_matrix"] - sim).max()), _capi.profile())
