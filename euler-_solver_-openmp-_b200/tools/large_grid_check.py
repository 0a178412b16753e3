import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
ldm = load_package()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t = time.time()
props = ldm.Properties.Generate(1, 0.10062893578616353, 0.08, 0.4, 0.2, 0.001, G)
print("generate", round(time.time() - t, 2), "s  G", props.G, "K", props.K, "mM", props.MaxNumForM, flush=True)
t = time.time()
eq = ldm.Equation(props)
print("create", round(time.time() - t, 2), "s", flush=True)
g = props.MomentsGrid
st = np.empty(1 + G * G); st[0] = 200.0
st[1:] = (200.0 ** 2 * (1 + 0.5 * np.exp(-(g[:, None] ** 2 + g[None, :] ** 2) / (2 * 0.08 ** 2)))).ravel()
eq.upload(st)
t = time.time(); eq.advance(2, 0.1); eq.synchronize(); print("first steps", round(time.time() - t, 2), "s", flush=True)
eq.advance(5, 0.1); eq.synchronize()
print("ms/step", eq.last_elapsed_ms() / 5)
fin = eq.download(); print("finite", np.isfinite(fin).all(), fin[0])
# cross-check a few nodes of the integrals against the sample-by-sample path
d = ldm.Equation(props, algorithm=ldm.ALGO_DIRECT)
t = time.time(); (w1, m1) = d.integrals(st); print("direct integrals", round(time.time() - t, 2), "s")
(w2, m2) = eq.integrals(st)
print("max rel dev MC prefix vs direct", np.max(np.abs(m1 - m2) / np.abs(m1)), abs(w1 - w2) / abs(w1))
