import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
ldm = load_package()
for name, (sw, sm, G) in {"14.4a": (0.15, 0.15, 0), "14.4b": (0.05, 0.05, 0), "14.4c": (0.015, 0.15, 0), "g64": (0.1, 0.1, 64)}.items():
    props = ldm.Properties.Generate(1, sw, sm, 0.4, 0.2, 0.001, G)
    eq = ldm.Equation(props)
    eq.upload(eq.GenerateEquilibriumValuesForNAndC(200))
    eq.advance(50, 0.1); eq.synchronize()
    eq.advance(1000, 0.1); eq.synchronize()
    print(name, "G", props.G, "us/step", eq.last_elapsed_ms())
