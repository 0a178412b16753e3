#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Streams the reference's Solver.cpp to stdout with the FOUR lines of INTEGRATION.md A added
(include/ldm_reference_binding.h explains them): nothing of the reference is stored in this repository, the patched
text exists only for the duration of the compile (oracle/Makefile, target _ref/solver_bound).

    python oracle/patch_binding.py "/root/reference/Euler Solver ver.5.3/Solver.cpp" > $tmp/solver_bound.cpp
"""
import sys

INCLUDE = '#include "ldm_reference_binding.h"'
ACCESSOR = "    Properties& LdmProps() { return props; }"
RHS = "        return ldm_binding::Rhs(props, vals, (int)chosen_type);"
SOLVE = ("        return ldm_binding::Solve(solver_results, current_equation.LdmProps(), (int)type, time_max, step, "
         "initial_value, debug_print_enabled);")


def main():
    lines = open(sys.argv[1], newline="").read().split("\n")
    out, added = [], 0
    in_euler, k = False, 0
    while k < len(lines):
        line = lines[k]
        text = line.rstrip("\r")
        out.append(line)
        if text == '#include "types.h"':                                              # Solver.cpp:14
            out.append(INCLUDE); added += 1
        elif text == "    Equation(Properties p) : props(p) {}":                       # Solver.cpp:413
            out.append(ACCESSOR); added += 1
        elif text == "    Vector operator () (CalcType times, Vector vals, ExtrapolationType chosen_type)":   # Solver.cpp:431
            k += 1
            assert lines[k].rstrip("\r") == "    {"
            out.append(lines[k]); out.append(RHS); added += 1
        elif text == "class EulerSolver":                                              # Solver.cpp:541
            in_euler = True
        elif in_euler and text == "    ) {":                                           # Solver.cpp:564: end of Solve's parameter list
            out.append(SOLVE); added += 1
            in_euler = False
        k += 1
    assert added == 4, f"expected to add 4 lines, added {added}: the reference is not the file this recipe was written for"
    sys.stdout.write("\n".join(out))


if __name__ == "__main__":
    main()
