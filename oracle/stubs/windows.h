// Stub for building the reference's Solver.cpp on Linux (test infrastructure only).
// The reference calls WinExec("python Visualiser.py", 1) after writing its CSV
// (Solver.cpp:980); the harness never needs the plot, so this is a no-op.
#pragma once
static inline unsigned WinExec(const char*, unsigned) { return 0; }
