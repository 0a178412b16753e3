// Stub for building the reference's Solver.cpp on Linux (test infrastructure only).
// getch() (Solver.cpp:1241) replays key codes from the LDM_REF_KEYS environment
// variable ("80,13,27" = down, enter, esc); when exhausted it returns ESC (27).
#pragma once
#include <cstdlib>
#include <cstring>
static inline int getch()
{
    static const char* cursor = std::getenv("LDM_REF_KEYS");
    if (!cursor || !*cursor) return 27;
    char* end = nullptr;
    long code = std::strtol(cursor, &end, 10);
    if (end == cursor) { cursor = nullptr; return 27; }
    cursor = (*end == ',') ? end + 1 : end;
    return (int)code;
}
