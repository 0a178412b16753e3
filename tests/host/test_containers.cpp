// Acceptance test of host/types.h beyond what the reference's test_types.cpp covers: the quirks callers can
// observe and the conversion helpers the solver seam uses.
#include <cassert>
#include <cmath>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "types.h"

int main()
{
    // construction, bounds-checked indexing, grow/shrink
    Vector v(std::vector<CalcType>{3, 1, 2});
    assert(v.size() == 3 && !v.empty() && v.first() == 3 && v.last() == 2);
    bool threw = false;
    try { v[3] = 1; } catch (const std::out_of_range&) { threw = true; }
    assert(threw);
    v.resize(5);
    assert(v[3] == 0 && v[4] == 0);
    v.resize(3);
    v.sort();
    assert(v[0] == 1 && v[2] == 3 && v.min_value() == 1 && v.max_value() == 3 && v.sum() == 6 && v.avg_value() == 2);
    assert(Vector(4).size() == 4 && Vector(4)[2] == 0 && Vector(2, 7.5)[1] == 7.5);

    // head / tail / erase_first / add / add_unique
    Vector h = v.head(2), t = v.tail(2);
    assert(h.size() == 2 && h[1] == 2 && t[0] == 2 && t[1] == 3);
    v.erase_first(1);
    assert(v.size() == 2 && v[0] == 2);
    v.add(9); v.add(h);
    assert(v.size() == 5 && v[2] == 9 && v[4] == 2);
    assert(!v.add_unique(9) && v.add_unique(10) && v.last() == 10);

    // setLinSpaced: min + i*step, last element is max verbatim
    Vector lin = Vector::setLinSpaced(5, 0, 4);
    for (int i = 0; i < 5; i++) assert(lin[i] == i);
    Vector sig = Vector::setLinSpaced(160, 1e-8, 0.2);
    assert(sig.size() == 160 && sig[0] == 1e-8 && sig.last() == 0.2 && sig[12] == 0.01509434886792453);

    // scalar and element-wise operators; unary minus is a multiplication by -1
    Vector a(std::vector<CalcType>{1, 2, 3}), b(std::vector<CalcType>{10, 20});
    Vector s = a + b;                       // zero-extends BOTH operands: b is a copy, a itself grows
    assert(s.size() == 3 && s[0] == 11 && s[1] == 22 && s[2] == 3);
    Vector shorter(std::vector<CalcType>{1}), longer(std::vector<CalcType>{1, 2, 3});
    Vector d = shorter - longer;
    assert(shorter.size() == 3 && d[2] == -3);
    assert((a * 2.0)[2] == 6 && (a / 2.0)[0] == 0.5 && (a + 1.0)[0] == 2 && (a - 1.0)[0] == 0);
    assert(std::signbit((-Vector(1))[0]));
    assert((a * a)[2] == 9 && (a / a)[1] == 1);
    assert(Vector::abs(-a)[1] == 2 && Vector::log(Vector(1, 1.0))[0] == 0);

    // range-for over a Vector mutates in place (the clamp loop of the solver)
    Vector neg(std::vector<CalcType>{-1, 2, -3});
    for (auto& e : neg) if (e < 0) e = 0;
    assert(neg[0] == 0 && neg[1] == 2 && neg[2] == 0);

    std::ostringstream out;
    out << a;
    assert(out.str() == "1, 2, 3\n");
    assert(a.MaxStep().first == 1 && a.MaxStep().second == 1);

    // Matrix
    Matrix m = VectorToMatrix(Vector(std::vector<CalcType>{1, 2, 3, 4}), 2);
    assert(m.size() == 2 && m[1][0] == 3 && m.max_value() == 4 && m.min_value() == 1);
    assert((m * 2.0)[1][1] == 8 && (m + 1.0)[0][0] == 2 && (-m)[0][1] == -2 && (m / 2.0)[0][0] == 0.5 && (m - 1.0)[1][1] == 3);
    assert(m.Row(1)[0] == 3 && m.Col(1)[0] == 2);
    assert(Matrix(3, 1.5)[2][2] == 1.5);
    assert(!IsSymmetric(m));
    assert(MatrixToVector(m)[2] == 3 && MatrixToVector(m, true).size() == 3);
    Matrix sym = VectorToMatrix(Vector(std::vector<CalcType>{1, 2, 3}), 2, true);
    assert(sym[1][0] == 2 && IsSymmetric(sym));

    // state layout of the solver seam: [N | C row-major]
    MatrixOfMatrix mm(1);
    mm(0, 0) = m;
    Vector flat = MatrixOfMatrixToVector(mm);
    assert(flat.size() == 4 && flat[1] == 2);
    MatrixOfMatrix back = VectorToMatrixOfMatrix(flat, 1, 2);
    assert(back(0, 0)[1][1] == 4);
    VectorOfMatrix vm(2);
    vm(1) = m;
    assert(vm.size() == 2 && vm(1)[0][1] == 2);

    // Function family (source compatibility only)
    Function f(Vector(std::vector<CalcType>{0, 1, 2}), Vector(std::vector<CalcType>{0, 10, 20}));
    assert(f.Interpolate(0.5) == 5 && f.Interpolate(5, -1) == -1 && f.findY(1) == 10 && f.ConstantSign().first);
    assert((f * 2.0).Y()[2] == 40 && Function::trim(f, 1).size() == 2 && f.IntegralRectangles() != 0);
    assert(MatrixOfFunction(2).size() == 2 && VectorOfFunction(3).size() == 3);

    // Time
    std::ostringstream clock_text;
    clock_text << Time((clock_t)(3725 * (long)CLOCKS_PER_SEC));
    assert(clock_text.str() == "1:02:05");
    assert(sgn(-2.0) == -1 && sgn(0) == 0 && sgn(3) == 1);

    std::cout << "containers ok\n";
    return 0;
}
