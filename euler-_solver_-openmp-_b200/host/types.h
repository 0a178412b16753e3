// types.h -- host-side containers of the B200 Law-Dieckmann solver front-end.
//
// API-compatible with the containers of the reference ("Euler Solver ver.5.3/types.h": Vector T:19-295,
// Matrix T:299-428, VectorOfMatrix T:432-444, MatrixOfMatrix T:448-465, Function family T:469-727,
// conversions T:731-840, Time T:844-868), so that code written against the reference -- including its
// test_types.cpp -- compiles and behaves the same.  Written from the interface description, not from the
// reference's text; the CUDA kernels never see these types (they take flat double arrays through ldm.h).
//
// Behaviours kept on purpose because callers can observe them:
//   * everything has value semantics; operator[] is bounds-checked (std::out_of_range);
//   * Vector (op) Vector zero-extends BOTH operands to the longer length, which resizes *this (T:208-265);
//   * unary minus multiplies by -1, so 0 becomes -0 (T:175-183);
//   * setLinSpaced takes the count as CalcType and appends max verbatim as the last element (T:120-130);
//   * size() returns int.
// One deliberate difference: Matrix::min_value returns the true minimum (the reference returns the largest
// row minimum, T:343-350, which nothing uses).
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <ctime>
#include <functional>
#include <numeric>
#include <ostream>
#include <string>
#include <utility>
#include <vector>

using namespace std;   // the reference header exports this too; code written against it relies on it

template <typename T>
int sgn(T val) { return (T(0) < val) - (val < T(0)); }

using CalcType = double;

//--------------------------------------------------------------------------------------------------- Vector
class Vector
{
    std::vector<CalcType> data_;

    template <typename Op>
    Vector mapped(Op op) const
    {
        Vector out;
        out.data_.reserve(data_.size());
        for (CalcType v : data_) out.data_.push_back(op(v));
        return out;
    }

    // element-wise with zero extension of both sides to the longer length (the left side is *this)
    template <typename Op>
    Vector zipped(Vector& rhs, Op op)
    {
        const size_t n = std::max(data_.size(), rhs.data_.size());
        data_.resize(n);
        rhs.data_.resize(n);
        Vector out;
        out.data_.reserve(n);
        for (size_t i = 0; i < n; i++) out.data_.push_back(op(data_[i], rhs.data_[i]));
        return out;
    }

public:
    Vector(std::vector<CalcType> values) : data_(std::move(values)) {}
    Vector(const int n = 0) : data_((size_t)std::max(n, 0), CalcType(0)) {}
    Vector(const int n, CalcType value) : data_((size_t)std::max(n, 0), value) {}

    std::vector<CalcType> get_vector() { return data_; }
    CalcType& operator[](const int index) { return data_.at((size_t)index); }
    const CalcType& operator[](const int index) const { return data_.at((size_t)index); }

    bool empty() const { return data_.empty(); }
    int size() const { return (int)data_.size(); }
    void resize(int n) { data_.resize((size_t)n); }
    void clear() { data_.clear(); }
    void sort() { std::sort(data_.begin(), data_.end()); }

    void add(CalcType value) { data_.push_back(value); }
    void add(Vector other) { data_.insert(data_.end(), other.data_.begin(), other.data_.end()); }
    bool add_unique(CalcType value)
    {
        if (std::find(data_.begin(), data_.end(), value) != data_.end()) return false;
        data_.push_back(value);
        return true;
    }

    CalcType first() const { return data_.at(0); }
    CalcType last() const { return data_.at(data_.size() - 1); }
    Vector head(int n) const
    {
        Vector out;
        for (int i = 0; i < n; i++) out.add(data_.at((size_t)i));
        return out;
    }
    Vector tail(int n) const
    {
        Vector out;
        for (int i = n; i > 0; i--) out.add(data_.at(data_.size() - (size_t)i));
        return out;
    }
    void erase_first(int n) { data_.erase(data_.begin(), data_.begin() + std::min<size_t>((size_t)std::max(n, 0), data_.size())); }

    CalcType min_value() const { return *std::min_element(data_.begin(), data_.end()); }
    CalcType max_value() const { return *std::max_element(data_.begin(), data_.end()); }
    CalcType sum() const { return std::accumulate(data_.begin(), data_.end(), CalcType(0)); }
    CalcType avg_value() const { return data_.empty() ? CalcType(0) : sum() / (CalcType)data_.size(); }

    static Vector setLinSpaced(CalcType num, CalcType min_val, CalcType max_val)
    {
        Vector out;
        const CalcType step = (max_val - min_val) / (num - 1);
        for (int i = 0; i < num - 1; i++) out.add(min_val + i * step);
        out.add(max_val);
        return out;
    }
    static Vector log(Vector arg) { return arg.mapped([](CalcType v) { return std::log(v); }); }
    static Vector abs(Vector arg) { return arg.mapped([](CalcType v) { return std::abs(v); }); }

    Vector operator+(CalcType s) const { return mapped([s](CalcType v) { return v + s; }); }
    Vector operator-(CalcType s) const { return mapped([s](CalcType v) { return v - s; }); }
    Vector operator*(CalcType s) const { return mapped([s](CalcType v) { return v * s; }); }
    Vector operator/(CalcType s) const { return mapped([s](CalcType v) { return v / s; }); }
    Vector operator-() const { return mapped([](CalcType v) { return -1 * v; }); }

    Vector operator+(Vector rhs) { return zipped(rhs, [](CalcType a, CalcType b) { return a + b; }); }
    Vector operator-(Vector rhs) { return zipped(rhs, [](CalcType a, CalcType b) { return a - b; }); }
    Vector operator*(Vector rhs) { return zipped(rhs, [](CalcType a, CalcType b) { return a * b; }); }
    Vector operator/(Vector rhs) { return zipped(rhs, [](CalcType a, CalcType b) { return a / b; }); }

    std::vector<CalcType>::iterator begin() { return data_.begin(); }
    std::vector<CalcType>::iterator end() { return data_.end(); }
    std::vector<CalcType>::const_iterator begin() const { return data_.begin(); }
    std::vector<CalcType>::const_iterator end() const { return data_.end(); }
    const CalcType* raw() const { return data_.data(); }   // flat view for the C ABI
    CalcType* raw() { return data_.data(); }

    friend std::ostream& operator<<(std::ostream& out, const Vector& v)
    {
        out << v.data_.at(0);
        for (size_t i = 1; i < v.data_.size(); i++) out << ", " << v.data_[i];
        return out << '\n';
    }

    // index and size of the largest gap between neighbours
    std::pair<int, CalcType> MaxStep() const
    {
        CalcType widest = data_.at(1) - data_.at(0);
        int where = 1;
        for (size_t i = 2; i < data_.size(); i++)
            if (data_[i] - data_[i - 1] > widest) { widest = data_[i] - data_[i - 1]; where = (int)i; }
        return std::make_pair(where, widest);
    }
};

//--------------------------------------------------------------------------------------------------- Matrix
class Matrix
{
    std::vector<Vector> rows_;

    template <typename Op>
    Matrix mapped(Op op)
    {
        Matrix out;
        out.resize((int)rows_.size(), rows_.at(0).size());
        for (int i = 0; i < out.size(); i++)
            for (int j = 0; j < out.size(); j++) out[i][j] = op(rows_.at((size_t)i)[j]);   // square, like the reference
        return out;
    }

public:
    Matrix(const int size = 0) { resize(size); }
    Matrix(const int size, CalcType value)
    {
        resize(size);
        for (Vector& r : rows_)
            for (CalcType& e : r) e = value;
    }

    Vector& operator[](const int row) { return rows_.at((size_t)row); }
    const Vector& operator[](const int row) const { return rows_.at((size_t)row); }
    int size() const { return (int)rows_.size(); }

    void resize(const int size) { resize(size, size); }
    void resize(const int r, const int c)
    {
        rows_.resize((size_t)r);
        for (Vector& row : rows_) row.resize(c);
    }

    Vector Row(const int row) const { return rows_.at((size_t)row); }
    Vector Col(const int col) const
    {
        Vector out;
        for (const Vector& r : rows_) out.add(r[col]);
        return out;
    }

    CalcType min_value() const
    {
        CalcType best = rows_.at(0).min_value();
        for (const Vector& r : rows_) best = std::min(best, r.min_value());
        return best;
    }
    CalcType max_value() const
    {
        CalcType best = rows_.at(0).max_value();
        for (const Vector& r : rows_) best = std::max(best, r.max_value());
        return best;
    }

    Matrix operator+(CalcType s) { return mapped([s](CalcType v) { return v + s; }); }
    Matrix operator-(CalcType s) { return mapped([s](CalcType v) { return v - s; }); }
    Matrix operator*(CalcType s) { return mapped([s](CalcType v) { return v * s; }); }
    Matrix operator/(CalcType s) { return mapped([s](CalcType v) { return v / s; }); }
    Matrix operator-() { return mapped([](CalcType v) { return -1 * v; }); }

    friend std::ostream& operator<<(std::ostream& out, const Matrix& m)
    {
        for (const Vector& r : m.rows_) out << r;
        return out;
    }
};

//------------------------------------------------------------------------- containers of matrices
class VectorOfMatrix
{
    std::vector<Matrix> items_;

public:
    VectorOfMatrix(const int n = 0) : items_((size_t)n) {}
    int size() const { return (int)items_.size(); }
    Matrix& operator()(const int index) { return items_.at((size_t)index); }
};

class MatrixOfMatrix
{
    std::vector<std::vector<Matrix>> items_;

public:
    MatrixOfMatrix(const int size = 0) : items_((size_t)size, std::vector<Matrix>((size_t)size)) {}
    int size() const { return (int)items_.size(); }
    Matrix& operator()(const int row, const int col) { return items_.at((size_t)row).at((size_t)col); }
};

//------------------------------------------------------------------------------------------------- Function
// Samples (x, y) of a scalar function.  Not used by the solver or by test_types.cpp; kept so that code
// written against the reference header still compiles.
class Function
{
    Vector x;
    Vector y;

    void from_symmetric_table(Vector grid, Matrix values, bool limited, CalcType max_x)
    {
        std::vector<std::pair<CalcType, CalcType>> samples;
        for (int i = 0; i < grid.size(); i++)
            for (int j = i; j < grid.size(); j++) {
                assert(std::abs(values[i][j] - values[j][i]) < 1e-08);
                samples.emplace_back(std::hypot(grid[i], grid[j]), values[i][j]);
            }
        std::stable_sort(samples.begin(), samples.end(),
                         [](const std::pair<CalcType, CalcType>& a, const std::pair<CalcType, CalcType>& b) { return a.first < b.first; });
        for (const auto& s : samples) {
            if (limited && !(s.first <= max_x)) break;
            x.add(s.first);
            y.add(s.second);
        }
    }

    Function with_y(std::function<CalcType(CalcType)> op)
    {
        Vector ny;
        for (int i = 0; i < y.size(); i++) ny.add(op(y[i]));
        return Function(x, ny);
    }

    CalcType lerp_at(CalcType at)
    {
        int i = 1;
        while (at > x[i]) i++;
        const CalcType w = (at - x[i - 1]) / (x[i] - x[i - 1]);
        return y[i - 1] + w * (y[i] - y[i - 1]);
    }

public:
    Function(const int size = 0) : x(size), y(size) {}
    Function(Vector x_grid, Vector y_values) : x(x_grid), y(y_values) {}
    Function(Vector x_grid, Matrix y_values) { from_symmetric_table(x_grid, y_values, false, 0); }       // radial profile
    Function(Vector x_grid, Matrix y_values, CalcType max_x_value) { from_symmetric_table(x_grid, y_values, true, max_x_value); }

    int size() { return x.size(); }
    Vector X() { return x; }
    Vector Y() { return y; }

    static Function trim(Function func, CalcType final_x_value)
    {
        Vector xs = func.X(), ys = func.Y(), tx, ty;
        int last = -1;
        for (int i = 0; i < xs.size() && last < 0; i++)
            if (xs[i] == final_x_value) last = i;
        assert(last >= 0);
        for (int i = 0; i <= last; i++) { tx.add(xs[i]); ty.add(ys[i]); }
        return Function(tx, ty);
    }

    CalcType findY(CalcType x_value)
    {
        assert((x_value >= x[0]) && (x_value <= x.last()));
        for (int i = 0; i < x.size(); i++)
            if (x[i] == x_value) return y[i];
        assert(false && "findY: no such abscissa");
        return 0;
    }

    std::pair<bool, CalcType> ConstantSign()
    {
        bool constant = true;
        const int first_sign = sgn(y[1] - y[0]);
        CalcType where = 0;
        for (int i = 2; i < y.size(); i++) {
            const int s = sgn(y[i] - y[i - 1]);
            if (s != first_sign && s != 0) { constant = false; where = x[i]; }
        }
        return std::make_pair(constant, where);
    }

    CalcType Interpolate(CalcType x_value) { return lerp_at(x_value); }
    CalcType Interpolate(CalcType x_value, CalcType asymptotic_value) { return x_value > x.last() ? asymptotic_value : lerp_at(x_value); }

    CalcType IntegralRectangles()
    {
        assert(x.size() % 2 == 1);
        CalcType integral = 0;
        const CalcType step = (2 * (x.last() - x[0])) / x.size();
        for (int i = 1 - x.size(); i < x.size() - 1; i += 2) integral += y[std::abs(i + 1)];
        assert(integral != 0);
        return integral * step;
    }

    Function operator+(CalcType s) { return with_y([s](CalcType v) { return v + s; }); }
    Function operator-(CalcType s) { return with_y([s](CalcType v) { return v - s; }); }
    Function operator*(CalcType s) { return with_y([s](CalcType v) { return v * s; }); }
    Function operator/(CalcType s) { return with_y([s](CalcType v) { return v / s; }); }
    Function operator-() { return with_y([](CalcType v) { return -1 * v; }); }

    friend std::ostream& operator<<(std::ostream& out, Function& f)
    {
        for (int i = 0; i < f.size(); i++) out << "X: " << f.x[i] << "   Y: " << f.y[i] << '\n';
        return out;
    }
};

class VectorOfFunction
{
    std::vector<Function> items_;

public:
    VectorOfFunction(int size = 0) : items_((size_t)size) {}
    int size() const { return (int)items_.size(); }
    Function& operator()(const int index) { return items_.at((size_t)index); }
};

class MatrixOfFunction
{
    std::vector<std::vector<Function>> items_;

public:
    MatrixOfFunction(int size = 0) : items_((size_t)size, std::vector<Function>((size_t)size)) {}
    int size() const { return (int)items_.size(); }
    Function& operator()(int row, int col) { return items_.at((size_t)row).at((size_t)col); }
};

//------------------------------------------------------------------------------------------- conversions
// Row-major flattening; `symmetric` packs / unpacks the upper triangle only.
inline int packed_size(int mat_size, bool symmetric) { return symmetric ? mat_size * (mat_size + 1) / 2 : mat_size * mat_size; }

inline Vector MatrixToVector(Matrix mat, bool symmetric = false)
{
    Vector out;
    for (int i = 0; i < mat.size(); i++)
        for (int j = symmetric ? i : 0; j < mat.size(); j++) out.add(mat[i][j]);
    return out;
}

inline Vector MatrixOfMatrixToVector(MatrixOfMatrix mm, bool symmetric = false)
{
    Vector out;
    for (int i = 0; i < mm.size(); i++)
        for (int j = 0; j < mm.size(); j++) out.add(MatrixToVector(mm(i, j), symmetric));
    return out;
}

inline Matrix VectorToMatrix(Vector vec, int mat_size, bool symmetric = false)
{
    assert(vec.size() == packed_size(mat_size, symmetric));
    Matrix out(mat_size);
    int next = 0;
    for (int i = 0; i < mat_size; i++)
        for (int j = symmetric ? i : 0; j < mat_size; j++) {
            out[i][j] = vec[next];
            if (symmetric) out[j][i] = vec[next];
            next++;
        }
    return out;
}

inline MatrixOfMatrix VectorToMatrixOfMatrix(Vector vec, int mat_of_mat_size, int mat_size, bool symmetric = false)
{
    const int block = packed_size(mat_size, symmetric);
    assert(vec.size() == block * mat_of_mat_size * mat_of_mat_size);
    MatrixOfMatrix out(mat_of_mat_size);
    int next = 0;
    for (int i = 0; i < mat_of_mat_size; i++)
        for (int j = 0; j < mat_of_mat_size; j++) {
            Vector piece;
            for (int k = 0; k < block; k++) piece.add(vec[next++]);
            out(i, j) = VectorToMatrix(piece, mat_size, symmetric);
        }
    return out;
}

inline MatrixOfFunction VectorToMatrixOfFunction(Vector x_vec, Vector y_vec, int mat_of_func_size)
{
    assert(y_vec.size() == mat_of_func_size * mat_of_func_size * x_vec.size());
    MatrixOfFunction out(mat_of_func_size);
    int next = 0;
    for (int i = 0; i < mat_of_func_size; i++)
        for (int j = 0; j < mat_of_func_size; j++) {
            Vector ys;
            for (int k = 0; k < x_vec.size(); k++) ys.add(y_vec[next++]);
            out(i, j) = Function(x_vec, ys);
        }
    return out;
}

inline bool IsSymmetric(Matrix mat)
{
    for (int i = 0; i < mat.size(); i++)
        for (int j = i + 1; j < mat.size(); j++)
            if (mat[i][j] != mat[j][i]) return false;
    return true;
}

//------------------------------------------------------------------------------------------------------ Time
// h:mm:ss from a clock_t difference.  The reference truncates the clock_t to int BEFORE dividing (T:853),
// which overflows after ~2147 s of CPU time; this one divides first.
class Time
{
    long long total_seconds_;

public:
    Time(clock_t ticks) : total_seconds_((long long)(ticks / CLOCKS_PER_SEC)) {}
    explicit Time(double seconds) : total_seconds_((long long)seconds) {}

    friend std::ostream& operator<<(std::ostream& out, const Time& t)
    {
        const long long s = t.total_seconds_ < 0 ? -t.total_seconds_ : t.total_seconds_;
        const long long h = s / 3600, m = (s % 3600) / 60, sec = s % 60;
        return out << h << ':' << (m < 10 ? "0" : "") << m << ':' << (sec < 10 ? "0" : "") << sec;
    }
};
