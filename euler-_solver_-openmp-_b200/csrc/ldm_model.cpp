// Host-side model setup (see ldm_model.h).  Compiled with -ffp-contract=off: the grids and kernel
// tables must equal the reference's bit for bit, because interval indices and the inside/outside
// decision at the cutoff circle are derived from them by exact comparisons.
//
// S:n = "/root/reference/Euler Solver ver.5.3/Solver.cpp" line n, T:n = types.h line n.
#include "ldm_model.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

namespace ldm {

namespace {

const double kPi = 3.1415927;     // S:23 -- the reference's own 8-digit constant, part of the normalisation
const double kEpsilon = 1E-08;    // S:24

// Vector::setLinSpaced (T:120-130): min + i*step for i < num-1, then max verbatim.
std::vector<double> lin_spaced(int num, double lo, double hi)
{
    std::vector<double> v;
    v.reserve((size_t)num);
    const double step = (hi - lo) / ((double)num - 1);
    for (int i = 0; i < num - 1; i++) v.push_back(lo + i * step);
    v.push_back(hi);
    return v;
}

// The reference's node formula (S:69), one Vector operator at a time:
//   abs(-log(-setLinSpaced(n,0,n-1) / n + 1) * 0.01 * n)
std::vector<double> stretched_nodes(int n)
{
    std::vector<double> v = lin_spaced(n, 0, n - 1);
    for (double& e : v) e = -1 * e;          // unary minus is "x * -1" (T:175-183)
    for (double& e : v) e = e / n;
    for (double& e : v) e = e + 1;
    for (double& e : v) e = std::log(e);
    for (double& e : v) e = -1 * e;
    for (double& e : v) e = e * 0.01;
    for (double& e : v) e = e * n;
    for (double& e : v) e = std::abs(e);
    return v;
}

// Properties::GenerateKernel (S:98-145) for the Normal kernel type.
int fill_kernel(const std::vector<double>& grid, double step, double sigma, double radius, std::vector<double>& table)
{
    const int K = (int)grid.size();
    table.assign((size_t)K * K, 0.0);
    int extent = 0;
    for (int x = 0; x < K; x++)
        for (int y = 0; y < K; y++) {
            const double r = std::hypot(grid[x], grid[y]);
            double value = 0;
            if (r <= radius) {
                const double norm = sigma * std::sqrt(2 * kPi) * (1 - std::exp((-radius * radius) / (2 * sigma * sigma)));
                const double weight = std::exp((-r * r) / (2 * sigma * sigma));
                value = weight / norm;
                if (value != 0) extent = std::max({extent, x + 1, y + 1});
            }
            table[(size_t)x * K + y] = value;
        }
    if (K % 2 == 1 && extent % 2 == 0 && extent != 2) extent++;       // S:131
    if (extent == 1) extent = 2;                                       // S:132

    double integral = 0;                                               // S:135-138
    for (int x = 1 - extent; x < extent - 1; x += 2)
        for (int y = 1 - extent; y < extent - 1; y += 2)
            integral += table[(size_t)std::abs(x + 1) * K + std::abs(y + 1)];
    if (integral == 0) throw std::invalid_argument("kernel integral is zero (reference assert, Solver.cpp:140)");
    integral *= step * step;
    for (double& e : table) e = e / integral;                          // S:142
    return extent;
}

double midpoint_integral(const double* table, int K, int extent, double step)
{
    double integral = 0;                                               // S:156-161
    for (int x = 1 - extent; x < extent - 1; x += 2)
        for (int y = 1 - extent; y < extent - 1; y += 2)
            integral += table[(size_t)std::abs(x + 1) * K + std::abs(y + 1)];
    integral *= step * step;                                           // pow(step, 2)
    return integral;
}

// index of the first node i >= 1 with x <= grid[i]  (the scan of S:334-336)
inline int upper_node(const double* grid, int n, double x)
{
    return (int)(std::lower_bound(grid + 1, grid + n, x) - grid);
}

OffsetSet make_offsets(const ldm_problem& p, const double* table, int extent)
{
    OffsetSet o;
    const int K = p.K;
    for (int i = 1 - extent; i < extent - 1; i += 2) {                 // S:300
        o.offset.push_back(i >= 0 ? p.kernels_grid[i + 1] : -p.kernels_grid[-i - 1]);   // S:307
        o.node.push_back(std::abs(i + 1));                             // S:311
    }
    o.n_off = (int)o.offset.size();
    std::vector<int> nodes = o.node;
    std::sort(nodes.begin(), nodes.end());
    nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
    o.A = (int)nodes.size();
    for (int n : o.node) o.alpha.push_back((int)(std::lower_bound(nodes.begin(), nodes.end(), n) - nodes.begin()));
    o.quadrant.resize((size_t)o.A * o.A);
    for (int a = 0; a < o.A; a++)
        for (int c = 0; c < o.A; c++) o.quadrant[(size_t)a * o.A + c] = table[(size_t)nodes[a] * K + nodes[c]];
    return o;
}

}  // namespace

double bilinear(const double* grid, int n, const double* values, double x, double y, double max_distance,
                bool squared_n, double asymptotic_value)
{
    const double distance = std::hypot(x, y);                          // S:329
    if (distance > max_distance) {                                     // Extrapolation, S:355-373
        if (std::abs(values[n - 1] - asymptotic_value) < kEpsilon) return asymptotic_value;
        return squared_n ? asymptotic_value : 0.0;
    }
    const int i = upper_node(grid, n, x), j = upper_node(grid, n, y);
    const double wx = (x - grid[i - 1]) / (grid[i] - grid[i - 1]);
    const double wy = (y - grid[j - 1]) / (grid[j] - grid[j - 1]);
    return values[(size_t)(i - 1) * n + (j - 1)] * (1 - wx) * (1 - wy) + values[(size_t)i * n + (j - 1)] * wx * (1 - wy) +
           values[(size_t)(i - 1) * n + j] * (1 - wx) * wy + values[(size_t)i * n + j] * wx * wy;
}

void generate_problem(double sigma_w, double sigma_m, double b, double d, double d_prime, int forced_G, OwnedProblem& out)
{
    if (!(sigma_w > 0) || !(sigma_m > 0)) throw std::invalid_argument("sigma_w and sigma_m must be positive");
    if (forced_G > 0 && forced_G < 4) throw std::invalid_argument("forced_G must be at least 4");

    // GenerateGrids (S:63-95)
    const double cutoff_distance = (sigma_w + sigma_m) * 5;
    const double max_radius = std::max(sigma_w, sigma_m) * 5;
    int nodes = forced_G > 0 ? forced_G : 5;
    std::vector<double> grid = stretched_nodes(nodes);
    if (forced_G <= 0)
        while (grid.back() < cutoff_distance - 0.01) grid = stretched_nodes(++nodes);   // S:70-74
    const double scale = cutoff_distance / grid.back();
    for (double& e : grid) e = e * scale;
    if (!(std::abs(grid.back() - cutoff_distance) < kEpsilon))
        throw std::invalid_argument("moment grid does not end at the cutoff (reference assert, Solver.cpp:77)");

    int K = (nodes - 1) * 2 - 1;
    if (std::min(sigma_w, sigma_m) < 0.015) {                          // S:80-84
        K = (K - 2) * 4 - 1;
        K = (K - 1) * 2 - 1;
    }
    if (!((K % 2 == 1) && (K >= 5))) throw std::invalid_argument("kernel grid size must be odd and >= 5 (Solver.cpp:88)");

    out.moments_grid = grid;
    out.kernels_grid = lin_spaced(K, 0, max_radius);
    const double h = out.kernels_grid[2] - out.kernels_grid[0];        // S:94

    // Generate (S:205-239): radius 3 sigma for the Normal kernel
    const int mW = fill_kernel(out.kernels_grid, h, sigma_w, sigma_w * 3, out.W);
    const int mM = fill_kernel(out.kernels_grid, h, sigma_m, sigma_m * 3, out.M);

    ldm_problem& v = out.view;
    v.species = 1;
    v.G = nodes; v.K = K; v.mM = mM; v.mW = mW;
    v.h = h; v.cutoff = grid.back();
    v.sigma_m = sigma_m; v.sigma_w = sigma_w; v.b = b; v.d = d; v.d_prime = d_prime;
    v.moments_grid = out.moments_grid.data();
    v.kernels_grid = out.kernels_grid.data();
    v.M = out.M.data();
    v.W = out.W.data();
}

void generate_problem_multi(int species, const double* sigma_w, const double* sigma_m, const double* b, const double* d,
                            const double* d_prime, int forced_G, OwnedProblem& out)
{
    const int s = species;
    if (s < 1 || s > LDM_MAX_SPECIES) throw std::invalid_argument("species count out of range");
    double max_w = 0, max_m = 0, min_sigma = 1e300;
    for (int i = 0; i < s * s; i++) { if (!(sigma_w[i] > 0)) throw std::invalid_argument("sigma_w must be positive"); max_w = std::max(max_w, sigma_w[i]); min_sigma = std::min(min_sigma, sigma_w[i]); }
    for (int i = 0; i < s; i++) { if (!(sigma_m[i] > 0)) throw std::invalid_argument("sigma_m must be positive"); max_m = std::max(max_m, sigma_m[i]); min_sigma = std::min(min_sigma, sigma_m[i]); }
    if (forced_G > 0 && forced_G < 4) throw std::invalid_argument("forced_G must be at least 4");

    // GenerateGrids(sW.max_value(), sM.max_value()) (S:203, S:63-95): min(sigma_w, sigma_m) of S:80 sees only those two
    const double cutoff_distance = (max_w + max_m) * 5;
    const double max_radius = std::max(max_w, max_m) * 5;
    int nodes = forced_G > 0 ? forced_G : 5;
    std::vector<double> grid = stretched_nodes(nodes);
    if (forced_G <= 0)
        while (grid.back() < cutoff_distance - 0.01) grid = stretched_nodes(++nodes);
    const double scale = cutoff_distance / grid.back();
    for (double& e : grid) e = e * scale;
    if (!(std::abs(grid.back() - cutoff_distance) < kEpsilon))
        throw std::invalid_argument("moment grid does not end at the cutoff (reference assert, Solver.cpp:77)");
    int K = (nodes - 1) * 2 - 1;
    if (std::min(max_w, max_m) < 0.015) {
        K = (K - 2) * 4 - 1;
        K = (K - 1) * 2 - 1;
    }
    if (!((K % 2 == 1) && (K >= 5))) throw std::invalid_argument("kernel grid size must be odd and >= 5 (Solver.cpp:88)");
    (void)min_sigma;
    out.moments_grid = grid;
    out.kernels_grid = lin_spaced(K, 0, max_radius);
    const double h = out.kernels_grid[2] - out.kernels_grid[0];

    out.W.clear(); out.M.clear(); out.mW_of.clear(); out.mM_of.clear();
    std::vector<double> table;
    for (int i = 0; i < s; i++)
        for (int j = 0; j < s; j++) {                                   // S:212-220
            out.mW_of.push_back(fill_kernel(out.kernels_grid, h, sigma_w[i * s + j], sigma_w[i * s + j] * 3, table));
            out.W.insert(out.W.end(), table.begin(), table.end());
        }
    for (int i = 0; i < s; i++) {                                       // S:229-234
        out.mM_of.push_back(fill_kernel(out.kernels_grid, h, sigma_m[i], sigma_m[i] * 3, table));
        out.M.insert(out.M.end(), table.begin(), table.end());
    }
    out.sigma_w_of.assign(sigma_w, sigma_w + s * s);
    out.sigma_m_of.assign(sigma_m, sigma_m + s);
    out.b_of.assign(b, b + s);
    out.d_of.assign(d, d + s);
    out.d_prime_of.assign(d_prime, d_prime + s * s);

    ldm_problem& v = out.view;
    v.species = s;
    v.G = nodes; v.K = K; v.mM = out.mM_of[0]; v.mW = out.mW_of[0];
    v.h = h; v.cutoff = grid.back();
    v.sigma_m = sigma_m[0]; v.sigma_w = sigma_w[0]; v.b = b[0]; v.d = d[0]; v.d_prime = d_prime[0];
    v.moments_grid = out.moments_grid.data();
    v.kernels_grid = out.kernels_grid.data();
    v.M = out.M.data();
    v.W = out.W.data();
    v.mM_of = out.mM_of.data(); v.mW_of = out.mW_of.data();
    v.sigma_m_of = out.sigma_m_of.data(); v.sigma_w_of = out.sigma_w_of.data();
    v.b_of = out.b_of.data(); v.d_of = out.d_of.data(); v.d_prime_of = out.d_prime_of.data();
}

ldm_problem species_view(const ldm_problem& p, int i, int j)
{
    ldm_problem v = p;
    const int s = p.species;
    const size_t KK = (size_t)p.K * p.K;
    v.species = 1;
    v.M = p.M + (size_t)i * KK;
    v.W = p.W + (size_t)(i * s + j) * KK;
    v.mM = p.mM_of[i]; v.mW = p.mW_of[i * s + j];
    v.sigma_m = p.sigma_m_of[i]; v.sigma_w = p.sigma_w_of[i * s + j];
    v.b = p.b_of[i]; v.d = p.d_of[i]; v.d_prime = p.d_prime_of[i * s + j];
    v.mM_of = nullptr; v.mW_of = nullptr; v.sigma_m_of = nullptr; v.sigma_w_of = nullptr; v.b_of = nullptr; v.d_of = nullptr; v.d_prime_of = nullptr;
    return v;
}

void check_kernels(const ldm_problem& p, double& integral_w, double& integral_m)
{
    integral_w = midpoint_integral(p.W, p.K, p.mW, p.h);
    integral_m = midpoint_integral(p.M, p.K, p.mM, p.h);
}

void build_tables(const ldm_problem& p, Tables& t)
{
    const int G = p.G, K = p.K;
    const double* g = p.moments_grid;
    t.G = G; t.K = K; t.h = p.h; t.hh = p.h * p.h; t.cutoff = p.cutoff;
    t.b = p.b; t.d = p.d; t.dd = p.d_prime;

    // ---- m-kernel quadrature on every moment node (the per-axis halves of S:307-308, S:334-345) ----
    t.offM = make_offsets(p, p.M, p.mM);
    const int n_off = t.offM.n_off;
    AxisTable& ax = t.axisM;
    const size_t cells = (size_t)G * n_off;
    ax.s.resize(cells); ax.upper.assign(cells, 0); ax.weight.assign(cells, 0.0); ax.limit.assign(cells, -1.0);
    for (int x = 0; x < G; x++)
        for (int o = 0; o < n_off; o++) {
            const double s = std::abs(g[x] + t.offM.offset[o]);
            const size_t c = (size_t)x * n_off + o;
            ax.s[c] = s;
            if (s <= p.cutoff) {           // otherwise hypot(s, anything) > cutoff: always outside
                const int i = upper_node(g, G, s);
                ax.upper[c] = i;
                ax.weight[c] = (s - g[i - 1]) / (g[i] - g[i - 1]);
            }
        }

    // ---- the cutoff circle (S:329-330), decided once on the host with the reference's hypot ----
    // Candidates for a sample's partner abscissa are exactly the values in ax.s, so the largest
    // candidate that is still inside is an exact threshold for the device (no hypot on the GPU).
    std::vector<double> sorted(ax.s);
    std::sort(sorted.begin(), sorted.end());
    sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    auto inside = [&](double a, double c) { return !(std::hypot(a, c) > p.cutoff); };
    double closest = 1e300;
    long long non_monotone = 0;
    for (size_t c = 0; c < cells; c++) {
        const double s = ax.s[c];
        if (!(s <= p.cutoff) || !inside(s, sorted.front())) { ax.limit[c] = -1.0; continue; }
        size_t lo = 0, hi = sorted.size();           // invariant: sorted[lo] inside, sorted[hi] outside (or end)
        while (hi - lo > 1) {
            const size_t mid = lo + (hi - lo) / 2;
            if (inside(s, sorted[mid])) lo = mid; else hi = mid;
        }
        ax.limit[c] = sorted[lo];
        for (size_t q = (lo >= 3 ? lo - 3 : 0); q < std::min(sorted.size(), lo + 4); q++) {
            if (inside(s, sorted[q]) != (q <= lo)) non_monotone++;
            closest = std::min(closest, std::abs(std::hypot(s, sorted[q]) - p.cutoff) / p.cutoff);
        }
    }
    if (non_monotone) throw std::runtime_error("hypot is not monotone across the sample set near the cutoff");
    t.s_max = 0;
    for (double v : sorted) if (v <= p.cutoff) t.s_max = v;
    t.far_inside.assign((size_t)G, -1.0);
    for (int x = 0; x < G; x++)
        for (int o = 0; o < n_off; o++)
            if (ax.upper[(size_t)x * n_off + o] != 0)
                t.far_inside[x] = std::max(t.far_inside[x], ax.s[(size_t)x * n_off + o]);
    char note[160];
    std::snprintf(note, sizeof note, "samples=%zu distinct=%zu closest |hypot-cutoff|/cutoff=%.3e", cells, sorted.size(), closest);
    t.census = note;

    // ---- kernels at the moment nodes (S:469-476, S:490-497): time-invariant ----
    t.m_nodes.resize((size_t)G * G);
    t.w_nodes.resize((size_t)G * G);
    for (int x = 0; x < G; x++)
        for (int y = 0; y < G; y++) {
            t.m_nodes[(size_t)x * G + y] = bilinear(p.kernels_grid, K, p.M, g[x], g[y], p.sigma_m * 3, false, 0.0);
            t.w_nodes[(size_t)x * G + y] = bilinear(p.kernels_grid, K, p.W, g[x], g[y], p.sigma_w * 3, false, 0.0);
        }

    // ---- WC = integral of w C at shift (0,0) (S:265-274, S:297-316) as a linear functional of C ----
    // Inside samples spread w * (bilinear weights) onto four nodes; outside samples contribute w * N^2.
    const OffsetSet offW = make_offsets(p, p.W, p.mW);
    t.wc_operator.assign((size_t)G * G, 0.0);
    t.wc_outside = 0;
    t.wc_out_w.clear(); t.wc_out_distance.clear();
    for (int oi = 0; oi < offW.n_off; oi++) {
        const double sx = std::abs(g[0] + offW.offset[oi]);
        for (int oj = 0; oj < offW.n_off; oj++) {
            const double sy = std::abs(g[0] + offW.offset[oj]);
            const double w = offW.quadrant[(size_t)offW.alpha[oi] * offW.A + offW.alpha[oj]];
            if (std::hypot(sx, sy) > p.cutoff) {
                t.wc_outside += w;
                if (w != 0.0) { t.wc_out_w.push_back(w); t.wc_out_distance.push_back(std::hypot(sx, sy)); }
                continue;
            }
            const int i = upper_node(g, G, sx), j = upper_node(g, G, sy);
            const double wx = (sx - g[i - 1]) / (g[i] - g[i - 1]);
            const double wy = (sy - g[j - 1]) / (g[j] - g[j - 1]);
            t.wc_operator[(size_t)(i - 1) * G + (j - 1)] += w * ((1 - wx) * (1 - wy));
            t.wc_operator[(size_t)i * G + (j - 1)] += w * (wx * (1 - wy));
            t.wc_operator[(size_t)(i - 1) * G + j] += w * ((1 - wx) * wy);
            t.wc_operator[(size_t)i * G + j] += w * (wx * wy);
        }
    }
}

void detect_separable(const OffsetSet& off, double tolerance, SeparableKernel& out)
{
    const int A = off.A;
    const double* Q = off.quadrant.data();
    out = SeparableKernel();
    out.a_star.assign((size_t)A, -1);
    for (int b = 0; b < A; b++)
        for (int a = 0; a < A; a++)
            if (Q[(size_t)b * A + a] != 0.0) out.a_star[b] = a;
    // staircase: every row is non-zero exactly on [0, a_star], a_star never grows with beta, the table is symmetric
    for (int b = 0; b < A; b++) {
        for (int a = 0; a <= out.a_star[b]; a++)
            if (Q[(size_t)b * A + a] == 0.0) return;
        if (b > 0 && out.a_star[b] > out.a_star[b - 1]) return;
        for (int a = 0; a < A; a++)
            if (Q[(size_t)b * A + a] != Q[(size_t)a * A + b]) return;
    }
    if (A == 0 || !(Q[0] > 0.0)) return;
    // q from the first column, then two sweeps of the rank-one least-squares update over the non-zero entries (averages
    // the entries' individual rounding errors)
    out.q.assign((size_t)A, 0.0);
    const double root = std::sqrt(Q[0]);
    for (int a = 0; a < A; a++) out.q[a] = Q[(size_t)a * A] / root;
    for (int sweep = 0; sweep < 2; sweep++) {
        std::vector<double> next((size_t)A, 0.0);
        for (int b = 0; b < A; b++) {
            double num = 0, den = 0;
            for (int a = 0; a <= out.a_star[b]; a++) { num += Q[(size_t)b * A + a] * out.q[a]; den += out.q[a] * out.q[a]; }
            next[b] = den > 0 ? num / den : 0.0;
        }
        out.q = next;
    }
    for (int b = 0; b < A; b++)
        for (int a = 0; a <= out.a_star[b]; a++) {
            const double model = out.q[b] * out.q[a];
            if (!(model > 0.0)) return;
            out.max_rel_dev = std::max(out.max_rel_dev, std::abs(Q[(size_t)b * A + a] - model) / model);
        }
    if (!(out.max_rel_dev <= tolerance)) return;
    out.beta_lo.assign((size_t)A, 0);
    out.beta_hi.assign((size_t)A, 0);
    for (int b = A - 1; b >= 0; b--) {
        const int a = out.a_star[b];
        if (a < 0) continue;
        if (out.beta_hi[a] == 0) out.beta_hi[a] = b + 1;
        out.beta_lo[a] = b;
    }
    out.ok = true;
}

void build_scan_pieces(const Tables& t, int bn, int warps, ScanPieces& out)
{
    const int G = t.G, n_off = t.offM.n_off;
    const int chunks = (G + bn - 1) / bn, blocks = (G + 31) / 32;
    out = ScanPieces();
    out.chunk_begin.push_back(0);
    std::vector<char> hit((size_t)n_off);
    for (int c = 0; c < chunks; c++) {
        const int q0 = c * bn, q1 = std::min(G, q0 + bn);      // columns [q0, q1)
        // the offset intervals of every y block
        std::vector<std::vector<std::pair<int, int>>> runs_of((size_t)blocks);
        long long total = 0;
        for (int yb = 0; yb < blocks; yb++) {
            std::fill(hit.begin(), hit.end(), 0);
            for (int y = yb * 32; y < std::min(G, yb * 32 + 32); y++)
                for (int o = 0; o < n_off; o++) {
                    const int u = t.axisM.upper[(size_t)y * n_off + o];
                    if (u == 0) continue;                       // sample beyond the grid
                    if ((u - 1 >= q0 && u - 1 < q1) || (u >= q0 && u < q1)) hit[(size_t)o] = 1;
                }
            // maximal runs of hit offsets; while there are too many, close the narrowest gap
            std::vector<std::pair<int, int>>& runs = runs_of[(size_t)yb];
            for (int o = 0; o < n_off; o++)
                if (hit[(size_t)o]) {
                    if (!runs.empty() && runs.back().second == o) runs.back().second = o + 1;
                    else runs.emplace_back(o, o + 1);
                }
            while ((int)runs.size() > kScanRanges) {
                size_t best = 1;
                for (size_t r = 2; r < runs.size(); r++)
                    if (runs[r].first - runs[r - 1].second < runs[best].first - runs[best - 1].second) best = r;
                runs[best - 1].second = runs[best].second;
                runs.erase(runs.begin() + (long)best);
            }
            for (const auto& r : runs) total += r.second - r.first;
        }
        // cut into pieces: every warp takes the next `quota` offsets of the chunk's list
        const int quota = (int)std::max<long long>(1, (total + warps - 1) / warps);
        const int first = (int)out.piece.size() / 4;
        std::vector<int> warp_begin((size_t)warps + 1, 0), block_begin((size_t)blocks + 1, 0);
        int w = 0, load = 0, count = 0;
        for (int yb = 0; yb < blocks; yb++) {
            block_begin[(size_t)yb] = count;
            for (auto [lo, hi] : runs_of[(size_t)yb])
                while (lo < hi) {
                    if (load == quota && w + 1 < warps) { w++; warp_begin[(size_t)w] = count; load = 0; }
                    const int take = w + 1 < warps ? std::min(quota - load, hi - lo) : hi - lo;
                    out.piece.insert(out.piece.end(), {yb, lo, lo + take, 0});
                    count++; load += take; lo += take;
                }
        }
        block_begin[(size_t)blocks] = count;
        for (int v = w + 1; v <= warps; v++) warp_begin[(size_t)v] = count;
        out.warp_begin.insert(out.warp_begin.end(), warp_begin.begin(), warp_begin.end());
        out.block_begin.insert(out.block_begin.end(), block_begin.begin(), block_begin.end());
        out.chunk_begin.push_back(first + count);
        out.max_pieces = std::max(out.max_pieces, count);
    }
}

void build_ring_operator(const Tables& t, int rank, int world, RingOperator& out)
{
    const int G = t.G, n_off = t.offM.n_off, A = t.offM.A;
    const AxisTable& ax = t.axisM;
    const int rows_per_rank = (G + world - 1) / world;
    auto slot = [&](int x) { return world == 1 ? x : (x % world) * rows_per_rank + x / world; };

    // per node: its in-square offsets ordered by decreasing abscissa (ties by offset index: deterministic)
    std::vector<std::vector<int>> order((size_t)G);
    std::vector<double> lim_min((size_t)G, 1e300);
    for (int n = 0; n < G; n++) {
        for (int o = 0; o < n_off; o++)
            if (ax.upper[(size_t)n * n_off + o] != 0) {
                order[n].push_back(o);
                lim_min[n] = std::min(lim_min[n], ax.limit[(size_t)n * n_off + o]);
            }
        std::sort(order[n].begin(), order[n].end(), [&](int a, int b) {
            const double sa = ax.s[(size_t)n * n_off + a], sb = ax.s[(size_t)n * n_off + b];
            return sa != sb ? sa > sb : a < b;
        });
    }

    const int rows = (G - rank + world - 1) / world;
    std::vector<RingOperator> per_row((size_t)rows);
    #pragma omp parallel
    {
        std::vector<double> window((size_t)G * G, 0.0);
        std::vector<char> seen((size_t)G * G, 0);
        std::vector<int> touched;
        #pragma omp for schedule(dynamic, 1)
        for (int lx = 0; lx < rows; lx++) {
            const int x = rank + world * lx;
            RingOperator& mine = per_row[lx];
            for (int y = 0; y < G; y++) {
                if (!(lim_min[x] < t.far_inside[y])) continue;
                touched.clear();
                for (int oi : order[x]) {
                    const size_t ci = (size_t)x * n_off + oi;
                    const double lim = ax.limit[ci];
                    if (!(lim < t.far_inside[y])) break;        // smaller abscissae only raise the limit
                    const int ux = ax.upper[ci];
                    const double wx = ax.weight[ci];
                    const double* q_row = t.offM.quadrant.data() + (size_t)t.offM.alpha[oi] * A;
                    for (int oj : order[y]) {
                        const size_t cj = (size_t)y * n_off + oj;
                        if (ax.s[cj] <= lim) break;             // inside the circle from here on
                        mine.samples++;
                        const double m = q_row[t.offM.alpha[oj]];
                        if (m == 0.0) continue;
                        const int uy = ax.upper[cj];
                        const double wy = ax.weight[cj];
                        const int idx[4] = {(ux - 1) * G + (uy - 1), ux * G + (uy - 1), (ux - 1) * G + uy, ux * G + uy};
                        const double w4[4] = {(1 - wx) * (1 - wy), wx * (1 - wy), (1 - wx) * wy, wx * wy};   // S:348-351
                        for (int c4 = 0; c4 < 4; c4++) {
                            if (!seen[idx[c4]]) { seen[idx[c4]] = 1; touched.push_back(idx[c4]); }
                            window[idx[c4]] += m * w4[c4];
                        }
                    }
                }
                if (touched.empty()) continue;
                std::sort(touched.begin(), touched.end());
                mine.point.push_back(lx * G + y);
                for (int n : touched) {
                    const int p = n / G, q = n - p * G;
                    mine.node.push_back(slot(p) * G + q);
                    mine.weight.push_back(window[n]);
                    window[n] = 0.0; seen[n] = 0;
                }
                mine.row_begin.push_back((int)mine.node.size());
            }
        }
    }
    out = RingOperator();
    out.row_begin.push_back(0);
    out.grid_row_begin.push_back(0);
    for (const RingOperator& r : per_row) {
        const int base = (int)out.node.size();
        out.point.insert(out.point.end(), r.point.begin(), r.point.end());
        for (int e : r.row_begin) out.row_begin.push_back(base + e);
        out.node.insert(out.node.end(), r.node.begin(), r.node.end());
        out.weight.insert(out.weight.end(), r.weight.begin(), r.weight.end());
        out.samples += r.samples;
        out.grid_row_begin.push_back((int)out.point.size());
    }
}

}  // namespace ldm
