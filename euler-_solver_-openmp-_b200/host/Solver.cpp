// Solver.cpp -- experiment front-end of the B200-native Law-Dieckmann solver.
//
// Keeps what a user of the reference's Solver.cpp sees -- the same seven experiments (14.4 a/b/c, 14.6 a/b, custom,
// kernels test; reference menu S:1325-1333), the same Data/ input CSVs (FileReader, S:665-816) and the same
// Data/VisData.csv output (FileWriter::WritePlotData, S:820-864) that Visualiser.py parses -- and the same class
// names at the solver seam (Properties, Equation, EulerSolver, SolverResults), but every numerical step goes
// through the C ABI of libldm_b200.so (include/ldm.h) into the CUDA kernels.  There is no CPU solver in here.
//
// Differences by design: the Win32 arrow-key console menu (getch / system("cls"), S:1189-1317) is replaced by a
// portable selector (argument, or a number typed at the prompt); the 14.6 sweep is one batched call
// (ldm_solve_batch) instead of an OpenMP loop with an unsynchronised push_back (S:940-963); paths accept both
// "Data/14/4/a.csv" and the reference's literal backslash spelling.
//
//   solver 14.4.b            solver 3              solver            (interactive)
//   S:n = line n of the reference's "Euler Solver ver.5.3/Solver.cpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <string>
#include <exception>
#include <thread>
#include <vector>

#include "../../include/ldm.h"
#include "types.h"

enum KernelType { Normal };
enum ExtrapolationType { Zero, SquaredN, Alpha, Exponential };   // S:28; only SquaredN reaches the pair density

static const CalcType EPSILON = 1E-08;   // S:24

static void ldm_require(int code, const char* what)
{
    if (code != LDM_OK) {
        std::cerr << "ldm: " << what << " failed (" << code << "): " << ldm_last_error() << std::endl;
        std::abort();                    // the reference aborts through assert on every error
    }
}

//====================================================================================== model properties
struct Properties
{
    int SpeciesNumber = 0;
    Vector b, d;
    Matrix d_prime;
    VectorOfMatrix M;
    MatrixOfMatrix W;
    Matrix MaxNumForW;
    Vector MaxNumForM;
    Vector SigmaM;
    Matrix SigmaW;
    Vector MomentsGrid;
    CalcType CutoffDistance = 0;
    Vector KernelsGrid;
    CalcType KernelsGridStep = 0;

    std::shared_ptr<ldm_problem> flat;   // what the kernels consume

    // Properties::Generate (S:186-241), one or several species.
    static Properties Generate(int num, Matrix sW, Vector sM, Vector b, Vector d, Matrix dd,
                               KernelType = Normal, KernelType = Normal)
    {
        ldm_problem* raw = nullptr;
        if (num == 1) {
            ldm_require(ldm_problem_generate(sW[0][0], sM[0], b[0], d[0], dd[0][0], 0, &raw), "ldm_problem_generate");
        } else {
            std::vector<double> w, m, vb, vd, vdd;
            for (int i = 0; i < num; i++) {
                m.push_back(sM[i]); vb.push_back(b[i]); vd.push_back(d[i]);
                for (int j = 0; j < num; j++) { w.push_back(sW[i][j]); vdd.push_back(dd[i][j]); }
            }
            ldm_require(ldm_problem_generate_multi(num, w.data(), m.data(), vb.data(), vd.data(), vdd.data(), 0, &raw), "ldm_problem_generate_multi");
        }
        Properties p;
        p.flat = std::shared_ptr<ldm_problem>(raw, ldm_problem_free);
        p.SpeciesNumber = num;
        p.SigmaM = sM; p.SigmaW = sW; p.b = b; p.d = d; p.d_prime = dd;
        p.MomentsGrid = Vector(std::vector<CalcType>(raw->moments_grid, raw->moments_grid + raw->G));
        p.KernelsGrid = Vector(std::vector<CalcType>(raw->kernels_grid, raw->kernels_grid + raw->K));
        p.CutoffDistance = raw->cutoff;
        p.KernelsGridStep = raw->h;
        std::cout << "number of mid-points for pair density grid: " << raw->G
                  << "\nnumber of mid-points for kernels grid:      " << raw->K << "\n\n";        // S:85-86
        const int K = raw->K;
        p.M = VectorOfMatrix(num);
        p.W = MatrixOfMatrix(num);
        p.MaxNumForM = Vector(num);
        p.MaxNumForW = Matrix(num);
        for (int i = 0; i < num; i++) {
            Matrix m(K);
            for (int x = 0; x < K; x++)
                for (int y = 0; y < K; y++) m[x][y] = raw->M[((size_t)i * K + x) * K + y];
            p.M(i) = m;
            p.MaxNumForM[i] = num == 1 ? (CalcType)raw->mM : (CalcType)raw->mM_of[i];
            for (int j = 0; j < num; j++) {
                Matrix w(K);
                for (int x = 0; x < K; x++)
                    for (int y = 0; y < K; y++) w[x][y] = raw->W[(((size_t)i * num + j) * K + x) * K + y];
                p.W(i, j) = w;
                p.MaxNumForW[i][j] = num == 1 ? (CalcType)raw->mW : (CalcType)raw->mW_of[i * num + j];
            }
        }
        return p;
    }

    void CheckKernels()      // S:147-183
    {
        double iw = 0, im = 0;
        ldm_require(ldm_problem_check_kernels(flat.get(), &iw, &im), "ldm_problem_check_kernels");
        std::cout << "    Checking kernels...\n      W_1_1: integral = " << iw;
        assert(std::abs(iw - 1) < EPSILON);
        std::cout << ".  Cool!!\n      M_1:   integral = " << im;
        assert(std::abs(im - 1) < EPSILON);
        std::cout << ".  Cool!!\n    Check complete!\n\n";
    }
};

//============================================================================================== equation
class Equation
{
    Properties props;
    std::shared_ptr<ldm_context> handle;     // copies of an Equation share one device context

public:
    // --closure on the command line (an extension, include/ldm.h ldm_set_closure): the closure every Equation of this run
    // is evaluated with; the reference's own (power-1) unless asked otherwise
    static int& closure_of_this_run() { static int closure = LDM_CLOSURE_POWER1; return closure; }

    Equation(Properties p, int device = 0) : props(p)
    {
        ldm_handle h = nullptr;
        ldm_require(ldm_create(props.flat.get(), device, &h), "ldm_create");
        handle = std::shared_ptr<ldm_context>(h, ldm_destroy);
        if (closure_of_this_run() != LDM_CLOSURE_POWER1) ldm_require(ldm_set_closure(h, closure_of_this_run()), "ldm_set_closure");
    }

    int state_size() const { const int n = props.flat->species; return n + n * n * props.flat->G * props.flat->G; }   // S:260
    ldm_handle raw() const { return handle.get(); }
    const Properties& properties() const { return props; }

    Vector GenerateEquilibriumValuesForNAndC(CalcType value)     // S:416-428
    {
        Vector out(state_size(), value * value);
        for (int i = 0; i < props.flat->species; i++) out[i] = value;
        return out;
    }

    Vector operator()(CalcType /*times*/, Vector vals, ExtrapolationType chosen_type)     // S:431-513
    {
        assert(chosen_type == SquaredN);
        assert(vals.size() == state_size());                                            // S:260
        Vector out(state_size());
        ldm_require(ldm_rhs(handle.get(), vals.raw(), out.raw()), "ldm_rhs");
        return out;
    }
};

//================================================================================================ solver
struct SolverResults
{
    Vector Times;
    std::vector<Vector> Values;
    std::vector<Vector> Derivatives;
    void clear() { Times.clear(); Values.clear(); Derivatives.clear(); }
};

class EulerSolver
{
public:
    // EulerSolver::Solve (S:557-592): full history of states and derivatives, like the reference.
    SolverResults Solve(Equation current_equation, ExtrapolationType type, CalcType time_max, CalcType step,
                        Vector initial_value, bool debug_print_enabled = false)
    {
        assert(type == SquaredN);
        const int n = ldm_count_points(time_max, step), width = current_equation.state_size();
        std::vector<double> times((size_t)n), values((size_t)n * width), derivs((size_t)n * width);
        int got = 0;
        ldm_require(ldm_solve(current_equation.raw(), initial_value.raw(), step, time_max, LDM_RECORD_ALL, times.data(),
                              values.data(), derivs.data(), &got), "ldm_solve");
        SolverResults r;
        for (int k = 0; k < got; k++) {
            r.Times.add(times[(size_t)k]);
            r.Values.emplace_back(std::vector<CalcType>(values.begin() + (size_t)k * width, values.begin() + (size_t)(k + 1) * width));
            r.Derivatives.emplace_back(std::vector<CalcType>(derivs.begin() + (size_t)k * width, derivs.begin() + (size_t)(k + 1) * width));
            if (debug_print_enabled) std::cout << "   t = " << times[(size_t)k] << ";   N = " << values[(size_t)k * width] << '\n';
        }
        return r;
    }

    // What Experiment_14_4 actually reads (S:1009-1016): N(t_k) and C[0][G-1](t_k) only; no state history crosses PCIe.
    void SolveObservables(Equation& equation, CalcType time_max, CalcType step, Vector initial_value, Vector& times,
                          Vector& n_of_t, Vector& c_at_cutoff, bool debug_print_enabled)
    {
        const int n = ldm_count_points(time_max, step);
        std::vector<double> t((size_t)n), obs((size_t)n * 2);
        int got = 0;
        ldm_require(ldm_solve(equation.raw(), initial_value.raw(), step, time_max, LDM_RECORD_OBSERVABLES, t.data(),
                              obs.data(), nullptr, &got), "ldm_solve");
        for (int k = 0; k < got; k++) {
            times.add(t[(size_t)k]);
            n_of_t.add(obs[(size_t)2 * k]);
            c_at_cutoff.add(obs[(size_t)2 * k + 1]);
            if (debug_print_enabled) std::cout << "   t = " << t[(size_t)k] << ";   N = " << obs[(size_t)2 * k] << '\n';
        }
    }
};

//======================================================================================= input and output
// Positional CSV reader (S:665-816): the label in the first field is ignored, only the row
// position and the fields after the first comma matter.  CRLF and LF inputs both work.
class FileReader
{
    std::string filename;
    std::vector<std::vector<std::string>> rows;
    size_t cursor = 0;

    static std::string other_spelling(std::string path)
    {
        for (char& c : path) c = (c == '\\') ? '/' : (c == '/' ? '\\' : c);
        return path;
    }

    void load()
    {
        std::ifstream file(filename);
        if (!file) file.open(other_spelling(filename));
        if (!file) {
            std::string slashed = filename;
            for (char& c : slashed) if (c == '\\') c = '/';
            file.open(slashed);
        }
        if (!file) throw std::runtime_error("cannot open " + filename);
        rows.clear();
        cursor = 0;
        std::string line;
        while (std::getline(file, line)) {
            while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
            std::vector<std::string> fields;
            std::stringstream stream(line);
            std::string word;
            while (std::getline(stream, word, ',')) fields.push_back(word);
            rows.push_back(fields);
        }
    }

    const std::vector<std::string>& next_row()
    {
        static const std::vector<std::string> empty;
        return cursor < rows.size() ? rows[cursor++] : (cursor++, empty);
    }
    void skip(int n) { cursor += (size_t)n; }

    static std::vector<CalcType> numbers_after_label(const std::vector<std::string>& row)
    {
        std::vector<CalcType> out;
        for (size_t i = 1; i < row.size(); i++)
            if (row[i].find_first_not_of(" \t") != std::string::npos) out.push_back(std::stod(row[i]));
        return out;
    }

    int GetInt()                                  // S:673-682: value in field 2, then one row skipped
    {
        const auto& row = next_row();
        skip(1);
        return std::stoi(row.at(1));
    }
    Vector GetVector(int size)                    // S:694-705
    {
        Vector v(numbers_after_label(next_row()));
        assert(v.size() == size);
        return v;
    }
    Matrix GetMatrix(int size)                    // S:707-727: `size` rows, padded to three
    {
        Matrix m(size);
        for (int r = 0; r < size; r++) {
            const auto values = numbers_after_label(next_row());
            for (int c = 0; c < (int)values.size() && c < size; c++) m[r][c] = values[(size_t)c];
        }
        if (size < 3) skip(3 - size);
        return m;
    }

    static void DebugPrint(Properties& props)     // S:729-740
    {
        std::cout << "MODEL PROPERTIES:\n   Species number: " << props.SpeciesNumber << "\n   Cutoff distance : " << props.CutoffDistance
                  << "\n   Sigma W:  " << props.SigmaW << "   Sigma M:  " << props.SigmaM << "         b:  " << props.b
                  << "         d:  " << props.d << "         d': " << props.d_prime << "\n\n";
    }

public:
    FileReader(std::string str) : filename(str) {}

    Properties GetProperties(bool debug_print_enabled = false)              // S:745-767
    {
        load();
        const int species = GetInt();
        skip(1);
        Matrix sW = GetMatrix(species);
        Vector sM = GetVector(species);
        skip(1);
        Vector b = GetVector(species), d = GetVector(species);
        Matrix dd = GetMatrix(species);
        Properties p = Properties::Generate(species, sW, sM, b, d, dd);
        if (debug_print_enabled) DebugPrint(p);
        p.CheckKernels();
        return p;
    }

    Properties GetWithoutSigmaW(CalcType sigmaW, bool debug_print_enabled = false)   // S:769-792
    {
        load();
        const int species = GetInt();
        skip(1);
        Matrix sW(species, sigmaW);
        Vector sM = GetVector(species);
        skip(3);
        Vector b = GetVector(species), d = GetVector(species);
        Matrix dd = GetMatrix(species);
        Properties p = Properties::Generate(species, sW, sM, b, d, dd);
        if (debug_print_enabled) DebugPrint(p);
        return p;
    }

    Properties GetWithoutSigmaM(CalcType sigmaM, bool debug_print_enabled = false)   // S:794-815
    {
        load();
        const int species = GetInt();
        skip(1);
        Matrix sW = GetMatrix(species);
        Vector sM(species, sigmaM);
        skip(1);
        Vector b = GetVector(species), d = GetVector(species);
        Matrix dd = GetMatrix(species);
        Properties p = Properties::Generate(species, sW, sM, b, d, dd);
        if (debug_print_enabled) DebugPrint(p);
        return p;
    }
};

// Data/VisData.csv in the layout Visualiser.py reads positionally (S:838-862): nine header
// lines, the label row, one row per point; numbers at the default ostream precision (6 significant digits).
class FileWriter
{
    std::string filename;

public:
    FileWriter(std::string str) : filename(str) {}

    void WritePlotData(std::string number, std::string title, std::vector<std::string> labels, std::vector<Vector> points)
    {
        assert(points.size() >= 2);
        assert(points.size() == labels.size());
        const int number_of_plots = (int)points.size() - 1;
        const int number_of_points = points.at(0).size();
        for (size_t i = 1; i < points.size(); i++) assert(points.at(i).size() == number_of_points);

        std::ostringstream text;
        text << "Experiment," << number << "\nPlot title," << title << "\n\nNumber of," << number_of_plots << "\nplots\n\nNumber of,"
             << number_of_points << "\npoints\n\n";
        for (size_t i = 0; i < labels.size(); i++) text << (i ? "," : "") << labels[i];
        text << '\n';
        for (int p = 0; p < number_of_points; p++) {
            text << points.at(0)[p];
            for (int plot = 1; plot <= number_of_plots; plot++) text << ',' << points.at((size_t)plot)[p];
            text << '\n';
        }
        // Visualiser.py opens the literal name 'Data\VisData.csv' (V:23): on Linux that is one file name with a
        // backslash in it, so both spellings are written.
        std::string slashed = filename, literal = filename;
        for (char& c : slashed) if (c == '\\') c = '/';
        for (char& c : literal) if (c == '/') c = '\\';
        for (const std::string& name : {slashed, literal}) {
            std::ofstream file(name, std::ios::out | std::ios::trunc);
            if (file) file << text.str();
        }
    }
};

//=========================================================================================== experiments
class Experiments
{
    Properties Props;
    std::vector<Properties> PropsVector;
    std::string reader_file_path;
    std::string visdata_path;
    std::string data_dir;

    static double seconds_since(std::chrono::steady_clock::time_point t0)
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    void Visualise()                                         // S:980: WinExec("python Visualiser.py", 1)
    {
        if (std::getenv("LDM_VISUALISE")) { int rc = std::system("python Visualiser.py"); (void)rc; }
    }

    // CalculateFirstMoment (S:885-906)
    void CalculateFirstMoment(CalcType t_max, Vector& times, Vector& N, Vector& CSquareRoot)
    {
        FileReader Reader(reader_file_path);
        Props = Reader.GetProperties(true);
        Equation equation(Props);
        EulerSolver solver;
        std::cout << "\n==========================CALCULATIONS START==========================\n";
        const auto start = std::chrono::steady_clock::now();
        Vector c_cut;
        if (Props.SpeciesNumber == 1) {
            solver.SolveObservables(equation, t_max, 0.1, equation.GenerateEquilibriumValuesForNAndC(200), times, N, c_cut, true);
        } else {
            // several species: the reference plots Values[k].first() and Values[k][MomentsGrid.size()] whatever the species
            // count (S:1080-1087), i.e. N of the first species and entry G of the state vector
            const int n = ldm_count_points(t_max, 0.1), width = equation.state_size(), G = Props.MomentsGrid.size();
            Vector init = equation.GenerateEquilibriumValuesForNAndC(200);
            std::vector<double> t((size_t)n), values((size_t)n * width);
            int got = 0;
            ldm_require(ldm_solve(equation.raw(), init.raw(), 0.1, t_max, LDM_RECORD_VALUES, t.data(), values.data(), nullptr, &got), "ldm_solve");
            for (int k = 0; k < got; k++) {
                times.add(t[(size_t)k]);
                N.add(values[(size_t)k * width]);
                c_cut.add(values[(size_t)k * width + G]);
                std::cout << "   t = " << t[(size_t)k] << ";   N = " << values[(size_t)k * width] << '\n';
            }
        }
        std::cout << "\n===========================CALCULATIONS END===========================\n";
        std::cout << "TIME:  " << Time(seconds_since(start)) << '\n';
        for (int k = 0; k < c_cut.size(); k++) CSquareRoot.add(std::sqrt(c_cut[k]));     // S:1014-1015
    }

    void single_solve_experiment(const std::string& number, const std::string& title)
    {
        Vector times, N, CSquareRoot;
        CalculateFirstMoment(100, times, N, CSquareRoot);
        FileWriter VisData(visdata_path);
        VisData.WritePlotData(number, title, {"Time", "N", "Square root of C at cutoff distance"}, {times, N, CSquareRoot});
        Visualise();
    }

public:
    Experiments(std::string data_directory = "Data") : data_dir(data_directory) { visdata_path = data_dir + "/VisData.csv"; }

    void Experiment_14_4(int input_data)                     // S:987-1031
    {
        static const char* files[] = {"a", "b", "c"};
        reader_file_path = data_dir + "/14/4/" + files[input_data] + ".csv";
        single_solve_experiment("14.4", std::string("(") + files[input_data] + ")");
    }

    // devices the sweep may use: all visible ones, or LDM_GPUS of them
    static int gpu_count()
    {
        int count = std::max(1, ldm_device_count());
        if (const char* env = std::getenv("LDM_GPUS")) count = std::max(1, std::min(count, std::atoi(env)));
        return count;
    }

    // cost of one Euler step of an entry ~ G^2 * (quadrature offsets per axis): both phases of the kernel walk that many
    // (node, offset) pairs per row.  Longest-processing-time-first onto the least loaded device.
    static std::vector<int> assign_devices(const std::vector<Properties>& props, int devices)
    {
        const int n = (int)props.size();
        std::vector<double> cost((size_t)n);
        std::vector<int> order((size_t)n), device_of((size_t)n, 0);
        for (int i = 0; i < n; i++) {
            const ldm_problem& f = *props[(size_t)i].flat;
            cost[(size_t)i] = (double)f.G * f.G * std::max(1, f.mM - 1) + 2000.0;     // + a floor for launch-bound entries
            order[(size_t)i] = i;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[(size_t)a] > cost[(size_t)b]; });
        std::vector<double> load((size_t)devices, 0.0);
        for (int i : order) {
            const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            device_of[(size_t)i] = d;
            load[(size_t)d] += cost[(size_t)i];
        }
        if (devices > 1) {
            std::cout << "sweep over " << devices << " devices, estimated load per device:";
            for (double l : load) std::cout << ' ' << l / 1e6;
            std::cout << " (1e6 pair-steps)\n";
        }
        return device_of;
    }

    // Experiment_14_6 (S:1033-1071) over CalculateWithVaryingSigma (S:909-977): 160 independent solves of 2001 points.
    void Experiment_14_6(char SigmaType, std::string RadiusType)
    {
        reader_file_path = data_dir + (SigmaType == 'M' ? "/14/6/b.csv" : "/14/6/a.csv");
        const std::string plot_title = SigmaType == 'M' ? "(b)" : "(a)";
        Vector SigmaVec = Vector::setLinSpaced(160, EPSILON, 0.2);                       // S:1042

        PropsVector.clear();
        for (int i = 0; i < SigmaVec.size(); i++) {
            FileReader Reader(reader_file_path);
            std::cout << "ENTRY " << i << ":\n";
            PropsVector.push_back(SigmaType == 'W' ? Reader.GetWithoutSigmaW(SigmaVec[i], true) : Reader.GetWithoutSigmaM(SigmaVec[i], true));
        }
        std::cout << "\n==========================CALCULATIONS START==========================\n";
        const auto start = std::chrono::steady_clock::now();
        const int n = (int)PropsVector.size();
        // Where the reference hands the entries to OpenMP threads in index order (static schedule: the twelve refined-grid
        // entries, 100x the cost of the others, all land on thread 0), the entries are dealt out over the visible GPUs by
        // estimated cost, largest first onto the least loaded device; each device then advances its slice as one batch.
        const std::vector<int> device_of = assign_devices(PropsVector, gpu_count());
        std::vector<Equation> equations;
        {
            std::vector<std::unique_ptr<Equation>> built((size_t)n);
            std::vector<std::thread> workers;
            const int n_workers = std::min(8, std::max(1, n));
            // a library error inside a worker (no device, out of memory, an entry the library rejects) is carried back to
            // this thread and thrown here; an exception that escaped a worker would end the program without its message
            std::vector<std::exception_ptr> failure((size_t)n_workers);
            for (int w = 0; w < n_workers; w++)
                workers.emplace_back([&, w] {
                    try {
                        for (int i = w; i < n; i += n_workers) built[(size_t)i].reset(new Equation(PropsVector[(size_t)i], device_of[(size_t)i]));
                    } catch (...) {
                        failure[(size_t)w] = std::current_exception();
                    }
                });
            for (std::thread& t : workers) t.join();
            for (const std::exception_ptr& e : failure) if (e) std::rethrow_exception(e);
            for (int i = 0; i < n; i++) equations.push_back(*built[(size_t)i]);
        }
        const double setup_seconds = seconds_since(start);
        std::vector<Vector> inits, lasts;
        std::vector<ldm_handle> handles;
        std::vector<const double*> in_ptrs;
        std::vector<double*> out_ptrs;
        for (int i = 0; i < n; i++) {
            inits.push_back(equations[(size_t)i].GenerateEquilibriumValuesForNAndC(200));
            lasts.emplace_back(equations[(size_t)i].state_size());
        }
        for (int i = 0; i < n; i++) {
            handles.push_back(equations[(size_t)i].raw());
            in_ptrs.push_back(inits[(size_t)i].raw());
            out_ptrs.push_back(lasts[(size_t)i].raw());
        }
        ldm_require(ldm_solve_batch(n, handles.data(), in_ptrs.data(), 0.1, 200, out_ptrs.data()), "ldm_solve_batch");
        std::cout << "\n===========================CALCULATIONS END===========================\n";
        const double elapsed = seconds_since(start);
        std::cout << "TIME:  " << Time(elapsed) << '\n';
        std::cout << "SECONDS:  " << elapsed << "  (" << n << " entries; " << setup_seconds << " s of it setting the entries up on the device)\n";

        Vector ConvergedN, ConvergedCSquareRoot;
        for (int i = 0; i < n; i++) {
            ConvergedN.add(lasts[(size_t)i][0]);                                                          // S:1051
            ConvergedCSquareRoot.add(std::sqrt(lasts[(size_t)i][PropsVector[(size_t)i].MomentsGrid.size()]));   // S:1054
            std::cout << "ENTRY " << i << ": sigma = " << SigmaVec[i] << ";  N = " << lasts[(size_t)i][0] << '\n';
        }
        FileWriter VisData(visdata_path);
        VisData.WritePlotData("14.6", plot_title, {RadiusType, "Equilibrium density", "Square root of C at cutoff distance"},
                              {SigmaVec, ConvergedN, ConvergedCSquareRoot});
        Visualise();
    }

    // Host-only: parse the experiment's input file and print the generated model (no device needed).
    void Describe(int index)
    {
        static const char* files[] = {"/14/4/a.csv", "/14/4/b.csv", "/14/4/c.csv", "/14/6/a.csv", "/14/6/b.csv", "/Custom/experiment.csv"};
        if (index < 0 || index > 5) { std::cout << "nothing to describe\n"; return; }
        FileReader Reader(data_dir + files[index]);
        Vector sigmas = Vector::setLinSpaced(160, EPSILON, 0.2);
        Properties p = index == 3 ? Reader.GetWithoutSigmaW(sigmas[80], true) : index == 4 ? Reader.GetWithoutSigmaM(sigmas[80], true) : Reader.GetProperties(true);
        std::cout.precision(17);
        std::cout << "DESCRIBE G=" << p.MomentsGrid.size() << " K=" << p.KernelsGrid.size() << " mM=" << p.MaxNumForM[0] << " mW=" << p.MaxNumForW[0][0]
                  << " h=" << p.KernelsGridStep << " cutoff=" << p.CutoffDistance << " sigmaW=" << p.SigmaW[0][0] << " sigmaM=" << p.SigmaM[0]
                  << " b=" << p.b[0] << " d=" << p.d[0] << " dd=" << p.d_prime[0][0] << '\n';
    }

    void Custom_Experiment()                                 // S:1073-1103
    {
        reader_file_path = data_dir + "/Custom/experiment.csv";
        single_solve_experiment("1", "My experiment");
    }

    // KernelTest (S:1105-1184): the radial profile of one kernel table, distinct radii in ascending order
    void KernelTest(char KernelSymbol, std::string case_str)
    {
        FileReader Reader(data_dir + "/14/4/" + case_str + ".csv");
        Properties props = Reader.GetProperties(true);
        Matrix kernel = KernelSymbol == 'W' ? props.W(0, 0) : props.M(0);
        const std::string title = std::string(KernelSymbol == 'W' ? "w" : "m") + "(r) kernel test for experiment " + case_str;

        std::vector<std::pair<CalcType, CalcType>> samples;
        const int K = props.KernelsGrid.size();
        for (int x = 0; x < K; x++)
            for (int y = 0; y < K; y++) samples.emplace_back(std::hypot(props.KernelsGrid[x], props.KernelsGrid[y]), kernel[x][y]);
        // the reference keeps the FIRST table entry seen for each distinct radius, then sorts by radius
        std::stable_sort(samples.begin(), samples.end(),
                         [](const std::pair<CalcType, CalcType>& a, const std::pair<CalcType, CalcType>& b) { return a.first < b.first; });
        Vector xValues, yValues;
        for (size_t i = 0; i < samples.size(); i++)
            if (i == 0 || samples[i].first != samples[i - 1].first) { xValues.add(samples[i].first); yValues.add(samples[i].second); }
        std::cout << "MAX NON-ZERO NUMBERS: ";
        if (KernelSymbol == 'W') std::cout << props.MaxNumForW << '\n'; else std::cout << props.MaxNumForM << '\n';
        FileWriter VisData(visdata_path);
        VisData.WritePlotData("1", title, {"r", "distribution"}, {xValues, yValues});
        Visualise();
    }
};

//================================================================================================== menu
static const std::vector<std::string> kCommands = {"14.4.a", "14.4.b", "14.4.c", "14.6.a", "14.6.b", "Custom experiment", "Kernels Test"};   // S:1325-1333

static int parse_choice(std::string text)
{
    while (!text.empty() && isspace((unsigned char)text.back())) text.pop_back();
    if (text.empty()) return -1;
    if (text == "q" || text == "quit" || text == "esc") return -2;
    for (size_t i = 0; i < kCommands.size(); i++)
        if (text == kCommands[i]) return (int)i;
    if (text == "custom") return 5;
    if (text == "kernels") return 6;
    char* end = nullptr;
    const long v = std::strtol(text.c_str(), &end, 10);
    return (end && *end == 0 && v >= 0 && v < (long)kCommands.size()) ? (int)v : -1;
}

static void run_choice(Experiments& run, int index)          // the switch of S:1259-1308
{
    switch (index) {
        case 0: case 1: case 2: run.Experiment_14_4(index); break;
        case 3: run.Experiment_14_6('W', "Competition radius"); break;
        case 4: run.Experiment_14_6('M', "Dispersal radius"); break;
        case 5: run.Custom_Experiment(); break;
        case 6:
            for (char kernel : {'W', 'M'})
                for (const char* c : {"a", "b", "c"}) run.KernelTest(kernel, c);
            std::cout << "Tests Complete!\n";
            break;
        default: std::cout << "ERROR: No such experiment!!\n";
    }
}

int main(int argc, char** argv)
{
    std::string data_dir = "Data";
    std::vector<std::string> args;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--data-dir" && i + 1 < argc) data_dir = argv[++i];
        else if (a == "--describe" && i + 1 < argc) { Experiments(data_dir).Describe(parse_choice(argv[++i])); return 0; }
        else if (a == "--closure" && i + 1 < argc) {
            const std::string name = argv[++i];
            const int closure = name == "power1" ? LDM_CLOSURE_POWER1 : name == "power2" ? LDM_CLOSURE_POWER2_SYMMETRIC : name == "power3" ? LDM_CLOSURE_POWER3 : 0;
            if (closure == 0) { std::cerr << "unknown closure '" << name << "' (power1, power2, power3)\n"; return 2; }
            Equation::closure_of_this_run() = closure;
        }
        else args.push_back(a);
    }
    Experiments run(data_dir);
    if (!args.empty()) {
        std::string joined = args[0];
        for (size_t i = 1; i < args.size(); i++) joined += " " + args[i];
        const int index = parse_choice(joined);
        if (index < 0) { std::cerr << "unknown experiment '" << joined << "'\n"; return 2; }
        run_choice(run, index);
        return 0;
    }
    for (;;) {
        std::cout << "Welcome!\nPlease choose experiment:\n";
        for (size_t i = 0; i < kCommands.size(); i++) std::cout << "  " << i << ") " << kCommands[i] << '\n';
        std::cout << "\n(type the number or the name and press ENTER; q to quit)\n> " << std::flush;
        std::string line;
        if (!std::getline(std::cin, line)) break;
        const int index = parse_choice(line);
        if (index == -2) break;
        if (index >= 0) run_choice(run, index);
    }
    return 0;
}
