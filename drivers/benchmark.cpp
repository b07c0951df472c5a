// benchmark -- times MPF() against host LAPACK on the matrices of a generator file and self-checks both.
// Drop-in for the reference driver (/root/reference/benchmark.cpp:146-270): same command line
//     benchmark filename [-v] [--no-check]
// same input format (count, then n and n*n values consumed sequentially as column-major storage, :186-194), same
// check (P*L*U == A element-wise, absolute tolerance 1e-10, :97-144) and the same CSV, "benchmark_times.csv" with
// header matrix_size,mpf_time,lapack_time and 10 fixed decimals (:168-169,265).
// and the same stdout lines and exit codes on that default path (:150,:226,:236; -1 on a bad command line / file).
// New, behind flags only:
//     --solve        also solve A x = A*1 with the mixed-precision LU + iterative refinement (mplu_gesv_host) and
//                    append the columns mplu_time,iters,backward_error,mplu_tflops to the CSV
//     --csv path     write the CSV somewhere else
//     --bin          `filename` is binary: int32 count, then per matrix int32 n and n*n float64 values (column-major):
//                    the text format costs ~20 bytes and a strtod per value, which keeps n = 32768 out of reach
//     --gen dd:N[:SEED]   no input file at all: the column-dominant synthetic system of mplu_generate (the reference
//                    generator's value set) is built ON THE DEVICE and solved there (mplu_gesv_device); MPF() and LAPACK
//                    are skipped (their columns are nan).  `benchmark --gen dd:32768 --solve` reproduces the headline.
// Host LAPACK/CBLAS are loaded at run time (dlopen) so the driver builds without lapacke.h: MPLU_LAPACK_LIB or the
// library path baked in at build time (scipy's OpenBLAS in this image, symbols prefixed scipy_); when none is found
// the lapack_time column is nan and products fall back to a plain triple loop.
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

#include "MPF.h"
#include "mplu.h"

namespace {

using dgetrf_fn = int (*)(int, int, int, double*, int, int*);
using dgemm_fn = void (*)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);

struct HostBlas {
    dgetrf_fn dgetrf = nullptr;
    dgemm_fn dgemm = nullptr;
} g_blas;

void load_host_blas() {
    const char* cands[] = {std::getenv("MPLU_LAPACK_LIB"),
#ifdef MPLU_DEFAULT_LAPACK_LIB
                           MPLU_DEFAULT_LAPACK_LIB,
#endif
                           "liblapacke.so.3", "liblapacke.so", "libopenblas.so.0", "libopenblas.so"};
    for (const char* c : cands) {
        if (!c || !*c) continue;
        void* h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (!h) continue;
        for (const char* pre : {"", "scipy_"}) {
            if (!g_blas.dgetrf) g_blas.dgetrf = (dgetrf_fn)dlsym(h, (std::string(pre) + "LAPACKE_dgetrf").c_str());
            if (!g_blas.dgemm) g_blas.dgemm = (dgemm_fn)dlsym(h, (std::string(pre) + "cblas_dgemm").c_str());
        }
        if (g_blas.dgetrf && g_blas.dgemm) return;
    }
}

// the L / U dump of the reference's print_LU (benchmark.cpp:27-57)
void show_factors(const double* lu, int n, bool verbose) {
    if (!verbose || n >= 10) return;
    for (int pass = 0; pass < 2; ++pass) {
        std::cout << (pass == 0 ? "L matrix:" : "U matrix:") << std::endl;
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) {
                if (pass == 0) {
                    if (i > j) std::cout << lu[(size_t)j * n + i] << " ";
                    else std::cout << (i == j ? "1 " : "0 ");
                } else {
                    if (i <= j) std::cout << lu[(size_t)j * n + i] << " ";
                    else std::cout << "0 ";
                }
            }
            std::cout << std::endl;
        }
        std::cout << std::endl;
    }
}

void show_matrix(const char* title, const double* a, int n, bool verbose) {
    if (!verbose || n >= 10) return;
    std::cout << title << std::endl;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) std::cout << a[(size_t)j * n + i] << " ";
        std::cout << std::endl;
    }
    std::cout << std::endl;
}

// C = L * U with L, U unpacked from the dgetrf-layout array
void product_of_factors(const double* lu, double* c, int n) {
    std::vector<double> L((size_t)n * n, 0.0), U((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            const double v = lu[(size_t)j * n + i];
            if (i > j) L[(size_t)j * n + i] = v;
            else U[(size_t)j * n + i] = v;
            if (i == j) L[(size_t)j * n + i] = 1.0;
        }
    if (g_blas.dgemm) {
        g_blas.dgemm(102 /*ColMajor*/, 111, 111, n, n, n, 1.0, L.data(), n, U.data(), n, 0.0, c, n);
        return;
    }
    std::fill(c, c + (size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j)
        for (int k = 0; k <= j; ++k) {
            const double u = U[(size_t)j * n + k];
            if (u == 0.0) continue;
            for (int i = k; i < n; ++i) c[(size_t)j * n + i] += L[(size_t)k * n + i] * u;
        }
}

// P*L*U == A ?  (swaps applied last to first, as the reference's row_permute)
bool factors_reproduce(const double* a, const double* lu, const int* ipiv, int n, bool verbose) {
    std::vector<double> plu((size_t)n * n);
    show_factors(lu, n, verbose);
    product_of_factors(lu, plu.data(), n);
    show_matrix("LU matrix:", plu.data(), n, verbose);
    for (int i = n - 1; i >= 0; --i) {
        const int p = ipiv[i] - 1;
        if (p != i)
            for (int j = 0; j < n; ++j) std::swap(plu[(size_t)j * n + i], plu[(size_t)j * n + p]);
    }
    show_matrix("PLU matrix:", plu.data(), n, verbose);
    bool ok = true;
    for (size_t e = 0; e < (size_t)n * n && ok; ++e) ok = std::fabs(a[e] - plu[e]) <= 1e-10;
    if (verbose) std::cout << "Correctitude: " << (ok ? "True" : "False") << std::endl;
    return ok;
}

double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

// LU+IR of the synthetic system generated on the device (no host copy of A at all)
int run_generated(const std::string& spec, const std::string& csv_path) {
    int n = 0;
    unsigned long long seed = 1;
    if (std::sscanf(spec.c_str(), "dd:%d:%llu", &n, &seed) < 1 || n <= 0) {
        std::cout << "Invalid --gen spec " << spec << " (dd:N[:SEED])" << std::endl;
        return -1;
    }
    mplu_context* ctx = nullptr;
    if (mplu_create(&ctx, 0) != 0) {
        std::cout << "mplu: no CUDA device" << std::endl;
        return -1;
    }
    double *dA = nullptr, *db = nullptr, *dx = nullptr;
    if (mplu_device_alloc((void**)&dA, (size_t)n * n * sizeof(double)) != 0 || mplu_device_alloc((void**)&db, n * sizeof(double)) != 0 ||
        mplu_device_alloc((void**)&dx, n * sizeof(double)) != 0 || mplu_generate(n, seed, 1, dA, n, db, mplu_stream(ctx)) != 0) {
        std::cout << "mplu: device allocation / generation failed for n = " << n << std::endl;
        return -1;
    }
    mplu_stats st;
    int rc = 0;
    double best = 1e30;
    for (int rep = 0; rep < 4 && (rc == 0 || rc == MPLU_E_NOCONV); ++rep) {  // the first call captures the schedule
        rc = mplu_gesv_device(ctx, n, dA, n, db, dx, nullptr, &st);
        if (rep > 0 && st.total_ms * 1e-3 < best) best = st.total_ms * 1e-3;
    }
    std::ofstream csv(csv_path);
    csv << "matrix_size,mpf_time,lapack_time,mplu_time,iters,backward_error,mplu_tflops\n" << std::fixed << std::setprecision(10);
    if (rc != 0 && rc != MPLU_E_NOCONV) {
        std::cout << "mplu_gesv_device failed with code " << rc << std::endl;
        return -1;
    }
    std::vector<double> x(n);
    mplu_device_to_host(x.data(), dx, n * sizeof(double));
    double err = 0.0;
    for (int i = 0; i < n; ++i) err = std::fmax(err, std::fabs(x[i] - 1.0));
    const double tflops = 2.0 / 3.0 * (double)n * n * n / best / 1e12;
    std::cout << "Matriz tamanyo: " << n << std::endl;
    std::cout << "mplu LU+IR (device-resident, generated dd seed " << seed << "): " << best << " s = " << tflops << " TFLOP/s, " << st.iters
              << " refinement iterations, backward error " << std::scientific << st.backward_error << ", max|x-1| " << err
              << std::defaultfloat << (rc ? "  (did not converge)" : "") << std::endl;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    csv << n << "," << nan << "," << nan << "," << best << "," << st.iters << "," << std::scientific << st.backward_error << std::fixed
        << "," << tflops << std::endl;
    mplu_device_free(dA); mplu_device_free(db); mplu_device_free(dx);
    mplu_destroy(ctx);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    bool verbose = false, check = true, solve = false, binary = false;
    std::string csv_path = "benchmark_times.csv", gen;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-v") verbose = true;
        else if (a == "--no-check") check = false;
        else if (a == "--solve") solve = true;
        else if (a == "--bin") binary = true;
        else if (a == "--csv" && i + 1 < argc) csv_path = argv[++i];
        else if (a == "--gen" && i + 1 < argc) gen = argv[++i];
    }
    if (!gen.empty()) return run_generated(gen, csv_path);
    if (argc < 2 || argv[1][0] == '-') {
        std::cout << "Usage: " << argv[0] << " filename [-v] [--no-check]   (also: [--solve] [--bin] [--csv path] | --gen dd:N[:SEED])" << std::endl;
        return -1;
    }
    std::ifstream in(argv[1], binary ? std::ios::binary : std::ios::in);
    if (!in.is_open()) {
        std::cout << "Failed to open " << argv[1] << std::endl;
        return -1;
    }
    auto read_int = [&](int& v) -> bool {
        if (binary) { int32_t t = 0; in.read(reinterpret_cast<char*>(&t), sizeof(t)); v = t; return (bool)in; }
        return (bool)(in >> v);
    };
    load_host_blas();
    std::ofstream csv(csv_path);
    csv << "matrix_size,mpf_time,lapack_time" << (solve ? ",mplu_time,iters,backward_error,mplu_tflops" : "") << "\n"
        << std::fixed << std::setprecision(10);
    int count = 0;
    if (!read_int(count) || count <= 0) {
        std::cout << "Invalid number of matrices in " << argv[1] << std::endl;
        return -1;
    }
    if (verbose) std::cout << "Number of matrices: " << count << std::endl;
    mplu_context* ctx = nullptr;

    for (int m = 0; m < count; ++m) {
        int n = 0;
        if (!read_int(n) || n <= 0) {
            std::cout << "Invalid matrix size in " << argv[1] << " n: " << n << std::endl;
            return -1;
        }
        const size_t nn = (size_t)n * n;
        std::vector<double> a(nn);
        if (binary) {
            in.read(reinterpret_cast<char*>(a.data()), (std::streamsize)(nn * sizeof(double)));
        } else {
            for (size_t e = 0; e < nn && in; ++e) in >> a[e];
        }
        if (!in) {
            std::cout << "Error while reading matrix data in " << argv[1] << std::endl;
            return -1;
        }
        show_matrix("Original matrix:", a.data(), n, verbose);
        std::vector<double> w(a);
        std::vector<int> ipiv(n);
        for (int i = 0; i < n; ++i) ipiv[i] = i + 1;  // MPF leaves the entry of a trailing 1x1 panel untouched

        auto t0 = std::chrono::high_resolution_clock::now();
        MPF(w.data(), n, 32, ipiv.data());
        const double mpf_time = seconds_since(t0);
        if (verbose) std::cout << "MPF() time: " << mpf_time << " seconds\n" << std::endl;
        if (check) {
            std::cout << "Checking correctness of MPF results..." << std::endl;
            if (!factors_reproduce(a.data(), w.data(), ipiv.data(), n, verbose))
                std::cout << "MPF produced incorrect results." << std::endl;
        }
        std::cout << "Matriz tamanyo: " << n << std::endl;  // sic: the reference's line (benchmark.cpp:236)

        double lapack_time = std::numeric_limits<double>::quiet_NaN();
        if (g_blas.dgetrf) {
            w = a;
            t0 = std::chrono::high_resolution_clock::now();
            const int info = g_blas.dgetrf(102 /*ColMajor*/, n, n, w.data(), n, ipiv.data());
            lapack_time = seconds_since(t0);
            if (info != 0) std::cout << "LAPACKE_dgetrf failed with error code " << info << std::endl;
            else if (verbose) std::cout << "LAPACKE_dgetrf time: " << lapack_time << " seconds\n" << std::endl;
            if (check && info == 0 && !factors_reproduce(a.data(), w.data(), ipiv.data(), n, verbose))
                std::cout << "LAPACKE_dgetrf produced incorrect results." << std::endl;
        }
        csv << n << "," << mpf_time << "," << lapack_time;

        if (solve) {
            double mplu_time = std::numeric_limits<double>::quiet_NaN(), be = mplu_time;
            int iters = -1;
            if (!ctx && mplu_create(&ctx, 0) != 0) ctx = nullptr;
            if (ctx) {
                std::vector<double> b(n, 0.0), x(n, 0.0);
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < n; ++i) b[i] += a[(size_t)j * n + i];  // b = A * ones
                mplu_stats st;
                t0 = std::chrono::high_resolution_clock::now();
                const int rc = mplu_gesv_host(ctx, n, a.data(), n, b.data(), x.data(), nullptr, &st);
                mplu_time = seconds_since(t0);
                if (rc == 0 || rc == MPLU_E_NOCONV) {
                    iters = st.iters;
                    be = st.backward_error;
                    double err = 0.0;
                    for (int i = 0; i < n; ++i) err = std::fmax(err, std::fabs(x[i] - 1.0));
                    std::cout << "mplu LU+IR: " << mplu_time << " s, " << iters << " refinement iterations, backward error "
                              << std::scientific << be << ", max|x-1| " << err << std::defaultfloat
                              << (rc ? "  (did not converge)" : "") << std::endl;
                } else {
                    std::cout << "mplu_gesv_host failed with code " << rc << std::endl;
                }
            } else {
                std::cout << "mplu: no CUDA device" << std::endl;
            }
            csv << "," << mplu_time << "," << iters << "," << std::scientific << be << std::fixed << ","
                << 2.0 / 3.0 * (double)n * n * n / mplu_time / 1e12;
        }
        csv << std::endl;
    }
    if (ctx) mplu_destroy(ctx);
    return 0;
}
