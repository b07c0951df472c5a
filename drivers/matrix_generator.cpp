// matrix_generator -- writes the text input of the benchmark driver.
// Same command line and file format as the reference generator (/root/reference/matrix_generator.cpp:8-11,53-85):
//     matrix_generator filename maxSize [step=2] [function=exp (exp/lin)] [sparsity=0.0] [kind=rand (rand/dd/spd:KAPPA[:SEED])]
// first line = number of matrices (written last, over a 16-character placeholder), then per matrix its size n and
// n rows of n values.  With kind=rand the libc rand() stream is consumed exactly like the reference does (one draw
// per element when sparsity == 0, an extra draw per element otherwise; never seeded), so the files are byte-identical.
// kind=dd (new) keeps the same draws but replaces the diagonal so that the matrix the benchmark driver ACTUALLY
// factors -- it reads the values sequentially as column-major storage (benchmark.cpp:192-194), i.e. the transpose of
// what is printed -- is strictly column diagonally dominant: partial pivoting then never swaps and the mixed-
// precision no-pivot solver applies.
// kind=spd:KAPPA[:SEED] (new) writes symmetric positive definite matrices H diag(sigma) H with one Householder
// reflector H = I - 2uu^T and singular values geometric in [1/KAPPA, 1] (BASELINE.json's condition-number sweep); it
// uses its own seeded generator and prints 17 significant digits; libc rand() is not touched.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cout << "Usage: " << argv[0]
                  << " filename maxSize [step=2] [function=exp (exp/lin)] [sparsity=0.0] [kind=rand (rand/dd)]" << std::endl;
        std::cout << "  sparsity: fraction of zeros in the matrix (0.0 = dense, 0.9 = 90% zeros)" << std::endl;
        std::cout << "  kind:     rand = reference distribution; dd = same draws, diagonally dominant" << std::endl;
        return -1;
    }
    std::ofstream out(argv[1]);
    if (!out.is_open()) {
        std::cout << "Failed to open " << argv[1] << std::endl;
        return -1;
    }
    const int max_size = std::atoi(argv[2]);
    if (max_size <= 0) {
        std::cout << "Invalid maxSize: " << max_size << std::endl;
        return -1;
    }
    int step = 2;
    if (argc > 3 && (step = std::atoi(argv[3])) <= 0) {
        std::cout << "Invalid step: " << step << std::endl;
        return -1;
    }
    bool geometric = true;
    if (argc > 4) {
        const std::string f = argv[4];
        if (f == "lin") geometric = false;
        else if (f != "exp") {
            std::cout << "Invalid function: " << f << ". Use 'exp' or 'lin'." << std::endl;
            return -1;
        }
    }
    double sparsity = 0.0;
    if (argc > 5) {
        sparsity = std::atof(argv[5]);
        if (sparsity < 0.0 || sparsity >= 1.0) {
            std::cout << "Invalid sparsity: " << sparsity << ". Must be in [0.0, 1.0)." << std::endl;
            return -1;
        }
    }
    bool dominant = false;
    double kappa = 0.0;  // > 0: SPD mode
    uint64_t seed = 1;
    if (argc > 6) {
        const std::string k = argv[6];
        if (k == "dd") dominant = true;
        else if (k.rfind("spd:", 0) == 0) {
            char* end = nullptr;
            kappa = std::strtod(k.c_str() + 4, &end);
            if (end && *end == ':') seed = std::strtoull(end + 1, nullptr, 10);
            if (!(kappa >= 1.0)) {
                std::cout << "Invalid kind: " << k << ". spd needs KAPPA >= 1." << std::endl;
                return -1;
            }
        } else if (k != "rand") {
            std::cout << "Invalid kind: " << k << ". Use 'rand', 'dd' or 'spd:KAPPA[:SEED]'." << std::endl;
            return -1;
        }
    }

    out << std::string(16, ' ') << std::endl;  // room for the matrix count
    int count = 0;
    std::vector<double> row;
    for (int n = 2; n <= max_size; n = geometric ? n * step : n + step) {
        out << n << std::endl;
        row.resize(n);
        if (kappa > 0.0) {
            // A = D - 2 u (Du)^T - 2 (Du) u^T + 4 (u^T D u) u u^T,  D = diag(sigma),  O(n^2)
            std::vector<double> u(n), du(n), sig(n);
            uint64_t x = seed * 0x9E3779B97F4A7C15ull + (uint64_t)n;
            double nrm = 0.0;
            for (int i = 0; i < n; ++i) {  // splitmix64 -> uniform (-1, 1)
                x += 0x9E3779B97F4A7C15ull;
                uint64_t z = x;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                u[i] = (double)(z >> 11) / 9007199254740992.0 * 2.0 - 1.0;
                nrm += u[i] * u[i];
            }
            nrm = std::sqrt(nrm);
            double udu = 0.0;
            for (int i = 0; i < n; ++i) {
                u[i] /= nrm;
                sig[i] = std::pow(kappa, -(double)i / (double)(n > 1 ? n - 1 : 1));
                du[i] = sig[i] * u[i];
                udu += u[i] * du[i];
            }
            out << std::setprecision(17);
            for (int i = 0; i < n; ++i) {
                for (int j = 0; j < n; ++j) {
                    double v = -2.0 * u[i] * du[j] - 2.0 * du[i] * u[j] + 4.0 * udu * u[i] * u[j];
                    if (i == j) v += sig[i];
                    out << v << " ";
                }
                out << std::endl;
            }
            out << std::setprecision(6) << std::endl;
            ++count;
            continue;
        }
        for (int i = 0; i < n; ++i) {
            double off = 0.0;
            for (int j = 0; j < n; ++j) {
                double v;
                // short-circuit keeps the reference's draw count: no sparsity draw when sparsity == 0
                if (sparsity > 0.0 && (static_cast<double>(std::rand()) / (RAND_MAX + 1.0)) < sparsity) v = 0.0;
                else v = static_cast<double>(std::rand() % 100) / 10.0;
                row[j] = v;
                if (j != i) off += v;  // values are non-negative
            }
            if (dominant) row[i] = off + 1.0;
            for (int j = 0; j < n; ++j) out << row[j] << " ";
            out << std::endl;
        }
        out << std::endl;
        ++count;
        std::cout << "Generating matrix of size " << (geometric ? n * step : n + step) << "\r" << std::flush;
    }
    out.seekp(0, std::ios::beg);
    out << count;
    out.close();
    std::cout << "\nnumber of matrices: " << count << std::endl;
    return 0;
}
