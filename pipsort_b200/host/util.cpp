// File readers / writers and the dense linear algebra of the host pre-processing.
// Behaviour follows util.cpp / pipsort.cpp of the reference (cited per function); the code is new.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "pipsort_host.h"

namespace pipsort_host {

// pipsort.cpp:28-44: one path per non-empty line
std::vector<std::string> read_dir(const std::string& fileName) {
    std::ifstream fin(fileName.c_str());
    if (!fin) {
        std::cout << "Error: unable to open " << fileName << std::endl;
        std::exit(1);
    }
    std::vector<std::string> out;
    std::string line;
    while (fin.good()) {
        std::getline(fin, line);
        if (!line.empty()) out.push_back(line);
    }
    return out;
}

// pipsort.cpp:46-66: comma separated non-negative integers, anything else is a format error
std::vector<int> read_sigma(const std::string& text) {
    std::vector<int> sizes;
    std::string cur;
    for (char ch : text) {
        if (ch == ',') {
            sizes.push_back((int)std::stod(cur));
            cur.clear();
        } else if (std::isdigit((unsigned char)ch)) {
            cur += ch;
        } else {
            std::cout << "Error: sample size is not in the right format" << std::endl;
            std::exit(1);
        }
    }
    if (!cur.empty()) sizes.push_back((int)std::stod(cur));
    return sizes;
}

// util.cpp:86-96: whitespace separated doubles; extraction stops at the first token that is not a number
void importData(const std::string& fileName, std::vector<double>& out) {
    std::ifstream file(fileName.c_str());
    if (!file) {
        std::cout << "Unable to open file; This is why";
        std::exit(1);
    }
    double v;
    while (file >> v) out.push_back(v);
}

// util.cpp:144-159
void importDataFirstColumn(const std::string& fileName, std::vector<std::string>& out) {
    std::ifstream fin(fileName.c_str());
    std::string line, tok;
    while (std::getline(fin, line)) {
        std::istringstream iss(line);
        iss >> tok;                      // an empty line repeats the previous token, like the reference
        out.push_back(tok);
    }
}

// util.cpp:132-142
void importDataSecondColumn(const std::string& fileName, std::vector<double>& out) {
    std::ifstream fin(fileName.c_str());
    std::string line, name;
    double v = 0.0;
    while (std::getline(fin, line)) {
        std::istringstream iss(line);
        iss >> name;
        iss >> v;
        out.push_back(v);
    }
}

// util.cpp:99-126: rsid,idx0,idx1 per line
void importSnpMap(const std::string& file, int numCols, std::vector<std::string>& firstCol,
                  std::vector<std::vector<int>>& remainingCols) {
    std::ifstream f(file.c_str());
    if (!f.is_open()) {
        std::cout << "Could not open file\n";
        std::exit(1);
    }
    std::string line, word;
    while (std::getline(f, line)) {
        std::stringstream s(line);
        for (int i = 0; i < numCols; i++) {
            std::getline(s, word, ',');
            if (i == 0) firstCol.push_back(word);
            else remainingCols[i - 1].push_back(std::stoi(word));
        }
    }
}

// util.cpp:183-187: appended, default stream formatting
void export2File(const std::string& fileName, double data) {
    std::ofstream outfile(fileName.c_str(), std::ios::out | std::ios::app);
    outfile << data << std::endl;
}

// determinant by LU with partial pivoting = sign * product of the diagonal (what gsl_linalg_LU_decomp +
// gsl_linalg_LU_det compute, util.cpp:214-215)
double lu_determinant(std::vector<double> a, int n) {
    double det = 1.0;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::fabs(a[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            const double v = std::fabs(a[(size_t)i * n + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(a[(size_t)k * n + j], a[(size_t)piv * n + j]);
            det = -det;
        }
        const double p = a[(size_t)k * n + k];
        det *= p;
        if (p == 0.0) return 0.0;
        for (int i = k + 1; i < n; i++) {
            const double m = a[(size_t)i * n + k] / p;
            if (m == 0.0) continue;
            double* ri = &a[(size_t)i * n];
            const double* rk = &a[(size_t)k * n];
            for (int j = k + 1; j < n; j++) ri[j] -= m * rk[j];
        }
    }
    return det;
}

// Symmetric eigen-decomposition: Householder tridiagonalisation followed by implicit QL iterations
// (the classical tred2 / tql2 pair).  On return the COLUMNS of a are the eigenvectors, w the eigenvalues.
// Only the lower triangle of the input is read (as gsl_eigen_symmv does, util.cpp:242).
void symmetric_eigen(std::vector<double>& a, int n, std::vector<double>& w) {
    std::vector<double> e(n, 0.0);
    w.assign(n, 0.0);
    auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) A(i, j) = A(j, i);
    // --- tridiagonalise -----------------------------------------------------------------------------
    for (int i = n - 1; i > 0; i--) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; k++) scale += std::fabs(A(i, k));
            if (scale == 0.0) {
                e[i] = A(i, l);
            } else {
                for (int k = 0; k <= l; k++) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
                double f = A(i, l);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; j++) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; k++) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; k++) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; j++) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; k++) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else {
            e[i] = A(i, l);
        }
        w[i] = h;
    }
    w[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; i++) {
        const int l = i - 1;
        if (w[i] != 0.0) {
            for (int j = 0; j <= l; j++) {
                double g = 0.0;
                for (int k = 0; k <= l; k++) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; k++) A(k, j) -= g * A(k, i);
            }
        }
        w[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; j++) A(j, i) = A(i, j) = 0.0;
    }
    // --- QL with implicit shifts --------------------------------------------------------------------
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                const double dd = std::fabs(w[m]) + std::fabs(w[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) break;
                double g = (w[l + 1] - w[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = w[m] - w[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = r = std::hypot(f, g);
                    if (r == 0.0) {
                        w[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    c = g / r;
                    g = w[i + 1] - p;
                    r = (w[i] - g) * s + 2.0 * c * b;
                    w[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (int k = 0; k < n; k++) {
                        f = A(k, i + 1);
                        A(k, i + 1) = s * A(k, i) + c * f;
                        A(k, i) = c * A(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                w[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
}

// model.h:171-264 + util.cpp:195-263 for one study
Prep preprocess_study(std::vector<double>& sigma, const std::vector<double>& z, int n) {
    Prep out;
    // makeSigmaPositiveSemiDefinite: 0.01 steps until the LU determinant is > 0 (an underflow to 0 counts as failure)
    double add = 0.0;
    for (;;) {
        std::vector<double> t = sigma;
        for (int i = 0; i < n; i++) t[(size_t)i * n + i] += add;
        if (lu_determinant(t, n) > 0) break;
        add += 0.01;
    }
    for (int i = 0; i < n; i++) sigma[(size_t)i * n + i] = sigma[(size_t)i * n + i] + add;
    out.add_diag = add;
    // eigen_decomp + abs(Omega) (model.h:213-232):  B = |Omega|^(1/2) Q^T,  S' = |Omega|^(-1/2) Q^T z
    std::vector<double> Q = sigma, w;
    symmetric_eigen(Q, n, w);
    double K = 0.0, mn = n ? w[0] : 0.0;
    bool all_pos = true;
    for (int i = 0; i < n; i++) {
        double qz = 0.0;
        for (int j = 0; j < n; j++) qz += Q[(size_t)j * n + i] * z[j];
        const double om = std::fabs(w[i]);
        const double sp = qz / std::sqrt(om);
        K += sp * sp;
        mn = std::min(mn, w[i]);
        if (!(w[i] > 0)) all_pos = false;
    }
    out.K = K;
    out.min_eig = mn;
    if (!all_pos) {
        // the effective LD  B^T B = Q |Omega| Q^T  differs from sigma only when an eigenvalue is negative
        std::vector<double> T((size_t)n * n);
        for (int i = 0; i < n; i++)
            for (int k = 0; k < n; k++) T[(size_t)i * n + k] = Q[(size_t)i * n + k] * std::fabs(w[k]);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                double s = 0.0;
                for (int k = 0; k < n; k++) s += T[(size_t)i * n + k] * Q[(size_t)j * n + k];
                sigma[(size_t)i * n + j] = s;
            }
    } else {
        // symmetrise from the lower triangle (the part the eigen-solver read)
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) sigma[(size_t)i * n + j] = sigma[(size_t)j * n + i];
    }
    return out;
}

}  // namespace pipsort_host
