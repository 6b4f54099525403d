// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_aic.cpp header).
//
// CPU restatement of the dense solvers of common/linalg.f90 and of the wrapper
// panel_solver_solve_system (src/panel_solver.f90:1802-2027):
//   lu_solve / lu_decomp / lu_back_sub        linalg.f90:118-342
//   decompose_blocks                          linalg.f90:376-456
//   block_ssor_solve                          linalg.f90:459-598
//   block_jacobi_solve                        linalg.f90:601-728
//   purcell_solve                             linalg.f90:731-794
//   get_lower_bandwidth                       linalg.f90:797-835
//   gen/apply Givens, QR_givens_solve_UP      linalg.f90:838-927
//   upper_triangular_back_sub                 linalg.f90:930-965
//   fast Givens, QR_fast_givens_solve_upper_pentagonal   linalg.f90:968-1165
//   arnoldi_update, GMRES, restarted_GMRES    linalg.f90:1208-1453
//   diagonal_preconditioner (with its bug)    linalg.f90:1798-1831
// Parity pinning: GMRES, block_jacobi_solve and block_ssor_solve (and through the block solvers decompose_blocks,
// lu_decomp, lu_back_sub, the DIAG scale and the N/5 block size) reproduce the reference's stored iteration histories
// studies/matrix_solvers/iterations/*.csv row by row (tests/test_oracle_solver_histories.py); GMRES / BJAC / FQRUP also run
// inside the golden tuples of test/test_machline.py (tests/test_oracle_golden.py).  restarted_GMRES and purcell_solve
// have no reference data: parity unpinned for those two.
// Matrices are column-major, A(i,j) = A[i + j*N], as in the reference.  Dense matvecs follow the
// natural ascending-k order (libgfortran's matmul may use FMA variants at run time, so the
// reference itself is only reproducible to rounding there).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include <omp.h>

#include "oracle.h"

namespace {

inline double fsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }

double norm2_gf(const double* x, int n) {  // gfortran NORM2
    double result = 0, scale = 1;
    for (int i = 0; i < n; ++i) {
        if (x[i] != 0) {
            double ax = std::fabs(x[i]);
            if (scale < ax) {
                double val = scale / ax;
                result = 1 + result * val * val;
                scale = ax;
            } else {
                double val = ax / scale;
                result += val * val;
            }
        }
    }
    return scale * std::sqrt(result);
}

// y = A(r0:r1, c0:c1) x(c0:c1) accumulating over columns in ascending order (axpy form)
void matvec_block(int N, const double* A, int r0, int r1, int c0, int c1, const double* x, double* y) {
    for (int i = r0; i < r1; ++i) y[i - r0] = 0.;
    for (int j = c0; j < c1; ++j) {
        const double* col = A + (size_t)j * N;
        double xj = x[j];
        for (int i = r0; i < r1; ++i) y[i - r0] += col[i] * xj;
    }
}
void matvec_full(int N, const double* A, const double* x, double* y) {
    int nt = omp_get_max_threads();
    if ((size_t)N * N < (size_t)1 << 16) nt = 1;
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int t = 0; t < nt; ++t) {
        int r0 = (int)((long long)N * t / nt), r1 = (int)((long long)N * (t + 1) / nt);
        if (r1 > r0) matvec_block(N, A, r0, r1, 0, N, x, y + r0);
    }
}

// linalg.f90:166-280; A is n x n with leading dimension ld
int lu_decomp(double* A, int ld, int n, int* indx) {
    std::vector<double> vv(n);
    const double tiny = 1.5e-20;
    int imax = 0;
    for (int i = 0; i < n; ++i) {
        double amax = 0.0;
        for (int j = 0; j < n; ++j)
            if (std::fabs(A[i + (size_t)j * ld]) > amax) amax = std::fabs(A[i + (size_t)j * ld]);
        if (amax <= tiny) return 1;
        vv[i] = 1.0 / amax;
    }
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < j; ++i) {
            double s = A[i + (size_t)j * ld];
            for (int k = 0; k < i; ++k) s = s - A[i + (size_t)k * ld] * A[k + (size_t)j * ld];
            A[i + (size_t)j * ld] = s;
        }
        double amax = 0.0;
        for (int i = j; i < n; ++i) {
            double s = A[i + (size_t)j * ld];
            for (int k = 0; k < j; ++k) s = s - A[i + (size_t)k * ld] * A[k + (size_t)j * ld];
            A[i + (size_t)j * ld] = s;
            double dum = vv[i] * std::fabs(s);
            if (dum >= amax) {
                imax = i;
                amax = dum;
            }
        }
        if (j != imax) {
            for (int k = 0; k < n; ++k) std::swap(A[imax + (size_t)k * ld], A[j + (size_t)k * ld]);
            vv[imax] = vv[j];
        }
        indx[j] = imax;
        if (j != n - 1) {
            double dum = 1.0 / A[j + (size_t)j * ld];
            for (int i = j + 1; i < n; ++i) A[i + (size_t)j * ld] = A[i + (size_t)j * ld] * dum;
        }
    }
    return 0;
}

// linalg.f90:283-342
void lu_back_sub(const double* A, int ld, int n, const int* indx, const double* b, double* x) {
    for (int i = 0; i < n; ++i) x[i] = b[i];
    int ii = -1;
    for (int i = 0; i < n; ++i) {
        int ll = indx[i];
        double sum = x[ll];
        x[ll] = x[i];
        if (ii != -1) {
            for (int j = ii; j < i; ++j) sum = sum - A[i + (size_t)j * ld] * x[j];
        } else if (sum != 0.0) {
            ii = i;
        }
        x[i] = sum;
    }
    for (int i = n - 1; i >= 0; --i) {
        double sum = x[i];
        for (int j = i + 1; j < n; ++j) sum = sum - A[i + (size_t)j * ld] * x[j];
        x[i] = sum / A[i + (size_t)i * ld];
    }
}

struct Blocks {
    int N_blocks = 0, block_size = 0, N_last = 0;
    std::vector<std::vector<double>> A_blocks;
    std::vector<std::vector<int>> ind_P;
    std::vector<int> i_start, i_end;  // [start, end) 0-based
};

// linalg.f90:376-456
int decompose_blocks(int N, const double* A, int block_size, Blocks& B) {
    B.block_size = block_size;
    B.N_blocks = N / block_size;
    if (N % block_size > 0) B.N_blocks += 1;
    B.A_blocks.resize(B.N_blocks);
    B.ind_P.resize(B.N_blocks);
    B.i_start.resize(B.N_blocks);
    B.i_end.resize(B.N_blocks);
    int bad = 0;
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < B.N_blocks; ++i) {
        int s = i * block_size;
        int e = (i == B.N_blocks - 1) ? N : (i + 1) * block_size;
        int n = e - s;
        B.i_start[i] = s;
        B.i_end[i] = e;
        if (i == B.N_blocks - 1) B.N_last = n;
        B.A_blocks[i].resize((size_t)n * n);
        B.ind_P[i].resize(n);
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < n; ++r) B.A_blocks[i][r + (size_t)c * n] = A[(s + r) + (size_t)(s + c) * N];
        if (lu_decomp(B.A_blocks[i].data(), n, n, B.ind_P[i].data())) {
#pragma omp critical
            bad = 1;
        }
    }
    return bad;
}

// linalg.f90:930-965; R is k x k with leading dimension ld
int upper_triangular_back_sub(int n, const double* R, int ld, const double* b, double* x) {
    for (int i = n - 1; i >= 0; --i) {
        x[i] = b[i];
        for (int j = i + 1; j < n; ++j) x[i] = x[i] - R[i + (size_t)j * ld] * x[j];
        if (R[i + (size_t)i * ld] != 0.) x[i] = x[i] / R[i + (size_t)i * ld];
        else return ML_SINGULAR;
    }
    return 0;
}

// linalg.f90:1208-1232 (Q is N x (k_max), H is (k_max+1) x k_max, column-major); k is 0-based here
void arnoldi_update(int N, const double* A, int k, double* Q, double* H, int ldh) {
    double* w = Q + (size_t)(k + 1) * N;
    matvec_full(N, A, Q + (size_t)k * N, w);
    for (int i = 0; i <= k; ++i) {
        const double* qi = Q + (size_t)i * N;
        double h = 0.;
        for (int r = 0; r < N; ++r) h += w[r] * qi[r];
        H[i + (size_t)k * ldh] = h;
        for (int r = 0; r < N; ++r) w[r] = w[r] - h * qi[r];
    }
    double nrm = norm2_gf(w, N);
    H[(k + 1) + (size_t)k * ldh] = nrm;
    for (int r = 0; r < N; ++r) w[r] = w[r] / nrm;
}

}  // namespace

// bench.py --impl reference under torchrun: the launcher exports OMP_NUM_THREADS=1, the reference run uses every core
extern "C" void orc_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

extern "C" int orc_lower_bandwidth(int N, const double* A) {  // linalg.f90:797-835
    int B_l = 0;
    for (int i = N - 1; i >= 0; --i) {
        bool found = false;
        int j = 0;
        for (j = 0; j < i; ++j) {
            found = std::fabs(A[i + (size_t)j * N]) > 1.0e-12;
            if (found) break;
        }
        if (found) B_l = std::max(B_l, i - j);
    }
    return B_l;
}

// lu_decomp alone (linalg.f90:166-280): A overwritten by its factors, indx = the 0-based pivot row of every column
extern "C" int orc_lu_decomp(int N, double* A, int* indx) { return lu_decomp(A, N, N, indx); }

extern "C" int orc_lu_solve(int N, double* A, const double* b, double* x) {  // linalg.f90:118-148
    std::vector<int> indx(N);
    if (lu_decomp(A, N, N, indx.data())) return ML_SINGULAR;
    lu_back_sub(A, N, N, indx.data(), b, x);
    return 0;
}

// linalg.f90:1235-1334
extern "C" int orc_gmres(int N, const double* A, const double* b, double tol, int max_iter, int* total_iter, double* x,
                         double* err_history) {
    int k_max = std::min(N, max_iter);
    std::vector<double> Q((size_t)N * (k_max + 1), 0.);  // one spare column keeps k+1 <= k_max-1 safe
    int ldh = k_max + 1;
    std::vector<double> H((size_t)ldh * k_max, 0.), c(k_max, 0.), s(k_max, 0.), E(N > k_max + 1 ? N : k_max + 1, 0.);
    E[0] = 1.;
    double beta = norm2_gf(b, N);
    for (int r = 0; r < N; ++r) Q[r] = b[r] / beta;
    int k = 0;  // 1-based count as in the reference
    double err = tol + 1;
    while (err > tol && k < k_max - 1) {
        k = k + 1;
        int kk = k - 1;  // 0-based column
        arnoldi_update(N, A, kk, Q.data(), H.data(), ldh);
        for (int i = 0; i < kk; ++i) {
            double temp = c[i] * H[i + (size_t)kk * ldh] + s[i] * H[(i + 1) + (size_t)kk * ldh];
            H[(i + 1) + (size_t)kk * ldh] = -s[i] * H[i + (size_t)kk * ldh] + c[i] * H[(i + 1) + (size_t)kk * ldh];
            H[i + (size_t)kk * ldh] = temp;
        }
        double hkk = H[kk + (size_t)kk * ldh], hk1 = H[(kk + 1) + (size_t)kk * ldh];
        double d = std::sqrt(hkk * hkk + hk1 * hk1);
        c[kk] = std::fabs(hkk) / d;
        s[kk] = fsign(1., hkk) * hk1 / d;
        H[kk + (size_t)kk * ldh] = c[kk] * hkk + s[kk] * hk1;
        H[(kk + 1) + (size_t)kk * ldh] = 0.;
        E[kk + 1] = -s[kk] * E[kk];
        E[kk] = c[kk] * E[kk];
        err = beta * std::fabs(E[kk + 1]);
        if (err_history) err_history[kk] = err;
        if (err < tol) break;
    }
    std::vector<double> y(std::max(k, 1)), rhs(std::max(k, 1));
    for (int i = 0; i < k; ++i) rhs[i] = beta * E[i];
    int st = upper_triangular_back_sub(k, H.data(), ldh, rhs.data(), y.data());
    if (st) return st;
    for (int r = 0; r < N; ++r) x[r] = 0.;
    for (int j = 0; j < k; ++j) {
        const double* qj = Q.data() + (size_t)j * N;
        for (int r = 0; r < N; ++r) x[r] += qj[r] * y[j];
    }
    *total_iter = k;
    return 0;
}

// linalg.f90:1337-1453
extern "C" int orc_restarted_gmres(int N, const double* A, const double* b, double tol, int max_iter, int restart_iter,
                                   int* total_iter_out, double* x) {
    int total_iter = 0;
    int k_max = std::min(restart_iter, N);
    int ldh = k_max + 1;
    std::vector<double> Q((size_t)N * (k_max + 1)), H((size_t)ldh * k_max), c(k_max), s(k_max), E(std::max(N, k_max + 1)), r0(N), Ax(N);
    for (int r = 0; r < N; ++r) x[r] = 0.;
    double err = tol + 1;
    while (err > tol && total_iter <= max_iter) {
        std::fill(Q.begin(), Q.end(), 0.);
        std::fill(H.begin(), H.end(), 0.);
        std::fill(c.begin(), c.end(), 0.);
        std::fill(s.begin(), s.end(), 0.);
        std::fill(E.begin(), E.end(), 0.);
        E[0] = 1.;
        matvec_full(N, A, x, Ax.data());
        for (int r = 0; r < N; ++r) r0[r] = b[r] - Ax[r];
        double beta = norm2_gf(r0.data(), N);
        for (int r = 0; r < N; ++r) Q[r] = r0[r] / beta;
        int k = 0;
        while (err > tol && k < k_max - 1) {
            k = k + 1;
            total_iter = total_iter + 1;
            int kk = k - 1;
            arnoldi_update(N, A, kk, Q.data(), H.data(), ldh);
            for (int i = 0; i < kk; ++i) {
                double temp = c[i] * H[i + (size_t)kk * ldh] + s[i] * H[(i + 1) + (size_t)kk * ldh];
                H[(i + 1) + (size_t)kk * ldh] = -s[i] * H[i + (size_t)kk * ldh] + c[i] * H[(i + 1) + (size_t)kk * ldh];
                H[i + (size_t)kk * ldh] = temp;
            }
            double hkk = H[kk + (size_t)kk * ldh], hk1 = H[(kk + 1) + (size_t)kk * ldh];
            double d = std::sqrt(hkk * hkk + hk1 * hk1);
            c[kk] = std::fabs(hkk) / d;
            s[kk] = fsign(1., hkk) * hk1 / d;
            H[kk + (size_t)kk * ldh] = c[kk] * hkk + s[kk] * hk1;
            H[(kk + 1) + (size_t)kk * ldh] = 0.;
            E[kk + 1] = -s[kk] * E[kk];
            E[kk] = c[kk] * E[kk];
            err = beta * std::fabs(E[kk + 1]);
        }
        if (k == 0) break;  // the reference would loop forever / read H(1:0): guard
        std::vector<double> y(k), rhs(k);
        for (int i = 0; i < k; ++i) rhs[i] = beta * E[i];
        int st = upper_triangular_back_sub(k, H.data(), ldh, rhs.data(), y.data());
        if (st) return st;
        std::vector<double> dx(N, 0.);
        for (int j = 0; j < k; ++j) {
            const double* qj = Q.data() + (size_t)j * N;
            for (int r = 0; r < N; ++r) dx[r] += qj[r] * y[j];
        }
        for (int r = 0; r < N; ++r) x[r] = x[r] + dx[r];
    }
    *total_iter_out = total_iter;
    return 0;
}

// Optional per-iteration history of the block solvers (the columns ||dx||, ||err|| of the reference's iteration file,
// linalg.f90:714-717, 585-587): orc_set_block_history(dx, err, capacity) arms it for the next block solve.
static double* g_hist_dx = nullptr;
static double* g_hist_err = nullptr;
static int g_hist_cap = 0;
extern "C" void orc_set_block_history(double* dx, double* err, int capacity) {
    g_hist_dx = dx;
    g_hist_err = err;
    g_hist_cap = capacity;
}
static void record_block_history(int iteration, const double* x, const double* x_new, int N, double err) {
    if (!g_hist_err || iteration > g_hist_cap) return;
    std::vector<double> d(N);
    for (int i = 0; i < N; ++i) d[i] = x[i] - x_new[i];
    if (g_hist_dx) g_hist_dx[iteration - 1] = norm2_gf(d.data(), N);
    g_hist_err[iteration - 1] = err;
}

// linalg.f90:601-728
extern "C" int orc_block_jacobi(int N, double* A, const double* b, int block_size, double tol, double rel, int max_iter,
                                int* total_iter, double* x) {
    double err = tol + 1.;
    for (int i = 0; i < N; ++i) x[i] = b[i] / A[i + (size_t)i * N];
    Blocks B;
    if (decompose_blocks(N, A, block_size, B)) return ML_SINGULAR;
    std::vector<double> x_new(N), vk(N);
    int iteration = 0;
    while (err >= tol && iteration < max_iter) {
        iteration += 1;
#pragma omp parallel for schedule(dynamic)
        for (int i = 0; i < B.N_blocks; ++i) {
            int s = B.i_start[i], e = B.i_end[i], n = e - s;
            std::vector<double> bi(n), t(n), xi(n);
            for (int r = 0; r < n; ++r) bi[r] = b[s + r];
            if (s > 0) {
                matvec_block(N, A, s, e, 0, s, x, t.data());
                for (int r = 0; r < n; ++r) bi[r] = bi[r] - t[r];
            } else if (i == B.N_blocks - 1) {
                // last block with an empty left part: b - matmul(empty) = b
            }
            if (i != B.N_blocks - 1) {
                matvec_block(N, A, s, e, e, N, x, t.data());
                for (int r = 0; r < n; ++r) bi[r] = bi[r] - t[r];
            }
            lu_back_sub(B.A_blocks[i].data(), n, n, B.ind_P[i].data(), bi.data(), xi.data());
            for (int r = 0; r < n; ++r) x_new[s + r] = xi[r];
        }
        for (int i = 0; i < N; ++i) x_new[i] = (1. - rel) * x[i] + rel * x_new[i];
        matvec_full(N, A, x_new.data(), vk.data());
        std::vector<double> dvec(N), rvec(N);
        for (int i = 0; i < N; ++i) {
            rvec[i] = vk[i] - b[i];
            dvec[i] = x[i] - x_new[i];
        }
        err = norm2_gf(rvec.data(), N);
        record_block_history(iteration, x, x_new.data(), N, err);
        for (int i = 0; i < N; ++i) x[i] = x_new[i];
    }
    *total_iter = iteration;
    return 0;
}

// linalg.f90:459-598
extern "C" int orc_block_ssor(int N, double* A, const double* b, int block_size, double tol, double rel, int max_iter,
                              int* total_iter, double* x) {
    double err = tol + 1.;
    int step = -1;
    for (int i = 0; i < N; ++i) x[i] = 0.;
    Blocks B;
    if (decompose_blocks(N, A, block_size, B)) return ML_SINGULAR;
    std::vector<double> x_new(N, 0.), vk(N);
    int iteration = 0;
    while (err >= tol && iteration < max_iter) {
        iteration += 1;
        int start, end;
        if (step == 1) {
            start = B.N_blocks - 1;
            end = 0;
            step = -1;
        } else {
            start = 0;
            end = B.N_blocks - 1;
            step = 1;
        }
        for (int i = start; step == 1 ? i <= end : i >= end; i += step) {
            int s = B.i_start[i], e = B.i_end[i], n = e - s;
            std::vector<double> bi(n), t(n), xi(n);
            for (int r = 0; r < n; ++r) bi[r] = b[s + r];
            const double* left = (step == 1) ? x_new.data() : x;
            const double* right = (step == 1) ? x : x_new.data();
            if (s > 0) {
                matvec_block(N, A, s, e, 0, s, left, t.data());
                for (int r = 0; r < n; ++r) bi[r] = bi[r] - t[r];
            }
            if (i != B.N_blocks - 1) {
                matvec_block(N, A, s, e, e, N, right, t.data());
                for (int r = 0; r < n; ++r) bi[r] = bi[r] - t[r];
            }
            lu_back_sub(B.A_blocks[i].data(), n, n, B.ind_P[i].data(), bi.data(), xi.data());
            for (int r = 0; r < n; ++r) x_new[s + r] = (1. - rel) * x[s + r] + rel * xi[r];
        }
        matvec_full(N, A, x_new.data(), vk.data());
        std::vector<double> rvec(N);
        for (int i = 0; i < N; ++i) rvec[i] = vk[i] - b[i];
        err = norm2_gf(rvec.data(), N);
        record_block_history(iteration, x, x_new.data(), N, err);
        for (int i = 0; i < N; ++i) x[i] = x_new[i];
    }
    *total_iter = iteration;
    return 0;
}

// linalg.f90:838-927
extern "C" int orc_qr_givens_up(int N, double* A, double* b, double* x) {
    int B_l = orc_lower_bandwidth(N, A);
    for (int j = 0; j < N; ++j) {
        for (int i = std::min(j + B_l, N - 1); i >= j + 1; --i) {
            double& xx = A[j + (size_t)j * N];
            double& yy = A[i + (size_t)j * N];
            if (yy != 0.) {
                double t = std::fabs(xx) + std::fabs(yy);
                double d = t * std::sqrt((xx / t) * (xx / t) + (yy / t) * (yy / t));
                double c = xx / d, s = yy / d;
                xx = d;
                yy = 0.;
                // apply to A(j, j+1:) and A(i, j+1:) -- length N-j-1 in the reference (the last
                // column is never rotated: linalg.f90:914 passes N-j-1 for a slice of N-j)
                for (int k = j + 1; k < j + 1 + (N - (j + 1) - 1); ++k) {
                    double a1 = A[j + (size_t)k * N], a2 = A[i + (size_t)k * N];
                    A[j + (size_t)k * N] = c * a1 + s * a2;
                    A[i + (size_t)k * N] = c * a2 - s * a1;
                }
                double b1 = b[j], b2 = b[i];
                b[j] = c * b1 + s * b2;
                b[i] = c * b2 - s * b1;
            }
        }
    }
    return upper_triangular_back_sub(N, A, N, b, x);
}

// linalg.f90:968-1165
extern "C" int orc_qr_fast_givens_up(int N, double* A, double* b, double* x) {
    std::vector<double> D(N, 1.);
    int B_l = orc_lower_bandwidth(N, A);
    for (int j = 0; j < N; ++j) {
        for (int i = std::min(j + B_l, N - 1); i >= j + 1; --i) {
            double& xx = A[j + (size_t)j * N];
            double& yy = A[i + (size_t)j * N];
            if (yy != 0.) {
                double &D1 = D[j], &D2 = D[i];
                double gamma = D1 / D2, ratio, a, bb, t, d;
                int rot_type;
                if (xx != 0.) ratio = (yy * yy) / (xx * xx);
                else ratio = gamma + 1.;
                if (D1 >= D2) {
                    if (ratio <= gamma) {
                        rot_type = 1;
                        t = yy / xx;
                        bb = t / gamma;
                        d = 1. + bb * t;
                        a = t / d;
                        D1 = D1 / d;
                        D2 = D2 * d;
                        xx = xx * d;
                    } else {
                        rot_type = 3;
                        a = xx / yy;
                        t = a * gamma;
                        d = 1. + a * t;
                        bb = t / d;
                        double temp = D2 * d;
                        D2 = D1 / d;
                        D1 = temp;
                        xx = yy;
                    }
                } else {
                    if (ratio <= gamma) {
                        rot_type = 2;
                        a = yy / xx;
                        t = a / gamma;
                        d = 1. + a * t;
                        bb = t / d;
                        D1 = D1 * d;
                        D2 = D2 / d;
                    } else {
                        rot_type = 4;
                        t = xx / yy;
                        bb = t * gamma;
                        d = 1. + bb * t;
                        a = t / d;
                        double temp = D2 / d;
                        D2 = D1 * d;
                        D1 = temp;
                        xx = yy * d;
                    }
                }
                yy = 0.;
                auto apply = [&](double& px, double& py) {  // linalg.f90:1074-1112
                    double temp;
                    switch (rot_type) {
                        case 1: px = px + bb * py; py = py - a * px; break;
                        case 2: py = py - a * px; px = px + bb * py; break;
                        case 3: temp = py; py = a * py - px; px = temp - bb * py; break;
                        case 4: temp = px; px = bb * px + py; py = a * px - temp; break;
                    }
                };
                for (int k = j + 1; k < j + 1 + (N - (j + 1) - 1); ++k) apply(A[j + (size_t)k * N], A[i + (size_t)k * N]);
                apply(b[j], b[i]);
            }
        }
    }
    return upper_triangular_back_sub(N, A, N, b, x);
}

// linalg.f90:731-794
extern "C" int orc_purcell(int N, const double* A, const double* b, double* x) {
    int M = N + 1;
    std::vector<double> V((size_t)M * M, 0.), V_s(M), d(M);
    std::vector<int> m(N);
    for (int i = 0; i < M; ++i) V[i + (size_t)i * M] = 1.;
    for (int i = N; i >= 1; --i) {
        int row = N - i;  // A(N+1-i,:) 0-based
        for (int k = 0; k <= i; ++k) {
            double sum = 0.;
            for (int c = 0; c < N; ++c) sum += A[row + (size_t)c * N] * V[c + (size_t)k * M];
            d[k] = sum - b[row] * V[N + (size_t)k * M];
        }
        int s = 0;
        for (int k = 1; k <= i; ++k)
            if (std::fabs(d[k]) > std::fabs(d[s])) s = k;
        for (int r = 0; r < M; ++r) V_s[r] = V[r + (size_t)s * M];
        for (int k = 0; k < i; ++k) m[k] = (k < s) ? k : k + 1;
        double denom = 1. / d[s];
        for (int k = 0; k < i; ++k) {
            double alpha = -d[m[k]] * denom;
            for (int r = 0; r < M; ++r) V[r + (size_t)k * M] = alpha * V_s[r] + V[r + (size_t)m[k] * M];
        }
    }
    for (int r = 0; r < N; ++r) x[r] = V[r] / V[N];
    return 0;
}

// panel_solver.f90:1802-2027 (square systems)
extern "C" int orc_solve_system(int N, const double* A, const double* I_known, const double* BC, const ml_solver_opts* opts,
                                double* x, ml_solve_info* info) {
    std::vector<double> b(N), A_p((size_t)N * N), b_p(N);
    for (int i = 0; i < N; ++i) b[i] = BC[i] - (I_known ? I_known[i] : 0.);
    if (opts->preconditioner == ML_PREC_DIAG) {
        // linalg.f90:1813-1816: every entry of A_ii_inv ends up as 1/A(N,N)
        double inv = 1. / A[(N - 1) + (size_t)(N - 1) * N];
        for (size_t k = 0; k < (size_t)N * N; ++k) A_p[k] = inv * A[k];
        for (int i = 0; i < N; ++i) b_p[i] = b[i] * inv;
    } else {
        std::memcpy(A_p.data(), A, sizeof(double) * (size_t)N * N);
        b_p = b;
    }
    int block_size = opts->block_size;
    if (block_size <= 0) block_size = N / 5;
    int iters = -1, st = 0;
    switch (opts->matrix_solver) {
        case ML_SOLVER_LU: st = orc_lu_solve(N, A_p.data(), b_p.data(), x); break;
        case ML_SOLVER_QRUP: st = orc_qr_givens_up(N, A_p.data(), b_p.data(), x); break;
        case ML_SOLVER_FQRUP: st = orc_qr_fast_givens_up(N, A_p.data(), b_p.data(), x); break;
        case ML_SOLVER_RGMRES:
            st = orc_restarted_gmres(N, A_p.data(), b_p.data(), opts->tol, opts->max_iterations, opts->restart_iterations, &iters, x);
            break;
        case ML_SOLVER_PURC: st = orc_purcell(N, A_p.data(), b_p.data(), x); break;
        case ML_SOLVER_BSSOR:
            st = orc_block_ssor(N, A_p.data(), b_p.data(), block_size, opts->tol, opts->rel, opts->max_iterations, &iters, x);
            break;
        case ML_SOLVER_BJAC:
            st = orc_block_jacobi(N, A_p.data(), b_p.data(), block_size, opts->tol, opts->rel, opts->max_iterations, &iters, x);
            break;
        case ML_SOLVER_GMRES:
        default: st = orc_gmres(N, A_p.data(), b_p.data(), opts->tol, opts->max_iterations, &iters, x, nullptr); break;
    }
    if (st) return st;
    std::vector<double> R(N);
    matvec_full(N, A, x, R.data());
    double mx = 0., ss = 0.;
    for (int i = 0; i < N; ++i) {
        R[i] = R[i] - b[i];
        mx = std::max(mx, std::fabs(R[i]));
        ss += R[i] * R[i];
    }
    if (info) {
        info->iterations = iters;
        info->res_max = mx;
        info->res_norm = std::sqrt(ss);
    }
    if (std::isnan(std::sqrt(ss))) return ML_NAN_RESIDUAL;
    return 0;
}

// panel_solver.f90:1842-1895, overdetermined least squares: A_p = diagonal_preconditioner(matmul(transpose(A), A)),
// b_p likewise from matmul(transpose(A), b); the residual is A x - b of the original system (:1992-1998)
extern "C" int orc_solve_system_ls(int n_cp, int N, const double* A, const double* I_known, const double* BC, const ml_solver_opts* opts,
                                   double* x, ml_solve_info* info) {
    std::vector<double> b(n_cp), AtA((size_t)N * N), Atb(N);
    for (int i = 0; i < n_cp; ++i) b[i] = BC[i] - (I_known ? I_known[i] : 0.);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < N; ++j) {
        const double* aj = A + (size_t)j * n_cp;
        for (int i = 0; i < N; ++i) {
            const double* ai = A + (size_t)i * n_cp;
            double acc = 0.;
            for (int r = 0; r < n_cp; ++r) acc = acc + ai[r] * aj[r];
            AtA[i + (size_t)j * N] = acc;
        }
        double acc = 0.;
        for (int r = 0; r < n_cp; ++r) acc = acc + aj[r] * b[r];
        Atb[j] = acc;
    }
    ml_solve_info tmp{};
    std::vector<double> zero(N, 0.);
    int st = orc_solve_system(N, AtA.data(), zero.data(), Atb.data(), opts, x, &tmp);
    if (st && st != ML_NAN_RESIDUAL) return st;
    double mx = 0., ss = 0.;
    for (int r = 0; r < n_cp; ++r) {
        double acc = 0.;
        for (int j = 0; j < N; ++j) acc = acc + A[r + (size_t)j * n_cp] * x[j];
        const double R = acc - b[r];
        mx = std::max(mx, std::fabs(R));
        ss += R * R;
    }
    if (info) {
        info->iterations = tmp.iterations;
        info->res_max = mx;
        info->res_norm = std::sqrt(ss);
    }
    if (std::isnan(std::sqrt(ss))) return ML_NAN_RESIDUAL;
    return 0;
}
