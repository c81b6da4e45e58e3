// Tiny dense host-side linear algebra for the Krylov control flow (inner x inner problems,
// inner <= 32): Hermitian pseudo-inverse (FGMRES_relaxation, src/Multigrid/FGMRES.jl:99-104) and
// the Hessenberg least-squares problem of restarted (F)GMRES.  These are O(inner^3) scalar
// operations on values that are already on the host for the stopping tests.
#pragma once
#include <cmath>
#include <complex>
#include <limits>
#include <vector>

namespace mgb200 {

typedef std::complex<double> zc;

// Cyclic Jacobi eigen-decomposition of a real symmetric N x N matrix (row-major, overwritten);
// V receives the eigenvectors as columns.
static inline void jacobi_eig_sym(int N, std::vector<double>& A, std::vector<double>& V) {
    V.assign((size_t)N * N, 0.0);
    for (int i = 0; i < N; ++i) V[(size_t)i * N + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) (i == j ? diag : off) += A[(size_t)i * N + j] * A[(size_t)i * N + j];
        if (off <= 1e-60 || off <= 1e-32 * diag) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double apq = A[(size_t)p * N + q];
                if (apq == 0.0) continue;
                double app = A[(size_t)p * N + p], aqq = A[(size_t)q * N + q];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {
                    double akp = A[(size_t)k * N + p], akq = A[(size_t)k * N + q];
                    A[(size_t)k * N + p] = c * akp - s * akq;
                    A[(size_t)k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {
                    double apk = A[(size_t)p * N + k], aqk = A[(size_t)q * N + k];
                    A[(size_t)p * N + k] = c * apk - s * aqk;
                    A[(size_t)q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    double vkp = V[(size_t)k * N + p], vkq = V[(size_t)k * N + q];
                    V[(size_t)k * N + p] = c * vkp - s * vkq;
                    V[(size_t)k * N + q] = s * vkp + c * vkq;
                }
            }
    }
}

// t = pinv(H) * xi for Hermitian H (n x n, row-major).  Singular values <= eps*n*max are
// dropped, the rule of Julia's LinearAlgebra.pinv default (rtol = eps*min(size)).
// A complex Hermitian H = A + iB is embedded as the real symmetric [[A,-B],[B,A]].
// eps: machine epsilon of the hierarchy's value type (the reference's H is Float32 / ComplexF32 for single-precision
// hierarchies, so its pinv drops singular values below eps(Float32) * n * max).
static inline void hermitian_pinv_apply(int n, const std::vector<zc>& H, const std::vector<zc>& xi,
                                        std::vector<zc>& t, double eps = std::numeric_limits<double>::epsilon()) {
    const int N = 2 * n;
    std::vector<double> M((size_t)N * N), V;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double a = H[(size_t)i * n + j].real(), b = H[(size_t)i * n + j].imag();
            M[(size_t)i * N + j] = a;
            M[(size_t)i * N + (j + n)] = -b;
            M[(size_t)(i + n) * N + j] = b;
            M[(size_t)(i + n) * N + (j + n)] = a;
        }
    jacobi_eig_sym(N, M, V);
    double smax = 0.0;
    for (int i = 0; i < N; ++i) smax = std::max(smax, std::fabs(M[(size_t)i * N + i]));
    const double tol = eps * n * smax;
    std::vector<double> rhs(N), out(N, 0.0);
    for (int i = 0; i < n; ++i) {
        rhs[i] = xi[i].real();
        rhs[i + n] = xi[i].imag();
    }
    for (int e = 0; e < N; ++e) {
        double lam = M[(size_t)e * N + e];
        if (!(std::fabs(lam) > tol)) continue;
        double proj = 0.0;
        for (int k = 0; k < N; ++k) proj += V[(size_t)k * N + e] * rhs[k];
        proj /= lam;
        for (int k = 0; k < N; ++k) out[k] += V[(size_t)k * N + e] * proj;
    }
    t.resize(n);
    for (int i = 0; i < n; ++i) t[i] = zc(out[i], out[i + n]);
}

// Pseudo-inverse of a general n x n complex matrix (row-major) by one-sided Jacobi SVD on the real
// embedding [[A,-B],[B,A]]; singular values <= rtol * max are dropped (numpy / Julia pinv rule).
static inline void general_pinv(int n, const std::vector<zc>& Ain, double rtol, std::vector<zc>& Pout) {
    const int N = 2 * n;
    std::vector<double> W((size_t)N * N), V((size_t)N * N, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double a = Ain[(size_t)i * n + j].real(), b = Ain[(size_t)i * n + j].imag();
            W[(size_t)i * N + j] = a;
            W[(size_t)i * N + (j + n)] = -b;
            W[(size_t)(i + n) * N + j] = b;
            W[(size_t)(i + n) * N + (j + n)] = a;
        }
    for (int i = 0; i < N; ++i) V[(size_t)i * N + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double maxoff = 0.0;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < N; ++k) {
                    double wp = W[(size_t)k * N + p], wq = W[(size_t)k * N + q];
                    alpha += wp * wp;
                    beta += wq * wq;
                    gamma += wp * wq;
                }
                if (gamma == 0.0) continue;
                double off = std::fabs(gamma) / std::sqrt(alpha * beta);
                if (!(off > 1e-15)) continue;
                maxoff = std::max(maxoff, off);
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < N; ++k) {
                    double wp = W[(size_t)k * N + p], wq = W[(size_t)k * N + q];
                    W[(size_t)k * N + p] = c * wp - s * wq;
                    W[(size_t)k * N + q] = s * wp + c * wq;
                    double vp = V[(size_t)k * N + p], vq = V[(size_t)k * N + q];
                    V[(size_t)k * N + p] = c * vp - s * vq;
                    V[(size_t)k * N + q] = s * vp + c * vq;
                }
            }
        if (maxoff <= 1e-15) break;
    }
    std::vector<double> sig(N);
    double smax = 0.0;
    for (int j = 0; j < N; ++j) {
        double s2 = 0.0;
        for (int k = 0; k < N; ++k) s2 += W[(size_t)k * N + j] * W[(size_t)k * N + j];
        sig[j] = std::sqrt(s2);
        smax = std::max(smax, sig[j]);
    }
    // pinv(M) = V diag(1/sigma) U^T with U[:,j] = W[:,j]/sigma_j  =>  pinv(M)[r][c] = sum_j V[r][j] W[c][j] / sigma_j^2
    std::vector<double> Pm((size_t)N * N, 0.0);
    for (int j = 0; j < N; ++j) {
        if (!(sig[j] > rtol * smax)) continue;
        const double inv = 1.0 / (sig[j] * sig[j]);
        for (int r = 0; r < N; ++r) {
            const double vr = V[(size_t)r * N + j] * inv;
            if (vr == 0.0) continue;
            for (int c = 0; c < N; ++c) Pm[(size_t)r * N + c] += vr * W[(size_t)c * N + j];
        }
    }
    Pout.resize((size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Pout[(size_t)i * n + j] = zc(Pm[(size_t)i * N + j], Pm[(size_t)(i + n) * N + j]);
}

// min_y || H(0:rows,0:cols) y - xi(0:rows) ||, H row-major with leading dimension ldh,
// rows = cols+1 (Hessenberg block).  Householder QR on a copy; returns the residual norm.
static inline double hessenberg_lsq(const std::vector<zc>& H, int ldh, int rows, int cols,
                                    const std::vector<zc>& xi, std::vector<zc>& y) {
    std::vector<zc> A((size_t)rows * cols), b(rows);
    for (int i = 0; i < rows; ++i) {
        b[i] = xi[i];
        for (int j = 0; j < cols; ++j) A[(size_t)i * cols + j] = H[(size_t)i * ldh + j];
    }
    for (int k = 0; k < cols; ++k) {
        double nrm = 0.0;
        for (int i = k; i < rows; ++i) nrm += std::norm(A[(size_t)i * cols + k]);
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) continue;
        zc akk = A[(size_t)k * cols + k];
        zc phase = (std::abs(akk) == 0.0) ? zc(1.0, 0.0) : akk / std::abs(akk);
        zc alpha = -phase * nrm;
        std::vector<zc> v(rows - k);
        for (int i = k; i < rows; ++i) v[i - k] = A[(size_t)i * cols + k];
        v[0] -= alpha;
        double vn = 0.0;
        for (auto& e : v) vn += std::norm(e);
        if (vn == 0.0) continue;
        for (int j = k; j < cols; ++j) {
            zc s = 0.0;
            for (int i = k; i < rows; ++i) s += std::conj(v[i - k]) * A[(size_t)i * cols + j];
            s *= 2.0 / vn;
            for (int i = k; i < rows; ++i) A[(size_t)i * cols + j] -= s * v[i - k];
        }
        zc s = 0.0;
        for (int i = k; i < rows; ++i) s += std::conj(v[i - k]) * b[i];
        s *= 2.0 / vn;
        for (int i = k; i < rows; ++i) b[i] -= s * v[i - k];
    }
    y.assign(cols, zc(0.0, 0.0));
    for (int k = cols - 1; k >= 0; --k) {
        zc s = b[k];
        for (int j = k + 1; j < cols; ++j) s -= A[(size_t)k * cols + j] * y[j];
        zc akk = A[(size_t)k * cols + k];
        y[k] = (std::abs(akk) == 0.0) ? zc(0.0, 0.0) : s / akk;
    }
    double res = 0.0;
    for (int i = cols; i < rows; ++i) res += std::norm(b[i]);
    return std::sqrt(res);
}

// min_Y || A Y - B ||_F for a dense rows x cols matrix A (row-major) and nb right-hand sides B (rows x nb,
// row-major): Householder QR on copies (block GMRES: the block-Hessenberg least-squares problem).  Columns whose
// R-diagonal vanishes get a zero solution component.  Returns the Frobenius norm of the residual.
static inline double dense_lsq(int rows, int cols, int nb, const std::vector<zc>& Ain, const std::vector<zc>& Bin,
                               std::vector<zc>& Y) {
    std::vector<zc> A(Ain), B(Bin);
    std::vector<zc> v(rows);
    for (int k = 0; k < cols && k < rows; ++k) {
        double nrm = 0.0;
        for (int i = k; i < rows; ++i) nrm += std::norm(A[(size_t)i * cols + k]);
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) continue;
        const zc akk = A[(size_t)k * cols + k];
        const zc phase = (std::abs(akk) == 0.0) ? zc(1.0, 0.0) : akk / std::abs(akk);
        const zc alpha = -phase * nrm;
        for (int i = k; i < rows; ++i) v[i] = A[(size_t)i * cols + k];
        v[k] -= alpha;
        double vn = 0.0;
        for (int i = k; i < rows; ++i) vn += std::norm(v[i]);
        if (vn == 0.0) continue;
        for (int j = k; j < cols; ++j) {
            zc s = 0.0;
            for (int i = k; i < rows; ++i) s += std::conj(v[i]) * A[(size_t)i * cols + j];
            s *= 2.0 / vn;
            for (int i = k; i < rows; ++i) A[(size_t)i * cols + j] -= s * v[i];
        }
        for (int j = 0; j < nb; ++j) {
            zc s = 0.0;
            for (int i = k; i < rows; ++i) s += std::conj(v[i]) * B[(size_t)i * nb + j];
            s *= 2.0 / vn;
            for (int i = k; i < rows; ++i) B[(size_t)i * nb + j] -= s * v[i];
        }
    }
    Y.assign((size_t)cols * nb, zc(0.0, 0.0));
    for (int j = 0; j < nb; ++j)
        for (int k = cols - 1; k >= 0; --k) {
            zc s = B[(size_t)k * nb + j];
            for (int c = k + 1; c < cols; ++c) s -= A[(size_t)k * cols + c] * Y[(size_t)c * nb + j];
            const zc akk = A[(size_t)k * cols + k];
            Y[(size_t)k * nb + j] = (std::abs(akk) == 0.0) ? zc(0.0, 0.0) : s / akk;
        }
    double res = 0.0;
    for (int i = cols; i < rows; ++i)
        for (int j = 0; j < nb; ++j) res += std::norm(B[(size_t)i * nb + j]);
    return std::sqrt(res);
}

// Upper Cholesky factor of a Hermitian positive definite m x m matrix (row-major): G = R^H R with a real positive
// diagonal.  Returns false when a pivot is not positive (G numerically singular).
static inline bool cholesky_upper(int m, const std::vector<zc>& G, std::vector<zc>& R) {
    R.assign((size_t)m * m, zc(0.0, 0.0));
    for (int j = 0; j < m; ++j) {
        double d = G[(size_t)j * m + j].real();
        for (int k = 0; k < j; ++k) d -= std::norm(R[(size_t)k * m + j]);
        if (!(d > 0.0)) return false;
        const double rjj = std::sqrt(d);
        R[(size_t)j * m + j] = rjj;
        for (int i = j + 1; i < m; ++i) {
            zc s = G[(size_t)j * m + i];
            for (int k = 0; k < j; ++k) s -= std::conj(R[(size_t)k * m + j]) * R[(size_t)k * m + i];
            R[(size_t)j * m + i] = s / rjj;
        }
    }
    return true;
}

// inverse of an upper triangular m x m matrix (row-major)
static inline void upper_inverse(int m, const std::vector<zc>& R, std::vector<zc>& Rinv) {
    Rinv.assign((size_t)m * m, zc(0.0, 0.0));
    for (int j = 0; j < m; ++j) {
        Rinv[(size_t)j * m + j] = 1.0 / R[(size_t)j * m + j];
        for (int i = j - 1; i >= 0; --i) {
            zc s = 0.0;
            for (int k = i + 1; k <= j; ++k) s += R[(size_t)i * m + k] * Rinv[(size_t)k * m + j];
            Rinv[(size_t)i * m + j] = -s / R[(size_t)i * m + i];
        }
    }
}

// X = A^{-1} B for an m x m matrix A and m x nb right-hand sides (row-major), LU with partial pivoting.
// Returns false when A is singular to working precision (pivot <= eps * m * max|A|).
static inline bool lu_solve_small(int m, int nb, const std::vector<zc>& Ain, const std::vector<zc>& Bin,
                                  std::vector<zc>& X) {
    std::vector<zc> A(Ain);
    X = Bin;
    double amax = 0.0;
    for (const zc& a : A) amax = std::max(amax, std::abs(a));
    const double tiny = 2.220446049250313e-16 * m * amax;
    for (int k = 0; k < m; ++k) {
        int piv = k;
        double best = std::abs(A[(size_t)k * m + k]);
        for (int i = k + 1; i < m; ++i)
            if (std::abs(A[(size_t)i * m + k]) > best) {
                best = std::abs(A[(size_t)i * m + k]);
                piv = i;
            }
        if (!(best > tiny)) return false;
        if (piv != k) {
            for (int j = 0; j < m; ++j) std::swap(A[(size_t)k * m + j], A[(size_t)piv * m + j]);
            for (int j = 0; j < nb; ++j) std::swap(X[(size_t)k * nb + j], X[(size_t)piv * nb + j]);
        }
        for (int i = k + 1; i < m; ++i) {
            const zc f = A[(size_t)i * m + k] / A[(size_t)k * m + k];
            if (f == zc(0.0, 0.0)) continue;
            for (int j = k + 1; j < m; ++j) A[(size_t)i * m + j] -= f * A[(size_t)k * m + j];
            for (int j = 0; j < nb; ++j) X[(size_t)i * nb + j] -= f * X[(size_t)k * nb + j];
        }
    }
    for (int j = 0; j < nb; ++j)
        for (int k = m - 1; k >= 0; --k) {
            zc s = X[(size_t)k * nb + j];
            for (int c = k + 1; c < m; ++c) s -= A[(size_t)k * m + c] * X[(size_t)c * nb + j];
            X[(size_t)k * nb + j] = s / A[(size_t)k * m + k];
        }
    return true;
}

}  // namespace mgb200
