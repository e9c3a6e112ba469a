// b2a_eigen_stub.h -- the small part of Eigen's dense API that /root/reference/src/aruco_slam.cpp uses,
// so that the UNMODIFIED reference source compiles in an image that has no Eigen (oracle/_ref build).
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  This is not Eigen: every expression evaluates eagerly
// into a plain row-major double array with textbook loops (products sum k = 0..K-1 in order), so results
// differ from an Eigen build by rounding order only (~1e-16 relative); control flow, formulas and the
// order of the matrix expressions are the reference's own, because its source is compiled as it is.
// Covered surface: Matrix<double,R,C> / MatrixXd / VectorXd / Matrix2d / Matrix3d / Vector3d, resize,
// setZero, rows, cols, (i), (i,j), [i], Identity, comma initialiser (scalars and vector blocks), block,
// block<R,C>, topLeftCorner / topRightCorner / bottomLeftCorner / bottomRightCorner / topRows, transpose,
// + - * (matrix and scalar), unary -, +=, inverse (partial-pivot LU, what Eigen does for dynamic sizes),
// norm (Frobenius), operator<< to a stream.
// Default-constructed fixed-size matrices hold NaN (Eigen leaves them uninitialised; its own
// EIGEN_INITIALIZE_MATRICES_BY_NAN option does the same): the reference reads ArucoMarker::
// last_observation_ before ever writing it (aruco_slam.cpp:193 after aruco_slam.h:71-72), and NaN makes that
// comparison false, which is SURVEY App. D item 2's reading of the undefined behaviour.
#ifndef B2A_EIGEN_STUB_H
#define B2A_EIGEN_STUB_H
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <limits>
#include <map>
#include <ostream>
#include <queue>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {

typedef std::ptrdiff_t Index;
const int Dynamic = -1;

class Dense;

// a rectangular view into a Dense (lvalue or rvalue use)
class Block {
public:
    Block(Dense &m, Index r0, Index c0, Index nr, Index nc) : m_(m), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
    Index rows() const { return nr_; }
    Index cols() const { return nc_; }
    inline double &operator()(Index i, Index j);
    inline double operator()(Index i, Index j) const;
    inline Block &operator=(const Dense &o);
    inline Block &operator=(const Block &o);
    inline Block &operator+=(const Dense &o);
    inline Block &operator+=(const Block &o);
    inline double norm() const;
private:
    Dense &m_;
    Index r0_, c0_, nr_, nc_;
};

class CommaInit {
public:
    CommaInit(Dense &m) : m_(m), row_(0), col_(0), blk_(1) {}
    inline CommaInit &add(double v);
    inline CommaInit &add(const Dense &b);
    template <class T> CommaInit &operator,(const T &v) { return add(v); }
private:
    CommaInit &add(int v) { return add((double)v); }
    CommaInit &add(float v) { return add((double)v); }
    Dense &m_;
    Index row_, col_, blk_;
};

class Dense {
public:
    Dense() : r_(0), c_(0) {}
    Dense(Index r, Index c) : r_(r), c_(c), v_((size_t)(r * c), 0.0) {}
    Dense(const Block &b) : r_(b.rows()), c_(b.cols()), v_((size_t)(b.rows() * b.cols()))
    {
        for (Index i = 0; i < r_; ++i) for (Index j = 0; j < c_; ++j) (*this)(i, j) = b(i, j);
    }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    void resize(Index r, Index c) { r_ = r; c_ = c; v_.assign((size_t)(r * c), 0.0); }
    void resize(Index n) { resize(n, 1); }                       // vectors
    Dense &setZero() { std::fill(v_.begin(), v_.end(), 0.0); return *this; }
    double &operator()(Index i, Index j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v_[(size_t)(i * c_ + j)]; }
    double operator()(Index i, Index j) const { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v_[(size_t)(i * c_ + j)]; }
    // linear index: column-major order as in Eigen; the reference only uses it on N x 1 objects
    double &operator()(Index k) { return c_ == 1 ? (*this)(k, 0) : (*this)(k % r_, k / r_); }
    double operator()(Index k) const { return c_ == 1 ? (*this)(k, 0) : (*this)(k % r_, k / r_); }
    double &operator[](Index k) { return (*this)(k); }
    double operator[](Index k) const { return (*this)(k); }
    const double *data() const { return v_.data(); }              // row-major (stub only)
    double *data() { return v_.data(); }

    Block block(Index r0, Index c0, Index nr, Index nc) { return Block(*this, r0, c0, nr, nc); }
    template <int NR, int NC> Block block(Index r0, Index c0) { return Block(*this, r0, c0, NR, NC); }
    Block topLeftCorner(Index nr, Index nc) { return Block(*this, 0, 0, nr, nc); }
    Block topRightCorner(Index nr, Index nc) { return Block(*this, 0, c_ - nc, nr, nc); }
    Block bottomLeftCorner(Index nr, Index nc) { return Block(*this, r_ - nr, 0, nr, nc); }
    Block bottomRightCorner(Index nr, Index nc) { return Block(*this, r_ - nr, c_ - nc, nr, nc); }
    Block topRows(Index nr) { return Block(*this, 0, 0, nr, c_); }

    Dense transpose() const
    {
        Dense t(c_, r_);
        for (Index i = 0; i < r_; ++i) for (Index j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    double norm() const
    {
        double s = 0;
        for (double x : v_) s += x * x;
        return std::sqrt(s);
    }
    Dense inverse() const
    {   // Gauss-Jordan with partial pivoting
        assert(r_ == c_);
        const Index n = r_;
        Dense a(*this), inv(n, n);
        for (Index i = 0; i < n; ++i) inv(i, i) = 1.0;
        for (Index col = 0; col < n; ++col) {
            Index piv = col;
            for (Index i = col + 1; i < n; ++i) if (std::fabs(a(i, col)) > std::fabs(a(piv, col))) piv = i;
            if (piv != col) for (Index j = 0; j < n; ++j) { std::swap(a(piv, j), a(col, j)); std::swap(inv(piv, j), inv(col, j)); }
            const double d = 1.0 / a(col, col);
            for (Index j = 0; j < n; ++j) { a(col, j) *= d; inv(col, j) *= d; }
            for (Index i = 0; i < n; ++i) {
                if (i == col) continue;
                const double f = a(i, col);
                if (f == 0.0) continue;
                for (Index j = 0; j < n; ++j) { a(i, j) -= f * a(col, j); inv(i, j) -= f * inv(col, j); }
            }
        }
        return inv;
    }
    Dense &operator+=(const Dense &o)
    {
        assert(r_ == o.r_ && c_ == o.c_);
        for (size_t k = 0; k < v_.size(); ++k) v_[k] += o.v_[k];
        return *this;
    }
    CommaInit operator<<(double v) { CommaInit ci(*this); ci.add(v); return ci; }
    CommaInit operator<<(const Dense &b) { CommaInit ci(*this); ci.add(b); return ci; }
protected:
    Index r_, c_;
    std::vector<double> v_;
};

inline double &Block::operator()(Index i, Index j) { return m_(r0_ + i, c0_ + j); }
inline double Block::operator()(Index i, Index j) const { return const_cast<const Dense &>(m_)(r0_ + i, c0_ + j); }
inline Block &Block::operator=(const Dense &o)
{
    assert(o.rows() == nr_ && o.cols() == nc_);
    for (Index i = 0; i < nr_; ++i) for (Index j = 0; j < nc_; ++j) (*this)(i, j) = o(i, j);
    return *this;
}
inline Block &Block::operator=(const Block &o) { return *this = Dense(o); }
inline Block &Block::operator+=(const Dense &o)
{
    assert(o.rows() == nr_ && o.cols() == nc_);
    for (Index i = 0; i < nr_; ++i) for (Index j = 0; j < nc_; ++j) (*this)(i, j) += o(i, j);
    return *this;
}
inline Block &Block::operator+=(const Block &o) { return *this += Dense(o); }
inline double Block::norm() const { return Dense(*this).norm(); }

// Eigen's CommaInitializer placement: items fill a row left to right; when the row is full the next item
// starts below the block just finished
inline CommaInit &CommaInit::add(double v)
{
    if (col_ == m_.cols()) { row_ += blk_; col_ = 0; blk_ = 1; }
    m_(row_, col_) = v;
    col_ += 1;
    return *this;
}
inline CommaInit &CommaInit::add(const Dense &b)
{
    if (col_ == m_.cols()) { row_ += blk_; col_ = 0; }
    if (col_ == 0) blk_ = b.rows();
    for (Index i = 0; i < b.rows(); ++i) for (Index j = 0; j < b.cols(); ++j) m_(row_ + i, col_ + j) = b(i, j);
    col_ += b.cols();
    return *this;
}

inline Dense operator+(const Dense &a, const Dense &b)
{
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    Dense r(a.rows(), a.cols());
    for (Index i = 0; i < a.rows(); ++i) for (Index j = 0; j < a.cols(); ++j) r(i, j) = a(i, j) + b(i, j);
    return r;
}
inline Dense operator-(const Dense &a, const Dense &b)
{
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    Dense r(a.rows(), a.cols());
    for (Index i = 0; i < a.rows(); ++i) for (Index j = 0; j < a.cols(); ++j) r(i, j) = a(i, j) - b(i, j);
    return r;
}
inline Dense operator-(const Dense &a)
{
    Dense r(a.rows(), a.cols());
    for (Index i = 0; i < a.rows(); ++i) for (Index j = 0; j < a.cols(); ++j) r(i, j) = -a(i, j);
    return r;
}
inline Dense operator*(const Dense &a, const Dense &b)
{
    assert(a.cols() == b.rows());
    const Index n = a.rows(), m = b.cols(), K = a.cols();
    Dense r(n, m);
    const double *A = a.data(), *B = b.data();
    double *R = r.data();
    for (Index i = 0; i < n; ++i)
        for (Index k = 0; k < K; ++k) {                 // k ascending for every (i, j): r(i,j) = sum_k a(i,k) b(k,j) in order
            const double x = A[i * K + k];
            if (x == 0.0) continue;                     // the reference's F, Hx, I - K Gx are mostly zeros; 0 * finite adds nothing
            const double *Bk = B + k * m;
            double *Ri = R + i * m;
            for (Index j = 0; j < m; ++j) Ri[j] += x * Bk[j];
        }
    return r;
}
inline Dense operator*(double s, const Dense &a)
{
    Dense r(a.rows(), a.cols());
    for (Index i = 0; i < a.rows(); ++i) for (Index j = 0; j < a.cols(); ++j) r(i, j) = s * a(i, j);
    return r;
}
inline Dense operator*(const Dense &a, double s) { return s * a; }
inline std::ostream &operator<<(std::ostream &os, const Dense &m)
{
    for (Index i = 0; i < m.rows(); ++i) { for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.rows()) os << "\n"; }
    return os;
}

template <class Scalar, int R, int C>
class Matrix : public Dense {
public:
    Matrix() : Dense(R > 0 ? R : 0, C > 0 ? C : 0)
    {
        if (R > 0 && C > 0) std::fill(v_.begin(), v_.end(), std::numeric_limits<double>::quiet_NaN());
    }
    Matrix(const Dense &o) : Dense(o) { check(); }
    Matrix(const Block &b) : Dense(b) { check(); }
    explicit Matrix(Index n) : Dense(C == 1 ? n : (R == 1 ? 1 : n), C == 1 ? 1 : n) {}             // VectorXd(n)
    Matrix(Index r, Index c) : Dense(r, c) {}                                                      // MatrixXd(r, c)
    Matrix(double a, double b, double c) : Dense(R, C) { (*this)(0) = a; (*this)(1) = b; (*this)(2) = c; }   // Vector3d(x, y, z)
    Matrix &operator=(const Dense &o) { Dense::operator=(o); check(); return *this; }
    Matrix &operator=(const Block &b) { Dense::operator=(Dense(b)); check(); return *this; }
    static Matrix Identity()
    {
        Matrix m; m.setZero();
        for (Index i = 0; i < std::min<Index>(m.rows(), m.cols()); ++i) m(i, i) = 1.0;
        return m;
    }
    static Matrix Identity(Index r, Index c)
    {
        Matrix m(r, c);
        for (Index i = 0; i < std::min(r, c); ++i) m(i, i) = 1.0;
        return m;
    }
private:
    void check() const { assert((R < 0 || rows() == R) && (C < 0 || cols() == C)); }
};

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;

}  // namespace Eigen
#endif
