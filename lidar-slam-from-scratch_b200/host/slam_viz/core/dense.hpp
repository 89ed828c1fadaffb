// dense.hpp — the matrix types of the public API.
//
// The reference's core headers use Eigen (types.hpp:3).  Where Eigen is installed this header simply includes it.
// Where it is not (this build image has no Eigen and no network) a minimal stand-in provides the handful of types
// and members the API surface needs — sizes, element access, contiguous data() — in namespace Eigen, so that code
// written against the reference's signatures compiles unchanged.  The mirror headers themselves only use members
// that both provide: rows(), cols(), size(), data(), resize(), operator()(i,j), operator()(i).
#pragma once
#if !defined(SLAM_B200_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define SLAM_B200_HAVE_EIGEN 1
#endif
#endif

#ifndef SLAM_B200_HAVE_EIGEN
#include <cstddef>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
constexpr int ColMajor = 0;
constexpr int RowMajor = 1;

template <typename T, int R, int C, int Options = ColMajor>
class Matrix {
public:
    Matrix() : rows_(R == Dynamic ? 0 : R), cols_(C == Dynamic ? 0 : C), v_((size_t)rows_ * cols_, T(0)) {}
    Matrix(long r, long c) : rows_(r), cols_(c), v_((size_t)r * c, T(0)) {}
    explicit Matrix(long n) : rows_(C == 1 ? n : 1), cols_(C == 1 ? 1 : n), v_((size_t)n, T(0)) {}
    Matrix(T x, T y, T z) : rows_(R == Dynamic ? 3 : R), cols_(C == Dynamic ? 1 : C), v_{x, y, z} {}
    long rows() const { return rows_; }
    long cols() const { return cols_; }
    long size() const { return rows_ * cols_; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
    void resize(long r, long c) { rows_ = r; cols_ = c; v_.assign((size_t)r * c, T(0)); }
    void resize(long n) { if (C == 1) resize(n, 1); else resize(1, n); }
    void setZero() { for (auto& x : v_) x = T(0); }
    T& operator()(long i, long j) { return v_[index(i, j)]; }
    const T& operator()(long i, long j) const { return v_[index(i, j)]; }
    T& operator()(long i) { return v_[(size_t)i]; }
    const T& operator()(long i) const { return v_[(size_t)i]; }
    T& operator[](long i) { return v_[(size_t)i]; }
    const T& operator[](long i) const { return v_[(size_t)i]; }
    // row(i): a view of one row, just enough for PointCloud::row (types.hpp:36-37)
    struct RowView {
        T* p;
        long stride, n;
        T& operator()(long j) const { return p[j * stride]; }
        long size() const { return n; }
        long cols() const { return n; }
    };
    struct ConstRowView {
        const T* p;
        long stride, n;
        const T& operator()(long j) const { return p[j * stride]; }
        long size() const { return n; }
        long cols() const { return n; }
    };
    RowView row(long i) { return RowView{&v_[index(i, 0)], Options == RowMajor ? 1 : rows_, cols_}; }
    ConstRowView row(long i) const { return ConstRowView{&v_[index(i, 0)], Options == RowMajor ? 1 : rows_, cols_}; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() {
        Matrix m;
        for (long i = 0; i < m.rows_ && i < m.cols_; ++i) m(i, i) = T(1);
        return m;
    }

private:
    size_t index(long i, long j) const {
        return Options == RowMajor ? (size_t)i * cols_ + j : (size_t)j * rows_ + i;
    }
    long rows_, cols_;
    std::vector<T> v_;
};

using Vector3d = Matrix<double, 3, 1>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using VectorXd = Matrix<double, Dynamic, 1>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;

}  // namespace Eigen
#endif
