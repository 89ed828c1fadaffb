// types.hpp — mirror of slam_viz/include/slam_viz/core/types.hpp:15-164 (same names, members and defaults).
// Pure host value types; nothing here touches the GPU.
#pragma once
#include <utility>
#include <vector>

#include "dense.hpp"

namespace slam {

// types.hpp:15-61
class PointCloud {
public:
    using Matrix = Eigen::Matrix<double, Eigen::Dynamic, 3, Eigen::RowMajor>;
    using Vector3 = Eigen::Vector3d;

    PointCloud() = default;
    explicit PointCloud(const Matrix& points) : points_(points) {}
    explicit PointCloud(Matrix&& points) : points_(std::move(points)) {}
    explicit PointCloud(const std::vector<Vector3>& points) {
        points_.resize((long)points.size(), 3);
        for (size_t i = 0; i < points.size(); ++i)
            for (int a = 0; a < 3; ++a) points_((long)i, a) = points[i](a);
    }

    const Matrix& points() const { return points_; }
    Matrix& points() { return points_; }
    size_t size() const { return static_cast<size_t>(points_.rows()); }
    bool empty() const { return points_.rows() == 0; }
    auto row(int i) const { return points_.row(i); }  // types.hpp:36-37
    auto row(int i) { return points_.row(i); }

    Vector3 centroid() const {  // types.hpp:44-46
        Vector3 c;
        for (int a = 0; a < 3; ++a) {
            double s = 0.0;
            for (long i = 0; i < points_.rows(); ++i) s += points_(i, a);
            c(a) = s / (double)points_.rows();
        }
        return c;
    }
    PointCloud centered() const {  // types.hpp:49-52
        Vector3 c = centroid();
        Matrix m(points_.rows(), 3);
        for (long i = 0; i < points_.rows(); ++i)
            for (int a = 0; a < 3; ++a) m(i, a) = points_(i, a) - c(a);
        return PointCloud(std::move(m));
    }
    PointCloud copy() const { return PointCloud(Matrix(points_)); }

private:
    Matrix points_;
};

// types.hpp:74-136
class Transformation {
public:
    using Matrix4 = Eigen::Matrix4d;
    using Matrix3 = Eigen::Matrix3d;
    using Vector3 = Eigen::Vector3d;

    Transformation() : matrix_(Matrix4::Identity()) {}
    explicit Transformation(const Matrix4& matrix) : matrix_(matrix) {}
    Transformation(const Matrix3& R, const Vector3& t) : matrix_(Matrix4::Identity()) {
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) matrix_(i, j) = R(i, j);
            matrix_(i, 3) = t(i);
        }
    }
    static Transformation from_rt(const Matrix3& R, const Vector3& t) { return Transformation(R, t); }
    static Transformation identity() { return Transformation(); }
    // row-major 4x4 as the C ABI carries it (sb_icp_result::transformation)
    static Transformation from_row_major(const double* m16) {
        Matrix4 m;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) m(i, j) = m16[4 * i + j];
        return Transformation(m);
    }
    void to_row_major(double* m16) const {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) m16[4 * i + j] = matrix_(i, j);
    }

    const Matrix4& matrix() const { return matrix_; }
    Matrix3 R() const {
        Matrix3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r(i, j) = matrix_(i, j);
        return r;
    }
    Vector3 t() const {
        Vector3 v;
        for (int i = 0; i < 3; ++i) v(i) = matrix_(i, 3);
        return v;
    }
    Vector3 apply(const Vector3& p) const {  // types.hpp:104-106
        Vector3 q;
        for (int i = 0; i < 3; ++i)
            q(i) = ((matrix_(i, 0) * p(0) + matrix_(i, 1) * p(1)) + matrix_(i, 2) * p(2)) + matrix_(i, 3);
        return q;
    }
    PointCloud apply(const PointCloud& cloud) const {  // types.hpp:109-115: P * R^T + t^T
        PointCloud::Matrix out(cloud.points().rows(), 3);
        for (long r = 0; r < cloud.points().rows(); ++r)
            for (int i = 0; i < 3; ++i)
                out(r, i) = ((cloud.points()(r, 0) * matrix_(i, 0) + cloud.points()(r, 1) * matrix_(i, 1)) +
                             cloud.points()(r, 2) * matrix_(i, 2)) + matrix_(i, 3);
        return PointCloud(std::move(out));
    }
    Transformation compose(const Transformation& other) const {  // types.hpp:118-120: this * other
        Matrix4 m;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0.0;
                for (int k = 0; k < 4; ++k) s += matrix_(i, k) * other.matrix_(k, j);
                m(i, j) = s;
            }
        return Transformation(m);
    }
    Transformation operator*(const Transformation& other) const { return compose(other); }
    Transformation inverse() const {  // types.hpp:128-132: (R^T, -R^T t)
        Matrix3 ri;
        Vector3 ti;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) ri(i, j) = matrix_(j, i);
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int j = 0; j < 3; ++j) s += ri(i, j) * matrix_(j, 3);
            ti(i) = -s;
        }
        return from_rt(ri, ti);
    }

private:
    Matrix4 matrix_;
};

// types.hpp:143-148
struct ICPConfig {
    int max_iterations = 50;
    double tolerance = 1e-6;
    double min_error = 1e-9;
    Transformation initial_transform = Transformation::identity();
};

// types.hpp:155-164
struct ICPResult {
    Transformation transformation;
    bool converged = false;
    int num_iterations = 0;
    std::vector<double> error_history;
    double final_error = 0.0;
    bool success() const { return converged && final_error < 0.1; }
};

}  // namespace slam
