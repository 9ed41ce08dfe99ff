/*
 * ref_standins.h — stand-ins that let the reference's OWN class text
 * (/root/reference/include/descriptor.h, classes scan_descriptor and scan_context_descriptor)
 * compile in this container, where Eigen, PCL, ROS and libnabo are absent.
 *
 * TEST INFRASTRUCTURE ONLY (used by oracle/ref_driver.cpp to build oracle/_ref/libscl_ref.so).
 * This is our code, not the reference's: just enough of each library's surface for the
 * expressions that class uses. What the stand-ins decide (and therefore what stays
 * "parity unpinned"): reductions run in index order (Eigen vectorises them), and the libnabo
 * stand-in is a linear scan with libnabo's published rules (sequential squared distance,
 * ascending results, d2 <= epsilon skipped without ALLOW_SELF_MATCH, strict < against the worst).
 */
#ifndef REF_STANDINS_H_
#define REF_STANDINS_H_

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <memory>
#include <utility>
#include <vector>

namespace Eigen {

template <typename T> class Mat;

template <typename T> struct BlockRef {
    Mat<T>* m; int r0, c0, nr, nc;
    T& at(int r, int c) const { return (*m)(r0 + r, c0 + c); }
    BlockRef& operator=(const BlockRef& o)
    {
        assert(nr == o.nr && nc == o.nc);
        for (int c = 0; c < nc; c++) for (int r = 0; r < nr; r++) at(r, c) = o.at(r, c);
        return *this;
    }
    BlockRef& operator=(const Mat<T>& o)
    {
        assert(nr == o.rows() && nc == o.cols());
        for (int c = 0; c < nc; c++) for (int r = 0; r < nr; r++) at(r, c) = o(r, c);
        return *this;
    }
};

/* dynamic, column-major, like Eigen::Matrix<T, Dynamic, Dynamic> */
template <typename T> class Mat {
public:
    Mat() : r_(0), c_(0) {}
    Mat(int r, int c) : r_(r), c_(c), d_((size_t)r * c) {}
    Mat(const BlockRef<T>& b) : r_(b.nr), c_(b.nc), d_((size_t)b.nr * b.nc)
    {
        for (int c = 0; c < c_; c++) for (int r = 0; r < r_; r++) (*this)(r, c) = b.at(r, c);
    }
    static Mat Ones(int r, int c) { Mat m(r, c); std::fill(m.d_.begin(), m.d_.end(), T(1)); return m; }
    static Mat Zero(int r, int c) { Mat m(r, c); std::fill(m.d_.begin(), m.d_.end(), T(0)); return m; }
    T& operator()(int r, int c) { return d_[(size_t)c * r_ + r]; }
    const T& operator()(int r, int c) const { return d_[(size_t)c * r_ + r]; }
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    T* data() { return d_.data(); }
    const T* data() const { return d_.data(); }
    BlockRef<T> block(int r0, int c0, int nr, int nc) const { return BlockRef<T>{const_cast<Mat*>(this), r0, c0, nr, nc}; }
    BlockRef<T> row(int i) const { return block(i, 0, 1, c_); }
    BlockRef<T> col(int i) const { return block(0, i, r_, 1); }
    T sum() const { T s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i]; return s; }
    T mean() const { return sum() / T(size()); }
    T squaredNorm() const { T s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i] * d_[i]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    T dot(const Mat& o) const { T s = 0; for (size_t i = 0; i < d_.size(); i++) s += d_[i] * o.d_[i]; return s; }
    Mat operator-(const Mat& o) const
    {
        assert(r_ == o.r_ && c_ == o.c_);
        Mat m(r_, c_);
        for (size_t i = 0; i < d_.size(); i++) m.d_[i] = d_[i] - o.d_[i];
        return m;
    }
    void conservativeResize(int r, int c)
    {
        Mat m = Zero(r, c);
        for (int cc = 0; cc < std::min(c, c_); cc++) for (int rr = 0; rr < std::min(r, r_); rr++) m(rr, cc) = (*this)(rr, cc);
        *this = m;
    }
protected:
    int r_, c_;
    std::vector<T> d_;
};

template <typename T> Mat<T> operator*(int s, const Mat<T>& m)
{
    Mat<T> o(m.rows(), m.cols());
    for (int c = 0; c < m.cols(); c++) for (int r = 0; r < m.rows(); r++) o(r, c) = T(s) * m(r, c);
    return o;
}

template <typename T> class Vec : public Mat<T> {
public:
    Vec() {}
    explicit Vec(int n) : Mat<T>(n, 1) {}
    Vec(const BlockRef<T>& b) : Mat<T>(b) {}
    T& operator[](int i) { return this->d_[i]; }
    const T& operator[](int i) const { return this->d_[i]; }
};

typedef Mat<double> MatrixXd;
typedef Mat<float> MatrixXf;
typedef Vec<double> VectorXd;
typedef Vec<float> VectorXf;
typedef Vec<int> VectorXi;

} // namespace Eigen

namespace pcl {
struct alignas(16) PointXYZI { float x, y, z, pad0; float intensity, pad1, pad2, pad3; };
template <typename P> struct PointCloud { std::vector<P> points; };
} // namespace pcl

namespace Nabo {
struct NNSearchF {
    Eigen::MatrixXf cloud; int dim;
    static NNSearchF* createKDTreeLinearHeap(const Eigen::MatrixXf& c, int d) { NNSearchF* s = new NNSearchF(); s->cloud = c; s->dim = d; return s; }
    void knn(const Eigen::VectorXf& q, Eigen::VectorXi& idx, Eigen::VectorXf& d2, int k) const
    {
        int count = 0;
        for (int i = 0; i < k; i++) { idx[i] = -1; d2[i] = std::numeric_limits<float>::infinity(); }
        for (int j = 0; j < cloud.cols(); j++) {
            float dist = 0;
            for (int d = 0; d < dim; d++) { const float diff = q[d] - cloud(d, j); dist += diff * diff; }
            if (!(dist > std::numeric_limits<float>::epsilon())) continue;
            if (!(dist < (count == k ? d2[k - 1] : std::numeric_limits<float>::infinity()))) continue;
            int i = count < k ? count : k - 1;
            for (; i > 0 && d2[i - 1] > dist; --i) { d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; }
            d2[i] = dist; idx[i] = j;
            if (count < k) count++;
        }
    }
};
} // namespace Nabo

#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)

#endif
