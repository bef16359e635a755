// ============================================================================
// oracle/qp_stubs/OsqpEigen/OsqpEigen.h -- stand-in for OsqpEigen + the slice of Eigen that
// src/planner/include/planner/qp_solver.hpp touches, so that THAT FILE compiles verbatim from /root/reference
// (oracle/Makefile target `ref`) and the matrices it builds can be captured.  TEST INFRASTRUCTURE ONLY.
//
// * namespace Eigen: a small dynamic dense matrix (row-major storage, eager evaluation) with exactly the operations the
//   reference uses: resize / setZero / (i,j) / block / row / col / middleRows / transpose / segment / unary minus /
//   scalar * / column * row (outer product) / comma initialisation with scalars, vectors and matrices / sparseView.
// * namespace OsqpEigen: Solver whose data() setters record the problem (Hessian, gradient, constraint matrix, bounds);
//   solveProblem() does not solve anything: it reports success with a zero solution.  The point is the capture.
// ============================================================================
#pragma once
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <vector>

namespace Eigen {

class Mat;
// read-only expression result: always a materialised matrix
class Mat {
public:
    Mat() {}
    Mat(std::size_t r, std::size_t c) : r_(r), c_(c), d_(r * c, 0.0) {}
    explicit Mat(std::size_t n) : r_(n), c_(1), d_(n, 0.0) {}
    std::size_t rows() const { return r_; }
    std::size_t cols() const { return c_; }
    std::size_t size() const { return d_.size(); }
    void resize(std::size_t r, std::size_t c) { r_ = r; c_ = c; d_.assign(r * c, 0.0); }
    void resize(std::size_t n) { resize(n, 1); }
    void setZero() { std::fill(d_.begin(), d_.end(), 0.0); }
    double &operator()(std::size_t i, std::size_t j) { return d_[i * c_ + j]; }
    double operator()(std::size_t i, std::size_t j) const { return d_[i * c_ + j]; }
    double &operator()(std::size_t i) { return d_[i]; }          // vectors (one column or one row)
    double operator()(std::size_t i) const { return d_[i]; }
    const double *data() const { return d_.data(); }

    struct Block {
        Mat &m; std::size_t i0, j0, p, q;
        Block &operator=(const Mat &o) {
            if (o.size() != p * q) throw std::runtime_error("qp_stubs: block size mismatch");   // Eigen would assert
            // shapes may be (p,q) or a vector laid out the other way (segment = row.transpose()); copy in order
            for (std::size_t a = 0; a < p; ++a)
                for (std::size_t b = 0; b < q; ++b) m(i0 + a, j0 + b) = (o.rows() == p) ? o(a, b) : o.d_[a * q + b];
            return *this;
        }
        Block &operator=(const Block &o) { return *this = Mat(o); }
        operator Mat() const {
            Mat r(p, q);
            for (std::size_t a = 0; a < p; ++a)
                for (std::size_t b = 0; b < q; ++b) r(a, b) = m(i0 + a, j0 + b);
            return r;
        }
        Mat transpose() const { return Mat(*this).transpose(); }
        Mat operator-() const { return -Mat(*this); }
        Mat operator*(const Block &o) const { return Mat(*this) * Mat(o); }
        Mat operator*(const Mat &o) const { return Mat(*this) * o; }
    };
    Block block(std::size_t i, std::size_t j, std::size_t p, std::size_t q) { return Block{*this, i, j, p, q}; }
    Mat block(std::size_t i, std::size_t j, std::size_t p, std::size_t q) const { return Mat(Block{const_cast<Mat &>(*this), i, j, p, q}); }
    Block row(std::size_t i) { return Block{*this, i, 0, 1, c_}; }
    Mat row(std::size_t i) const { return block(i, 0, 1, c_); }
    Block col(std::size_t j) { return Block{*this, 0, j, r_, 1}; }
    Mat col(std::size_t j) const { return block(0, j, r_, 1); }
    Block middleRows(std::size_t i, std::size_t n) { return Block{*this, i, 0, n, c_}; }
    Mat middleRows(std::size_t i, std::size_t n) const { return block(i, 0, n, c_); }
    Block segment(std::size_t i, std::size_t n) { return c_ == 1 ? Block{*this, i, 0, n, 1} : Block{*this, 0, i, 1, n}; }
    Mat transpose() const {
        Mat r(c_, r_);
        for (std::size_t a = 0; a < r_; ++a)
            for (std::size_t b = 0; b < c_; ++b) r(b, a) = (*this)(a, b);
        return r;
    }
    Mat operator-() const { Mat r = *this; for (double &v : r.d_) v = -v; return r; }
    Mat operator*(const Mat &o) const {
        if (c_ != o.r_) throw std::runtime_error("qp_stubs: product size mismatch");
        Mat r(r_, o.c_);
        for (std::size_t a = 0; a < r_; ++a)
            for (std::size_t b = 0; b < o.c_; ++b) {
                double s = 0.0;
                for (std::size_t k = 0; k < c_; ++k) s += (*this)(a, k) * o(k, b);
                r(a, b) = s;
            }
        return r;
    }
    Mat operator*(const Block &o) const { return *this * Mat(o); }
    friend Mat operator*(double s, const Mat &m) { Mat r = m; for (double &v : r.d_) v *= s; return r; }
    static Mat Ones(std::size_t n) { Mat r(n, 1); for (double &v : r.d_) v = 1.0; return r; }
    const Mat &sparseView() const { return *this; }

    // comma initialisation, row-major fill: scalars one by one; a vector or matrix as a block of rows
    struct Comma {
        Mat &m; std::size_t at;
        Comma &push(double v) {
            if (at >= m.size()) throw std::runtime_error("qp_stubs: too many comma-initialiser entries");
            m.d_[at++] = v; return *this;
        }
        Comma &push(const Mat &o) {
            // stacked blocks: every use in qp_solver.hpp stacks full-width matrices / whole vectors vertically
            if (o.cols() != m.cols() && !(m.cols() == 1)) throw std::runtime_error("qp_stubs: comma block width mismatch");
            for (std::size_t k = 0; k < o.size(); ++k) push(o.d_[k]);
            return *this;
        }
        Comma &operator,(double v) { return push(v); }
        Comma &operator,(const Mat &o) { return push(o); }
    };
    Comma operator<<(double v) { Comma c{*this, 0}; c.push(v); return c; }
    Comma operator<<(const Mat &o) { Comma c{*this, 0}; c.push(o); return c; }

private:
    std::size_t r_ = 0, c_ = 0;
    std::vector<double> d_;
    friend struct Block;
    friend struct Comma;
};

typedef Mat MatrixXd;
typedef Mat VectorXd;
typedef Mat MatrixX4d;
struct Vector2d : public Mat { Vector2d() : Mat(2, 1) {} };
template <class T> struct SparseMatrix : public Mat {
    SparseMatrix() {}
    SparseMatrix(const Mat &m) : Mat(m) {}
};
}  // namespace Eigen

namespace OsqpEigen {
enum class ErrorExitFlag { NoError, DataValidationError };
enum class Status { Solved, Unsolved };

// what the last Solver was given (read by oracle/ref_qp.cpp after QPSolver::solve returns)
struct Capture {
    int n = 0, m = 0;
    Eigen::Mat hessian, gradient, constraints, lower, upper;
};
inline Capture &last_capture() { static Capture c; return c; }

class Solver {
public:
    struct Settings {
        void setVerbosity(bool) {}
        void setWarmStart(bool) {}
    };
    struct Data {
        void setNumberOfVariables(int n) { last_capture().n = n; }
        void setNumberOfConstraints(int m) { last_capture().m = m; }
        bool setHessianMatrix(const Eigen::Mat &h) { last_capture().hessian = h; return true; }
        bool setGradient(Eigen::Mat &g) { last_capture().gradient = g; return true; }
        bool setLinearConstraintsMatrix(const Eigen::Mat &a) { last_capture().constraints = a; return true; }
        bool setLowerBound(Eigen::Mat &l) { last_capture().lower = l; return true; }
        bool setUpperBound(Eigen::Mat &u) { last_capture().upper = u; return true; }
    };
    Settings *settings() { return &settings_; }
    Data *data() { return &data_; }
    bool initSolver() { return true; }
    ErrorExitFlag solveProblem() { return ErrorExitFlag::NoError; }
    double getObjValue() const { return 0.0; }
    Status getStatus() const { return Status::Solved; }
    Eigen::Mat getSolution() const { return Eigen::Mat(static_cast<std::size_t>(last_capture().n), 1); }

private:
    Settings settings_;
    Data data_;
};
}  // namespace OsqpEigen
