// oracle/shim/Eigen/EigenShim.h -- TEST INFRASTRUCTURE ONLY.
//
// Eigen is not vendored by the reference and not installed in this image.  The reference uses it
// as a CONTAINER on the decode path (SparseMatrix<int,RowMajor>: setFromTriplets, nonZeros,
// outerSize, InnerIterator -- MyLdpc.cpp:85-109,188-191) and for small integer linear algebra in
// its encoder (MyLdpc.cpp:137-165,633-682; MyLdpc.h:240-337).  This file provides exactly the
// subset of that interface the reference calls, written from Eigen's documented semantics with
// plain std::vector storage, so that MyLdpc.cpp compiles UNMODIFIED into oracle/_ref.  All decode
// arithmetic that oracle/_ref executes is the reference's own code; nothing here computes floats.
//   * setFromTriplets: duplicates are summed, inner indices end up sorted ascending.
//   * InnerIterator walks one outer vector (a row for RowMajor, a column for ColMajor) in
//     ascending inner index.
//   * Expressions are evaluated eagerly; `a.col(i) += s * a.col(j)` therefore reads all of the
//     right-hand side before writing, which equals Eigen's coefficient-wise evaluation for the
//     aliasing patterns the reference uses (same column, or two different columns).
#ifndef ORACLE_SHIM_EIGEN_H_
#define ORACLE_SHIM_EIGEN_H_
#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1 };

template <class T> struct VecTmp { std::vector<T> v; };

template <class T, int R, int C> class Matrix;

template <class T> struct ColXpr {
    Matrix<T, Dynamic, Dynamic> *m;
    int c;
    ColXpr &operator+=(const VecTmp<T> &o);
    ColXpr &operator-=(const VecTmp<T> &o);
};
template <class T> struct RowXpr {
    Matrix<T, Dynamic, Dynamic> *m;
    int r;
    RowXpr &operator+=(const VecTmp<T> &o);
    RowXpr &operator-=(const VecTmp<T> &o);
};

template <class T, int R, int C> class Matrix {
public:
    Matrix() : r_(0), c_(0) {}
    Matrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c, T(0)) {}
    int rows() const { return r_; }
    int cols() const { return c_; }
    void resize(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, T(0)); }
    T &operator()(int r, int c) { return d_[(size_t)r * c_ + c]; }
    const T &operator()(int r, int c) const { return d_[(size_t)r * c_ + c]; }
    ColXpr<T> col(int c) { ColXpr<T> x = {this, c}; return x; }
    RowXpr<T> row(int r) { RowXpr<T> x = {this, r}; return x; }
    Matrix operator-() const { Matrix o(r_, c_); for (size_t i = 0; i < d_.size(); ++i) o.d_[i] = -d_[i]; return o; }
    Matrix operator+(const Matrix &b) const { Matrix o(r_, c_); for (size_t i = 0; i < d_.size(); ++i) o.d_[i] = d_[i] + b.d_[i]; return o; }
    Matrix operator-(const Matrix &b) const { Matrix o(r_, c_); for (size_t i = 0; i < d_.size(); ++i) o.d_[i] = d_[i] - b.d_[i]; return o; }
    Matrix operator*(const Matrix &b) const {
        Matrix o(r_, b.c_);
        for (int i = 0; i < r_; ++i)
            for (int k = 0; k < c_; ++k) {
                const T a = (*this)(i, k);
                if (a == T(0)) continue;
                for (int j = 0; j < b.c_; ++j) o(i, j) += a * b(k, j);
            }
        return o;
    }
private:
    int r_, c_;
    std::vector<T> d_;
};

template <class T> VecTmp<T> operator*(T s, const ColXpr<T> &x) {
    VecTmp<T> t; t.v.resize(x.m->rows());
    for (int i = 0; i < x.m->rows(); ++i) t.v[i] = s * (*x.m)(i, x.c);
    return t;
}
template <class T> VecTmp<T> operator*(T s, const RowXpr<T> &x) {
    VecTmp<T> t; t.v.resize(x.m->cols());
    for (int j = 0; j < x.m->cols(); ++j) t.v[j] = s * (*x.m)(x.r, j);
    return t;
}
template <class T> ColXpr<T> &ColXpr<T>::operator+=(const VecTmp<T> &o) { for (int i = 0; i < m->rows(); ++i) (*m)(i, c) += o.v[i]; return *this; }
template <class T> ColXpr<T> &ColXpr<T>::operator-=(const VecTmp<T> &o) { for (int i = 0; i < m->rows(); ++i) (*m)(i, c) -= o.v[i]; return *this; }
template <class T> RowXpr<T> &RowXpr<T>::operator+=(const VecTmp<T> &o) { for (int j = 0; j < m->cols(); ++j) (*m)(r, j) += o.v[j]; return *this; }
template <class T> RowXpr<T> &RowXpr<T>::operator-=(const VecTmp<T> &o) { for (int j = 0; j < m->cols(); ++j) (*m)(r, j) -= o.v[j]; return *this; }

template <class T> class Triplet {
public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(int r, int c, const T &v = T(0)) : r_(r), c_(c), v_(v) {}
    int row() const { return r_; }
    int col() const { return c_; }
    const T &value() const { return v_; }
private:
    int r_, c_;
    T v_;
};

template <class T, int Options = ColMajor> class SparseMatrix {
public:
    typedef std::vector<std::pair<int, T> > Vec;
    SparseMatrix() : rows_(0), cols_(0) {}
    SparseMatrix(int r, int c) : rows_(r), cols_(c), outer_(Options == RowMajor ? r : c) {}
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    int outerSize() const { return (int)outer_.size(); }
    int innerSize() const { return Options == RowMajor ? cols_ : rows_; }
    void resize(int r, int c) { rows_ = r; cols_ = c; outer_.assign(Options == RowMajor ? r : c, Vec()); }
    void setZero() { for (size_t k = 0; k < outer_.size(); ++k) outer_[k].clear(); }
    void makeCompressed() {}
    int nonZeros() const { size_t n = 0; for (size_t k = 0; k < outer_.size(); ++k) n += outer_[k].size(); return (int)n; }
    T &insert(int r, int c) {
        Vec &v = outer_[Options == RowMajor ? r : c];
        const int inner = Options == RowMajor ? c : r;
        typename Vec::iterator it = v.begin();
        while (it != v.end() && it->first < inner) ++it;
        it = v.insert(it, std::make_pair(inner, T(0)));
        return it->second;
    }
    T coeff(int r, int c) const {
        const Vec &v = outer_[Options == RowMajor ? r : c];
        const int inner = Options == RowMajor ? c : r;
        for (size_t i = 0; i < v.size(); ++i) if (v[i].first == inner) return v[i].second;
        return T(0);
    }
    template <class It> void setFromTriplets(It b, It e) {
        setZero();
        for (It t = b; t != e; ++t) outer_[Options == RowMajor ? t->row() : t->col()].push_back(
            std::make_pair(Options == RowMajor ? t->col() : t->row(), t->value()));
        for (size_t k = 0; k < outer_.size(); ++k) {
            Vec &v = outer_[k];
            std::stable_sort(v.begin(), v.end(), [](const std::pair<int, T> &x, const std::pair<int, T> &y) { return x.first < y.first; });
            Vec m;
            for (size_t i = 0; i < v.size(); ++i) {
                if (!m.empty() && m.back().first == v[i].first) m.back().second += v[i].second;  // duplicates are summed
                else m.push_back(v[i]);
            }
            v.swap(m);
        }
    }
    Matrix<T, Dynamic, Dynamic> block(int r0, int c0, int nr, int nc) const {
        Matrix<T, Dynamic, Dynamic> o(nr, nc);
        for (size_t k = 0; k < outer_.size(); ++k)
            for (size_t i = 0; i < outer_[k].size(); ++i) {
                const int r = Options == RowMajor ? (int)k : outer_[k][i].first;
                const int c = Options == RowMajor ? outer_[k][i].first : (int)k;
                if (r >= r0 && r < r0 + nr && c >= c0 && c < c0 + nc) o(r - r0, c - c0) = outer_[k][i].second;
            }
        return o;
    }
    class InnerIterator {
    public:
        InnerIterator(const SparseMatrix &m, int outer) : v_(const_cast<Vec *>(&m.outer_[outer])), outer_(outer), i_(0) {}
        operator bool() const { return i_ < v_->size(); }
        InnerIterator &operator++() { ++i_; return *this; }
        int row() const { return Options == RowMajor ? outer_ : (*v_)[i_].first; }
        int col() const { return Options == RowMajor ? (*v_)[i_].first : outer_; }
        int index() const { return (*v_)[i_].first; }
        const T &value() const { return (*v_)[i_].second; }
        T &valueRef() { return (*v_)[i_].second; }
    private:
        Vec *v_;
        int outer_;
        size_t i_;
    };
    // element access used by the products below
    const Vec &outerVec(int k) const { return outer_[k]; }
    Vec &outerVec(int k) { return outer_[k]; }
private:
    int rows_, cols_;
    std::vector<Vec> outer_;
    friend class InnerIterator;
};

// column-major sparse * sparse and sparse + sparse (the reference's encoder, MyLdpc.cpp:652-657)
template <class T> SparseMatrix<T, ColMajor> operator*(const SparseMatrix<T, ColMajor> &a, const SparseMatrix<T, ColMajor> &b) {
    SparseMatrix<T, ColMajor> o(a.rows(), b.cols());
    std::vector<T> acc(a.rows());
    std::vector<char> hit(a.rows());
    for (int j = 0; j < b.cols(); ++j) {
        std::fill(acc.begin(), acc.end(), T(0));
        std::fill(hit.begin(), hit.end(), 0);
        const typename SparseMatrix<T, ColMajor>::Vec &bj = b.outerVec(j);
        for (size_t q = 0; q < bj.size(); ++q) {
            const typename SparseMatrix<T, ColMajor>::Vec &ak = a.outerVec(bj[q].first);
            for (size_t p = 0; p < ak.size(); ++p) { acc[ak[p].first] += ak[p].second * bj[q].second; hit[ak[p].first] = 1; }
        }
        for (int i = 0; i < a.rows(); ++i) if (hit[i]) o.outerVec(j).push_back(std::make_pair(i, acc[i]));
    }
    return o;
}
template <class T> SparseMatrix<T, ColMajor> operator+(const SparseMatrix<T, ColMajor> &a, const SparseMatrix<T, ColMajor> &b) {
    SparseMatrix<T, ColMajor> o(a.rows(), a.cols());
    for (int j = 0; j < a.cols(); ++j) {
        const typename SparseMatrix<T, ColMajor>::Vec &x = a.outerVec(j), &y = b.outerVec(j);
        size_t p = 0, q = 0;
        while (p < x.size() || q < y.size()) {
            if (q >= y.size() || (p < x.size() && x[p].first < y[q].first)) o.outerVec(j).push_back(x[p++]);
            else if (p >= x.size() || y[q].first < x[p].first) o.outerVec(j).push_back(y[q++]);
            else { o.outerVec(j).push_back(std::make_pair(x[p].first, x[p].second + y[q].second)); ++p; ++q; }
        }
    }
    return o;
}

}  // namespace Eigen
#endif
