// dind_oracle.cpp — CPU restatement of DataInterpolationsND.jl's batched-evaluation path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product path (the package under
// datainterpolationsnd.jl_b200/, libdind.so, include/dind.h) may import, link or call
// this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg use it, and only as the checker / CPU baseline.
//
// Parity status: PINNED by the reference's own known-answer tests (tests/test_oracle_kat.py
// re-states test/test_interpolations.jl, test/test_derivatives.jl,
// test/test_datainterpolations_comparison.jl and test/test_interface.jl of the reference),
// and cross-checked against SciPy (tests/golden/).  The reference is pure Julia and there
// is no Julia in this image, so it cannot be executed here; there is no oracle/_ref.
// Mixed-kind dimensions (eval_mixed below) have NO reference semantics: parity unpinned,
// pinned only by slice-equivalence against the homogeneous evaluators.
//
// Every function cites the reference file:line (relative to /root/reference) it follows.
// Operation order follows the reference; compile WITHOUT fp contraction / fast-math
// (-ffp-contract=off) so that it rounds like Julia (which never fuses a*b+c on its own).
//
// Array conventions are the reference's (Julia): column-major, 1-based indices inside
// the restated functions, `u` of shape (n_1..n_N, out...), `out` of shape
// (n_points, out...) or (g_1..g_N, out...).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

enum Kind { LINEAR = 0, CONSTANT = 1, BSPLINE = 2 };
constexpr int MAXN = 8;
constexpr int MAXDEG = 15;

// ---------------------------------------------------------------------------
// Julia Base `isless` for IEEE floats (Base/float.jl): total order with
// -0.0 < +0.0 and NaN greater than everything (all NaNs equal).
// ---------------------------------------------------------------------------
template <class T>
inline bool jl_isless(T a, T b) {
    const bool an = std::isnan(a), bn = std::isnan(b);
    if (an) return false;
    if (bn) return true;
    if (a < b) return true;
    if (a == b) return std::signbit(a) && !std::signbit(b);
    return false;
}

// Base.searchsortedfirst(v, x): first index i (1-based) with !isless(v[i], x); n+1 if none.
template <class T>
inline int64_t searchsortedfirst(const T* v, int64_t n, T x) {
    int64_t lo = 0, hi = n + 1;  // Julia's own loop (Base/sort.jl): lo = 0, hi = n+1
    while (lo < hi - 1) {
        int64_t m = lo + ((hi - lo) >> 1);
        if (jl_isless(v[m - 1], x)) lo = m; else hi = m;
    }
    return hi;
}

// Base.searchsortedlast(v, x): last index i (1-based) with !isless(x, v[i]); 0 if none.
template <class T>
inline int64_t searchsortedlast(const T* v, int64_t n, T x) {
    int64_t lo = 0, hi = n + 1;
    while (lo < hi - 1) {
        int64_t m = lo + ((hi - lo) >> 1);
        if (jl_isless(x, v[m - 1])) hi = m; else lo = m;
    }
    return lo;
}

inline int64_t clampi(int64_t x, int64_t lo, int64_t hi) {
    // Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
    return x > hi ? hi : (x < lo ? lo : x);
}

// One interpolation dimension (src/interpolation_dimensions.jl:15-27, 61-76, 119-147).
template <class T>
struct Dim {
    int kind = LINEAR;
    int degree = 0;
    const T* t = nullptr;      // distinct knots
    int64_t n_t = 0;
    std::vector<T> knots_all;  // BSpline only
    const T* t_eval = nullptr;
    int64_t n_eval = 0;
    int64_t n_u() const {      // expected size of u along this dim
        // interpolation_utils.jl:39-52, spline_utils.jl:223-225
        return kind == BSPLINE ? (int64_t)knots_all.size() - degree - 1 : n_t;
    }
};

// src/interpolation_utils.jl:104-134 — get_left / get_idx_bounds / get_idx_shift / get_idx.
template <class T>
inline int64_t get_idx(const Dim<T>& d, T x) {
    const T* t;
    int64_t n;
    if (d.kind == BSPLINE) { t = d.knots_all.data(); n = (int64_t)d.knots_all.size(); }
    else { t = d.t; n = d.n_t; }
    const bool left = (d.kind == LINEAR);                         // :104-105
    int64_t lb = 1, ub_shift = -1;                                // :107
    if (d.kind == BSPLINE) { lb = d.degree + 1; ub_shift = -d.degree - 1; }  // :108-110
    const int64_t idx_shift = (d.kind == LINEAR) ? -1 : 0;        // :112-113
    const int64_t ub = n + ub_shift;                              // :128
    if (left) return clampi(searchsortedfirst(t, n, x) + idx_shift, lb, ub);  // :130
    return clampi(searchsortedlast(t, n, x) + idx_shift, lb, ub);             // :132
}

// src/spline_utils.jl:1-20 — expand_knot_vector_kernel (one "work item" per knot i).
template <class T>
void expand_knot_vector(std::vector<T>& knots_all, const T* t, const int64_t* mult, int64_t n) {
    int64_t total = 0;
    for (int64_t i = 0; i < n; ++i) total += mult[i];
    knots_all.assign((size_t)total, T(0));
    for (int64_t i = 1; i <= n; ++i) {
        int64_t mult_sum = 0;
        for (int64_t j = 1; j <= i - 1; ++j) mult_sum += mult[j - 1];
        for (int64_t k = mult_sum + 1; k <= mult_sum + mult[i - 1]; ++k) knots_all[k - 1] = t[i - 1];
    }
}

// src/spline_utils.jl:22
template <class T>
inline T safe_div(T a, T b) { return b == T(0) ? T(0) : a / b; }

// src/spline_utils.jl:24-50 — one entry of one Cox–de Boor stage. b and K are 1-based here.
template <class T>
inline T cox_de_boor(const T* b, const T* K, T t, int64_t idx, int degree, int d, int k, bool deriv) {
    if (k <= degree - d || k == degree + 2) return T(0);
    const int64_t i = idx + k - degree - 1;
    const T ti = K[i], ti1 = K[i + 1], tip = K[i + d], tip1 = K[i + d + 1];
    if (deriv) {
        return T(d) * (safe_div(b[k], tip - ti) - safe_div(b[k + 1], tip1 - ti1));
    }
    return b[k] * safe_div(t - ti, tip - ti) + b[k + 1] * safe_div(tip1 - t, tip1 - ti1);
}

// src/spline_utils.jl:64-114 — get_basis_function_values (on the fly) +
// _compute_basis_function_values.  Writes degree+1 values to out[0..degree].
template <class T>
void get_basis_function_values(const Dim<T>& d, T t, int64_t idx, int derivative_order, T* out) {
    const int p = d.degree;
    if (derivative_order > p) {                                   // :79-81
        for (int k = 0; k <= p; ++k) out[k] = T(0);
        return;
    }
    T b[MAXDEG + 4], nb[MAXDEG + 4];                              // 1-based, length p+2 (+1 guard)
    for (int k = 1; k <= p + 2; ++k) b[k] = (k == p + 1) ? T(1) : T(0);   // :86-89
    b[p + 3] = T(0);
    const T* K = d.knots_all.data() - 1;                          // 1-based view
    for (int dd = 1; dd <= p; ++dd) {                             // :104-112
        const bool deriv = dd > p - derivative_order;             // :105
        for (int k = 1; k <= p + 2; ++k) nb[k] = cox_de_boor(b, K, t, idx, p, dd, k, deriv);
        for (int k = 1; k <= p + 2; ++k) b[k] = nb[k];
    }
    for (int k = 1; k <= p + 1; ++k) out[k - 1] = b[k];           // :97
}

// The interpolation object (src/DataInterpolationsND.jl:29-51).
template <class T>
struct Interp {
    int n_in = 0;
    Dim<T> dims[MAXN];
    const T* u = nullptr;
    int64_t u_stride[MAXN];    // column-major strides of the first n_in dims
    int64_t comp_stride = 0;   // prod(size(u)[1:n_in])
    int64_t n_comp = 1;        // prod(size(u)[n_in+1:end])
    bool scalar_out = true;    // N_out == 0
    const T* weights = nullptr;  // NURBSWeights (spline_utils.jl:232-234) or null
    const T* basis_cache[MAXN] = {nullptr};  // optional basis_function_eval caches (unused: recomputed)
};

template <class T>
inline int64_t u_offset(const Interp<T>& A, const int64_t* J /*1-based*/) {
    int64_t off = 0;
    for (int i = 0; i < A.n_in; ++i) off += (J[i] - 1) * A.u_stride[i];
    return off;
}

// Status of one point evaluation: whether `out` was written.
enum PointStatus { WRITTEN = 0, UNTOUCHED = 1 };

// src/interpolation_methods.jl:1-39 — Linear.
// out[c * out_stride] for c in 0..n_comp-1 is the view of the output for this point.
template <class T>
PointStatus interpolate_linear(T* out, int64_t out_stride, const Interp<T>& A, const T* t,
                               const int64_t* idx, const int* orders) {
    const int N = A.n_in;
    for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] = T(0);   // :9 make_zero!!
    for (int i = 0; i < N; ++i) if (orders[i] > 1) return WRITTEN;       // :10
    T t1[MAXN], t2[MAXN];
    for (int i = 0; i < N; ++i) { t1[i] = A.dims[i].t[idx[i] - 1]; t2[i] = A.dims[i].t[idx[i]]; }  // :12-13
    T t_vol = T(1);                                                      // :16-19
    for (int i = 0; i < N; ++i) t_vol *= t2[i] - t1[i];
    const int64_t ncorner = (int64_t)1 << N;
    for (int64_t corner = 0; corner < ncorner; ++corner) {               // :22 (dim 1 fastest)
        T c = T(1) / t_vol;                                              // :23 inv(t_vol)
        int64_t J[MAXN];
        for (int i = 0; i < N; ++i) {                                    // :24-30
            const bool right = (corner >> i) & 1;
            if (right) c *= (orders[i] == 0) ? (t[i] - t1[i]) : T(1);
            else       c *= (orders[i] == 0) ? (t2[i] - t[i]) : -T(1);
            J[i] = idx[i] + (right ? 1 : 0);                             // :31
        }
        const int64_t off = u_offset(A, J);
        for (int64_t cc = 0; cc < A.n_comp; ++cc)                        // :32-36
            out[cc * out_stride] += c * A.u[off + cc * A.comp_stride];
    }
    return WRITTEN;
}

// src/interpolation_methods.jl:41-66 — Constant (the struct's `left` field is never read).
template <class T>
PointStatus interpolate_constant(T* out, int64_t out_stride, const Interp<T>& A, const T* t,
                                 const int64_t* idx, const int* orders) {
    const int N = A.n_in;
    bool any_deriv = false;
    for (int i = 0; i < N; ++i) any_deriv |= orders[i] > 0;
    if (any_deriv) {                                                     // :49-55
        bool on_knot = false;
        for (int i = 0; i < N; ++i) {
            const Dim<T>& d = A.dims[i];
            // !isempty(searchsorted(t, x))  <=>  searchsortedfirst <= searchsortedlast
            if (searchsortedfirst(d.t, d.n_t, t[i]) <= searchsortedlast(d.t, d.n_t, t[i])) on_knot = true;
        }
        if (on_knot) {                                                   // typed_nan, interpolation_utils.jl:163-172
            for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] = std::numeric_limits<T>::quiet_NaN();
            return WRITTEN;
        }
        // `out` is returned unchanged: zero for the scalar case (fresh make_out,
        // interpolation_parallel.jl:129-131), untouched memory for N_out > 0 views.
        if (A.scalar_out) { out[0] = T(0); return WRITTEN; }
        return UNTOUCHED;
    }
    int64_t J[MAXN];
    for (int i = 0; i < N; ++i) {                                        // :57-59
        const Dim<T>& d = A.dims[i];
        J[i] = (t[i] >= d.t[d.n_t - 1]) ? d.n_t : idx[i];
    }
    const int64_t off = u_offset(A, J);
    for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] = A.u[off + c * A.comp_stride];  // :60-64
    return WRITTEN;
}

// src/interpolation_methods.jl:69-98 (BSpline) and :101-141 (NURBS, when A.weights != null).
template <class T>
PointStatus interpolate_bspline(T* out, int64_t out_stride, const Interp<T>& A, const T* t,
                                const int64_t* idx, const int* orders, const T* const* cached_basis) {
    const int N = A.n_in;
    for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] = T(0);   // :79 / :112
    T B[MAXN][MAXDEG + 1];
    int deg[MAXN];
    for (int i = 0; i < N; ++i) {                                        // :80-83 (spline_utils.jl:149-162)
        deg[i] = A.dims[i].degree;
        if (cached_basis && cached_basis[i]) {                           // view(basis_function_eval, k, :, r+1), spline_utils.jl:118-146
            for (int k = 0; k <= deg[i]; ++k) B[i][k] = cached_basis[i][k];
        } else {                                                         // on the fly, spline_utils.jl:64-98
            get_basis_function_values(A.dims[i], t[i], idx[i], orders[i], B[i]);
        }
    }
    T denom = T(0);                                                      // :117
    int I[MAXN];
    for (int i = 0; i < N; ++i) I[i] = 1;
    for (;;) {                                                           // :85 CartesianIndices, dim 1 fastest
        T B_product = B[0][I[0] - 1];                                    // :86 prod(...) = left fold
        for (int i = 1; i < N; ++i) B_product = B_product * B[i][I[i] - 1];
        int64_t cp[MAXN];
        for (int i = 0; i < N; ++i) cp[i] = idx[i] + I[i] - deg[i] - 1;  // :87-89
        const int64_t off = u_offset(A, cp);
        if (A.weights) {                                                 // :124-131
            const T product = A.weights[off] * B_product;
            denom += product;
            for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] += product * A.u[off + c * A.comp_stride];
        } else {                                                         // :90-94
            for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] += B_product * A.u[off + c * A.comp_stride];
        }
        int i = 0;
        for (; i < N; ++i) { if (++I[i] <= deg[i] + 1) break; I[i] = 1; }
        if (i == N) break;
    }
    if (A.weights)                                                       // :134-138
        for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] /= denom;
    return WRITTEN;
}

// ---------------------------------------------------------------------------
// EXTENSION WITHOUT REFERENCE SEMANTICS: mixed kinds in one
// NDInterpolation.  The reference cannot evaluate these (interp_dims is a homogeneous
// NTuple, src/DataInterpolationsND.jl:36).  Defined as the tensor product of the
// per-dimension stencils of the homogeneous evaluators above:
//   Constant -> 1 point (index rule of interpolation_methods.jl:57-59), weight 1;
//   Linear   -> 2 points, weights (t2-x, x-t1)/(t2-t1) or (-1, 1)/(t2-t1) (order 1), 0 for order > 1;
//   BSpline  -> degree+1 points, Cox–de Boor values.
// A Constant dimension with derivative order > 0 follows the Constant rule (NaN when any
// Constant dimension sits on a knot, else zero / untouched).  NURBS weights multiply in as in :124-138.
// ---------------------------------------------------------------------------
template <class T>
PointStatus interpolate_mixed(T* out, int64_t out_stride, const Interp<T>& A, const T* t,
                              const int64_t* idx, const int* orders, const T* const* cached_basis) {
    const int N = A.n_in;
    bool const_deriv = false, on_knot = false;
    for (int i = 0; i < N; ++i) {
        const Dim<T>& d = A.dims[i];
        if (d.kind == CONSTANT) {
            if (orders[i] > 0) const_deriv = true;
            // like :50 — every Constant dimension is tested, not only the differentiated one
            if (searchsortedfirst(d.t, d.n_t, t[i]) <= searchsortedlast(d.t, d.n_t, t[i])) on_knot = true;
        }
    }
    if (const_deriv) {
        if (on_knot) {
            for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] = std::numeric_limits<T>::quiet_NaN();
            return WRITTEN;
        }
        if (A.scalar_out) { out[0] = T(0); return WRITTEN; }
        return UNTOUCHED;
    }
    for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] = T(0);
    T W[MAXN][MAXDEG + 1];
    int64_t first[MAXN];
    int width[MAXN];
    for (int i = 0; i < N; ++i) {
        const Dim<T>& d = A.dims[i];
        if (d.kind == CONSTANT) {
            width[i] = 1; W[i][0] = T(1);
            first[i] = (t[i] >= d.t[d.n_t - 1]) ? d.n_t : idx[i];
        } else if (d.kind == LINEAR) {
            width[i] = 2; first[i] = idx[i];
            const T t1 = d.t[idx[i] - 1], t2 = d.t[idx[i]];
            const T inv = T(1) / (t2 - t1);
            if (orders[i] == 0) { W[i][0] = inv * (t2 - t[i]); W[i][1] = inv * (t[i] - t1); }
            else if (orders[i] == 1) { W[i][0] = -inv; W[i][1] = inv; }
            else { W[i][0] = T(0); W[i][1] = T(0); }
        } else {
            width[i] = d.degree + 1; first[i] = idx[i] - d.degree;
            if (cached_basis && cached_basis[i]) for (int k = 0; k <= d.degree; ++k) W[i][k] = cached_basis[i][k];
            else get_basis_function_values(d, t[i], idx[i], orders[i], W[i]);
        }
    }
    T denom = T(0);
    int I[MAXN];
    for (int i = 0; i < N; ++i) I[i] = 0;
    for (;;) {
        T w = W[0][I[0]];
        for (int i = 1; i < N; ++i) w = w * W[i][I[i]];
        int64_t cp[MAXN];
        for (int i = 0; i < N; ++i) cp[i] = first[i] + I[i];
        const int64_t off = u_offset(A, cp);
        if (A.weights) { w = A.weights[off] * w; denom += w; }
        for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] += w * A.u[off + c * A.comp_stride];
        int i = 0;
        for (; i < N; ++i) { if (++I[i] < width[i]) break; I[i] = 0; }
        if (i == N) break;
    }
    if (A.weights)
        for (int64_t c = 0; c < A.n_comp; ++c) out[c * out_stride] /= denom;
    return WRITTEN;
}

template <class T>
inline PointStatus interpolate(T* out, int64_t out_stride, const Interp<T>& A, const T* t,
                               const int64_t* idx, const int* orders, int homogeneous_kind,
                               const T* const* cached_basis) {
    switch (homogeneous_kind) {
        case LINEAR: return interpolate_linear(out, out_stride, A, t, idx, orders);
        case CONSTANT: return interpolate_constant(out, out_stride, A, t, idx, orders);
        case BSPLINE: return interpolate_bspline(out, out_stride, A, t, idx, orders, cached_basis);
        default: return interpolate_mixed(out, out_stride, A, t, idx, orders, cached_basis);
    }
}

// src/interpolation_parallel.jl:105-138 — eval_kernel, one "work item" per output point;
// the KA CPU backend's static split of ndrange over threads is mirrored by
// `omp for schedule(static)`.
double g_last_eval_seconds = 0.0;   // duration of the last eval_kernel loop (caches excluded)

template <class T>
void eval_driver(T* out, const Interp<T>& A, const int* orders, bool grid, int kind, int nthreads) {
    const int N = A.n_in;
    int64_t npts = 1;
    if (grid) for (int i = 0; i < N; ++i) npts *= A.dims[i].n_eval;
    else npts = A.dims[0].n_eval;
    const int nt =
#ifdef _OPENMP
        nthreads > 0 ? nthreads : omp_get_max_threads();
#else
        1;
    (void)nthreads;
#endif
    // The per-dimension caches the reference fills at CONSTRUCTION time, outside eval_*:
    // idx_eval (set_eval_idx!, interpolation_utils.jl:143-161) and, for BSpline dimensions,
    // basis_function_eval for the requested derivative order (set_basis_function_eval!,
    // spline_utils.jl:164-191; the reference fills orders 0..max, only the one read is built here).
    std::vector<int64_t> idx_cache[MAXN];
    std::vector<T> basis_cache[MAXN];
    for (int i = 0; i < N; ++i) {
        const Dim<T>& d = A.dims[i];
        idx_cache[i].resize((size_t)d.n_eval);
        const int W = d.degree + 1;
        if (d.kind == BSPLINE) basis_cache[i].resize((size_t)d.n_eval * W);
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
        for (int64_t k = 0; k < d.n_eval; ++k) {
            idx_cache[i][k] = get_idx(d, d.t_eval[k]);
            if (d.kind == BSPLINE) get_basis_function_values(d, d.t_eval[k], idx_cache[i][k], orders[i], &basis_cache[i][k * W]);
        }
    }
#ifdef _OPENMP
    const double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
    for (int64_t k = 0; k < npts; ++k) {
        T t[MAXN];
        int64_t idx[MAXN];
        const T* bc[MAXN];
        int64_t rem = k;
        for (int i = 0; i < N; ++i) {
            int64_t ki = k;
            if (grid) { ki = rem % A.dims[i].n_eval; rem /= A.dims[i].n_eval; }
            t[i] = A.dims[i].t_eval[ki];                                 // :117 / :120
            idx[i] = idx_cache[i][ki];                                   // :118 / :121
            bc[i] = A.dims[i].kind == BSPLINE ? &basis_cache[i][ki * (A.dims[i].degree + 1)] : nullptr;
        }
        interpolate(out + k, npts, A, t, idx, orders, kind, bc);         // :128-137
    }
#ifdef _OPENMP
    g_last_eval_seconds = omp_get_wtime() - t0;
#endif
}

// --- plumbing from the flat C arguments to Interp<T> -----------------------
template <class T>
int build_interp(Interp<T>& A, int n_in, const int* kinds, const void* const* t, const int64_t* n_t,
                 const int* degrees, const int64_t* const* mults, const void* u, const int64_t* u_dims,
                 int u_ndims, const void* weights, const void* const* t_eval, const int64_t* n_eval) {
    if (n_in < 1 || n_in > MAXN || u_ndims < n_in) return -1;
    A.n_in = n_in;
    int64_t stride = 1;
    for (int i = 0; i < n_in; ++i) {
        Dim<T>& d = A.dims[i];
        d.kind = kinds[i];
        d.degree = degrees ? degrees[i] : 0;
        d.t = (const T*)t[i];
        d.n_t = n_t[i];
        if (d.degree > MAXDEG) return -1;
        for (int64_t j = 1; j < d.n_t; ++j) if (!(d.t[j] - d.t[j - 1] > T(0))) return -2;  // validate_t, interpolation_utils.jl:34-37
        if (d.kind == BSPLINE) {
            std::vector<int64_t> m;
            if (mults && mults[i]) m.assign(mults[i], mults[i] + d.n_t);
            else {                                                      // interpolation_dimensions.jl:168-173
                m.assign((size_t)d.n_t, 1);
                m[0] = d.degree + 1; m[(size_t)d.n_t - 1] = d.degree + 1;
            }
            for (int64_t j = 0; j < d.n_t; ++j) if (m[j] < 1 || m[j] > d.degree + 1) return -3;  // :175
            expand_knot_vector(d.knots_all, d.t, m.data(), d.n_t);      // :177-185
        }
        d.t_eval = t_eval ? (const T*)t_eval[i] : nullptr;
        d.n_eval = n_eval ? n_eval[i] : 0;
        if (u_dims[i] != d.n_u()) return -4;                            // validate_size_u
        A.u_stride[i] = stride;
        stride *= u_dims[i];
    }
    A.comp_stride = stride;
    A.n_comp = 1;
    for (int i = n_in; i < u_ndims; ++i) A.n_comp *= u_dims[i];
    A.scalar_out = (u_ndims == n_in);
    A.u = (const T*)u;
    A.weights = (const T*)weights;
    return 0;
}

inline int homogeneous_kind(int n_in, const int* kinds) {
    for (int i = 1; i < n_in; ++i) if (kinds[i] != kinds[0]) return -1;
    return kinds[0];
}

template <class T>
int eval_T(int n_in, const int* kinds, const void* const* t, const int64_t* n_t, const int* degrees,
           const int64_t* const* mults, const void* u, const int64_t* u_dims, int u_ndims,
           const void* weights, const void* const* t_eval, const int64_t* n_eval, const int* orders,
           int grid, void* out, int nthreads) {
    Interp<T> A;
    int rc = build_interp<T>(A, n_in, kinds, t, n_t, degrees, mults, u, u_dims, u_ndims, weights, t_eval, n_eval);
    if (rc) return rc;
    for (int i = 0; i < n_in; ++i) if (orders[i] < 0) return -5;        // interpolation_utils.jl:5-12
    if (weights) for (int i = 0; i < n_in; ++i) if (orders[i] != 0) return -6;  // :29-31
    if (!grid) for (int i = 1; i < n_in; ++i) if (n_eval[i] != n_eval[0]) return -7;
    eval_driver<T>((T*)out, A, orders, grid != 0, homogeneous_kind(n_in, kinds), nthreads);
    return 0;
}

template <class T>
int idx_T(int kind, const void* t, int64_t n_t, int degree, const int64_t* mult, const void* t_eval,
          int64_t n_eval, int64_t* idx_out) {
    Dim<T> d;
    d.kind = kind; d.degree = degree; d.t = (const T*)t; d.n_t = n_t;
    if (kind == BSPLINE) {
        std::vector<int64_t> m;
        if (mult) m.assign(mult, mult + n_t);
        else { m.assign((size_t)n_t, 1); m[0] = degree + 1; m[(size_t)n_t - 1] = degree + 1; }
        expand_knot_vector(d.knots_all, d.t, m.data(), n_t);
    }
    const T* te = (const T*)t_eval;
    for (int64_t i = 0; i < n_eval; ++i) idx_out[i] = get_idx(d, te[i]);   // set_idx_kernel, interpolation_utils.jl:156-161
    return 0;
}

// basis_function_eval cache (src/spline_utils.jl:164-191): out is (n_eval, degree+1, max_deriv+1), column-major.
template <class T>
int basis_T(const void* t, int64_t n_t, int degree, const int64_t* mult, const void* t_eval,
            int64_t n_eval, int max_deriv, void* out) {
    Dim<T> d;
    d.kind = BSPLINE; d.degree = degree; d.t = (const T*)t; d.n_t = n_t;
    std::vector<int64_t> m;
    if (mult) m.assign(mult, mult + n_t);
    else { m.assign((size_t)n_t, 1); m[0] = degree + 1; m[(size_t)n_t - 1] = degree + 1; }
    expand_knot_vector(d.knots_all, d.t, m.data(), n_t);
    const T* te = (const T*)t_eval;
    T* o = (T*)out;
    T vals[MAXDEG + 1];
    for (int r = 0; r <= max_deriv; ++r)
        for (int64_t i = 0; i < n_eval; ++i) {
            get_basis_function_values(d, te[i], get_idx(d, te[i]), r, vals);
            for (int k = 0; k <= degree; ++k) o[i + n_eval * (k + (int64_t)(degree + 1) * r)] = vals[k];
        }
    return 0;
}

template <class T>
int knots_T(const void* t, int64_t n_t, int degree, const int64_t* mult, void* out, int64_t* n_out) {
    std::vector<int64_t> m;
    if (mult) m.assign(mult, mult + n_t);
    else { m.assign((size_t)n_t, 1); m[0] = degree + 1; m[(size_t)n_t - 1] = degree + 1; }
    std::vector<T> ka;
    expand_knot_vector(ka, (const T*)t, m.data(), n_t);
    *n_out = (int64_t)ka.size();
    if (out) std::memcpy(out, ka.data(), ka.size() * sizeof(T));
    return 0;
}

}  // namespace

extern "C" {

// dtype: 0 = Float32, 1 = Float64.  Returns 0 on success, negative on validation failure.
int dindo_eval(int dtype, int n_in, const int* kinds, const void* const* t, const int64_t* n_t,
               const int* degrees, const int64_t* const* mults, const void* u, const int64_t* u_dims,
               int u_ndims, const void* weights, const void* const* t_eval, const int64_t* n_eval,
               const int* orders, int grid, void* out, int nthreads) {
    if (dtype == 0)
        return eval_T<float>(n_in, kinds, t, n_t, degrees, mults, u, u_dims, u_ndims, weights, t_eval, n_eval, orders, grid, out, nthreads);
    return eval_T<double>(n_in, kinds, t, n_t, degrees, mults, u, u_dims, u_ndims, weights, t_eval, n_eval, orders, grid, out, nthreads);
}

int dindo_get_idx(int dtype, int kind, const void* t, int64_t n_t, int degree, const int64_t* mult,
                  const void* t_eval, int64_t n_eval, int64_t* idx_out) {
    if (dtype == 0) return idx_T<float>(kind, t, n_t, degree, mult, t_eval, n_eval, idx_out);
    return idx_T<double>(kind, t, n_t, degree, mult, t_eval, n_eval, idx_out);
}

int dindo_basis(int dtype, const void* t, int64_t n_t, int degree, const int64_t* mult,
                const void* t_eval, int64_t n_eval, int max_deriv, void* out) {
    if (dtype == 0) return basis_T<float>(t, n_t, degree, mult, t_eval, n_eval, max_deriv, out);
    return basis_T<double>(t, n_t, degree, mult, t_eval, n_eval, max_deriv, out);
}

int dindo_expand_knots(int dtype, const void* t, int64_t n_t, int degree, const int64_t* mult,
                       void* out, int64_t* n_out) {
    if (dtype == 0) return knots_T<float>(t, n_t, degree, mult, out, n_out);
    return knots_T<double>(t, n_t, degree, mult, out, n_out);
}

// seconds spent in the last evaluation loop proper (the reference's eval_* call: caches are
// built by the constructors and are not part of it)
double dindo_last_eval_seconds(void) { return g_last_eval_seconds; }

int dindo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
