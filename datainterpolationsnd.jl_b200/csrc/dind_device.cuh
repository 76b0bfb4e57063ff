// dind_device.cuh — device-side building blocks shared by all kernels of libdind.so.
//
// Restates, for the GPU, the per-point math of the reference:
//   interval search  : src/interpolation_utils.jl:104-134 (get_idx) on Base.searchsortedfirst/last
//   Cox–de Boor basis: src/spline_utils.jl:22-114
//   stencils         : src/interpolation_methods.jl:1-141
// B200 notes: everything here is register/shared-memory math feeding HBM/L2-bound gathers;
// no tensor-core work (per-axis bands are 2..8 wide).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace dind {

constexpr int MAXN = 8;  // DIND_MAX_DIMS
constexpr int MAXW = 8;  // DIND_MAX_DEGREE + 1

enum Kind : int { LINEAR = 0, CONSTANT = 1, BSPLINE = 2 };

// ---- Julia `isless` total order via sortable integer keys -------------------------------
// key(a) < key(b)  <=>  isless(a, b):  -0.0 < +0.0, every NaN is the greatest element.
template <typename T> struct KeyOf;
template <> struct KeyOf<float> { using type = int32_t; static constexpr int32_t nan = 0x7fffffff; };
template <> struct KeyOf<double> { using type = int64_t; static constexpr int64_t nan = 0x7fffffffffffffffLL; };

__host__ __device__ __forceinline__ int32_t to_key(float v) {
#ifdef __CUDA_ARCH__
    int32_t b = __float_as_int(v);
#else
    int32_t b; memcpy(&b, &v, 4);
#endif
    if ((b & 0x7fffffff) > 0x7f800000) return 0x7fffffff;  // NaN (either sign) sorts last
    return b ^ ((b >> 31) & 0x7fffffff);
}
__host__ __device__ __forceinline__ int64_t to_key(double v) {
#ifdef __CUDA_ARCH__
    int64_t b = __double_as_longlong(v);
#else
    int64_t b; memcpy(&b, &v, 8);
#endif
    if ((b & 0x7fffffffffffffffLL) > 0x7ff0000000000000LL) return 0x7fffffffffffffffLL;
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

// Number of keys <= kx (searchsortedlast) / < kx (searchsortedfirst - 1) in a sorted array.
// Branch-free binary lifting: ceil(log2(n+1)) predicated steps, no divergence.
template <typename K>
__device__ __forceinline__ int count_le(const K* __restrict__ keys, int n, K kx) {
    int lo = 0;
    int step = 1 << (31 - __clz(n | 1));
    for (; step > 0; step >>= 1) {
        int j = lo + step;
        if (j <= n && keys[j - 1] <= kx) lo = j;
    }
    return lo;
}
template <typename K>
__device__ __forceinline__ int count_lt(const K* __restrict__ keys, int n, K kx) {
    int lo = 0;
    int step = 1 << (31 - __clz(n | 1));
    for (; step > 0; step >>= 1) {
        int j = lo + step;
        if (j <= n && keys[j - 1] < kx) lo = j;
    }
    return lo;
}

__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x > hi ? hi : (x < lo ? lo : x); }

// ---- bucket-accelerated search ---------------------------------------------------------------
// lut_bucket is monotone non-decreasing in x (subtraction of a constant, multiplication by a
// positive constant, clamping and truncation all are), so for a query in bucket b every knot in
// a lower bucket is < x and every knot in a higher bucket is > x: the exact count — still decided
// by key comparisons only, hence bit-exact — lies in [lut[b], lut[b+1]], typically 0-2 knots.
// The table is built on the device with this same function (dind_generic.cu: lut_kernel).
template <typename T>
__device__ __forceinline__ int lut_bucket(T x, T t0, T scale, int n_buckets) {
    const T v = fmin(fmax((x - t0) * scale, T(0)), T(n_buckets - 1));   // NaN -> 0
    return (int)v;
}

template <typename K, bool LE>
__device__ __forceinline__ int count_range(const K* __restrict__ keys, int lo, int hi, K kx) {
    const int len = hi - lo;
    if (len <= 2) {
        // the usual case (~4 buckets per knot), branch-free: the keys are sorted, so the count over the one or
        // two candidates is a prefix of passed comparisons.  lo < n whenever len > 0.
        const int j0 = max(lo - (int)(len <= 0), 0);   // any valid address when there is no candidate (lo may be n)
        const int j1 = len == 2 ? lo + 1 : j0;
        const K k0 = __ldg(keys + j0);
        const K k1 = __ldg(keys + j1);
        const bool p0 = len >= 1 && (LE ? k0 <= kx : k0 < kx);
        const bool p1 = len == 2 && (LE ? k1 <= kx : k1 < kx);
        return lo + (int)p0 + (int)(p0 && p1);
    }
    int cnt = lo;
    for (int step = 1 << (31 - __clz(len)); step > 0; step >>= 1) {
        const int j = cnt + step;
        if (j <= hi && (LE ? keys[j - 1] <= kx : keys[j - 1] < kx)) cnt = j;
    }
    return cnt;
}

// ---- per-dimension device descriptor ------------------------------------------------------
template <typename T>
struct DimParams {
    using K = typename KeyOf<T>::type;
    const T* knots;        // t (Linear/Constant) or knots_all (BSpline)
    const K* keys;         // to_key(knots)
    const T* recip;        // Linear: 1/(t[c+1]-t[c]), n_knots-1 entries
                           // BSpline: R_d[i] = safe 1/(K[i+d]-K[i]), layout [(d-1)*n_knots + i]
    const T* t_eval;       // unstructured: per point; grid: per axis
    const int32_t* idx_eval;  // optional cache of the reference's 1-based idx_eval
    // per-axis tables for grid evaluation (built for the requested derivative order)
    const int32_t* tab_i0;    // first stencil node (0-based) per grid coordinate
    const T* tab_w;           // weights [g * width + a]
    const uint8_t* tab_flag;  // bit0: Constant-with-derivative on a knot
    int64_t n_eval;
    int64_t u_stride;      // element stride of this dimension in u
    int64_t out_stride;    // grid evaluation: element stride of this axis in out
    int32_t n_knots;
    int32_t kind;
    int32_t degree;
    int32_t width;         // stencil width: 1 Constant, 2 Linear, degree+1 BSpline
    int32_t n_u;           // size(u, dim)
    int32_t order;         // derivative order requested for this dimension
    // bucket table accelerating the interval search (null = plain binary search): for bucket
    // b = lut_bucket(x), count_lt(x) and count_le(x) lie in [lut[b].x, lut[b].y] (.y = the next bucket's .x;
    // stored as pairs so that a query is ONE 8-byte load)
    const int2* lut;
    T lut_t0, lut_scale;
    int32_t lut_n;
    // cell-sorted plan with packed cells (EvalParams::cells): this dimension's first stencil node sits in bits
    // [cell_shift, cell_shift + popc(cell_mask)) of the point's cell word; cell_add turns it into the reference's
    // 1-based idx (Linear: + 1, BSpline: + degree + 1)
    int32_t cell_shift, cell_mask, cell_add;
    // cell-record copy of u (EvalParams::u_rec): stride of this dimension's first node, in records
    int64_t rec_stride;
};

// Experiment knobs (DIND_* environment variables; not part of the API).  Read from the environment ONCE per
// handle, in dind_create, and carried to the launchers by pointer: the evaluation path never calls getenv.
struct Knobs {
    int chunk = 0;       // DIND_CHUNK: pencil chunk length (0 = chosen by the launcher)
    int pf = 1;          // DIND_PF: one-run-ahead register prefetch of the pencil kernel
    int pf1 = -1;        // DIND_PF1: L1 prefetch distance (-1 = default of the instantiation)
    int l2_warm = -1;    // DIND_L2_WARM: bulk L2 warm-up of u (-1 = default)
    int c3_tile = -1;    // DIND_C3_TILE: points per work item of the tile kernels (0 = one-point kernels, -1 = default)
    bool no_lut = false, no_split = false, no_cells = false, no_aos = false, no_small = false, no_slice = false, no_plan_recs = false, no_batch_vec = false, no_den_cache = false;
    int tile_smem = -1;      // DIND_TILE_SMEM: cubic tile kernel with the stencils staged in shared memory (1 / 0; -1 = where the plan is dense)
    int plan_buckets = -1;   // DIND_PLAN_BUCKETS: un-permutation through window-ordered slots (1) or original-order staging (0); -1 = by record size
    bool persist = false;      // DIND_PERSIST=1: the persistent two-stage 4-D grid kernel instead of two launches
    int rec = -1;        // DIND_REC: cell-record copy of u for multilinear tables at iid points (0 = never, 1 = always, -1 = by size)
};

template <typename T>
struct EvalParams {
    const Knobs* knobs;    // host pointer (never dereferenced on the device); null = defaults
    DimParams<T> dim[MAXN];
    const T* u;            // (n_1..n_N, comps): NURBS: pre-multiplied by the weights
    const T* wts;          // NURBS weights (n_1..n_N) or null
    const T* u_aos;        // optional component-fastest copy of u: u_aos[c + n_comp * i] (unstructured, 2..4 components)
    const T* u_rec;        // optional cell-record copy of u_aos (dind_fast_unstructured_rec.cu): all 2^N corners of a cell contiguous
    T* out;                // (n_points, comps)
    int64_t n_points;
    int64_t u_comp_stride;
    int64_t out_comp_stride;
    int64_t out_point_stride;  // 1 for the reference layout (points fastest); n_comp for the plan's component-fastest staging
    int32_t n_comp;
    int32_t n_in;
    int32_t scalar_out;    // N_out == 0
    int32_t const_deriv;   // some Constant dimension has derivative order > 0
    int32_t zero_result;   // Linear with order > 1 somewhere: result is identically zero
    int32_t use_cached_idx;
    const uint32_t* perm;  // cell-sorted plan: output position of point k of the (sorted) t_eval arrays; null = identity
    const uint32_t* cells; // cell-sorted plan: per plan position the packed first stencil nodes of all dimensions (the
                           // plan's sort key, kept): the kernels decode the interval instead of searching; null = search
    int32_t n_stencil;       // grid: > 0 = dim[0..n_stencil) are the stencil axes and the rest identity/selector ("point") axes,
                             // already in kernel order (dind_fast_grid_split.cu); 0 = let the dispatcher classify the axes
    int32_t u_is_raw;        // NURBS: u has NOT been pre-multiplied by the weights (only the split grid path accepts that)
    int32_t split_den_cached; // NURBS, split grid path: the denominator's stage-1 result is already in the scratch (same weights, knots, t_eval)
    int32_t coherent_points; // neighbouring points share their stencil (cell-sorted plan): prefer narrow, broadcast loads
    // cell-sorted plan, work items (dind_plan.cu: plan_build_items): item i = plan positions [items[i], items[i + 1]) —
    // at most item_pts consecutive points of ONE cell, evaluated by one thread that loads every stencil node once
    const uint32_t* items;
    int64_t n_items;
    int32_t item_pts;
    int32_t staged_records;  // out is the plan's staging buffer: component-fastest records of out_point_stride
                             // elements, padded and aligned as plan_record_stride() says (vector stores allowed)
};

// Elements per staging record for ct result values of esz bytes: records of up to 32 bytes are padded to
// 8 / 16 / 32 bytes so that one record is one aligned vector store and never straddles a 32-byte sector.
__host__ __device__ inline int plan_record_stride(int ct, int esz) {
    const int bytes = ct * esz;
    if (bytes <= 8) return ct;
    if (bytes <= 16) return 16 / esz;
    if (bytes <= 32) return 32 / esz;
    return ((bytes + 15) / 16) * (16 / esz);
}

// count of knots < x (LE = false: searchsortedfirst - 1) or <= x (LE = true: searchsortedlast) in
// Julia's isless order, through the bucket table when the dimension has one.
template <typename T, bool LE>
__device__ __forceinline__ int dim_count(const DimParams<T>& D, T x, typename KeyOf<T>::type kx) {
    using K = typename KeyOf<T>::type;
    const int n = D.n_knots;
    if (D.lut == nullptr) return LE ? count_le(D.keys, n, kx) : count_lt(D.keys, n, kx);
    if (kx == KeyOf<T>::nan) return n;   // NaN sorts after every knot
    const int2 r = __ldg(D.lut + lut_bucket(x, D.lut_t0, D.lut_scale, D.lut_n));
    return count_range<K, LE>(D.keys, r.x, r.y, kx);
}

// the reference's 1-based idx of dimension D from the plan's packed cell word
template <typename T>
__device__ __forceinline__ int cell_idx(const DimParams<T>& D, uint32_t cell) {
    return (int)((cell >> D.cell_shift) & (uint32_t)D.cell_mask) + D.cell_add;
}

// ---- Cox–de Boor in registers (src/spline_utils.jl:24-114), runtime degree ---------------
// idx0: 0-based knot interval (reference idx - 1).  Writes p+1 values to w[0..p].
// Division by the knot spans is replaced by the precomputed safe reciprocals R_d[i]
// (safe_div(a, 0) = 0 of spline_utils.jl:22 <=> R = 0).
template <typename T>
__device__ __forceinline__ void bspline_basis_rt(const T* __restrict__ Kn, const T* __restrict__ R, int nk,
                                                 int p, int idx0, T x, int order, T* w) {
    if (order > p) {                       // :79-81
        for (int s = 0; s <= p; ++s) w[s] = T(0);
        return;
    }
    T b[MAXW + 1];
    for (int s = 0; s <= p + 1; ++s) b[s] = (s == p) ? T(1) : T(0);   // :86-89
    for (int d = 1; d <= p; ++d) {         // :104-112
        const bool deriv = d > p - order;  // :105
        const T* Rd = R + (int64_t)(d - 1) * nk;
        for (int s = p - d; s <= p; ++s) { // slots below p-d stay zero (:33-34)
            const int i = idx0 - p + s;
            T v;
            if (deriv) v = T(d) * (b[s] * Rd[i] - b[s + 1] * Rd[i + 1]);               // :38-41
            else v = b[s] * ((x - Kn[i]) * Rd[i]) + b[s + 1] * ((Kn[i + d + 1] - x) * Rd[i + 1]);  // :45-46
            b[s] = v;                      // in place: b[s+1] is still the previous stage's value
        }
    }
    for (int s = 0; s <= p; ++s) w[s] = b[s];   // :97
}

// Round-to-nearest operations the compiler may NOT contract into FMAs: the same point must get the same bits from
// every kernel variant that evaluates it (one point or several per thread, derivative order known at compile time or
// not), which "a * b + c" would leave to each variant's own contraction choices.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// Compile-time degree version for the specialised kernels: fully unrolled, knots of the
// 2P+2 wide window are read once.
template <typename T, int P>
__device__ __forceinline__ void bspline_basis_ct(const T* __restrict__ Kn, const T* __restrict__ R, int nk,
                                                 int idx0, T x, int order, T (&w)[P + 1]) {
    if (order > P) {
#pragma unroll
        for (int s = 0; s <= P; ++s) w[s] = T(0);
        return;
    }
    T kn[2 * P + 2];                        // kn[j] = K[idx0 - P + j]
#pragma unroll
    for (int j = 0; j < 2 * P + 2; ++j) kn[j] = Kn[idx0 - P + j];
    T b[P + 2];
#pragma unroll
    for (int s = 0; s <= P + 1; ++s) b[s] = (s == P) ? T(1) : T(0);
#pragma unroll
    for (int d = 1; d <= P; ++d) {
        const bool deriv = d > P - order;
        const T* Rd = R + (d - 1) * nk + (idx0 - P);
#pragma unroll
        for (int s = P - d; s <= P; ++s) {
            // R_d at i = idx0-P+s and i+1; the lowest slot only needs the second term
            // (b[P-d] is zero on entry), the top slot only the first (b[P+1] == 0).
            T v;
            // explicit operations (mul_rn / sub_rn / fma): see mul_rn
            if (deriv) {
                if (s > P - d && s < P) v = mul_rn(T(d), fma(b[s], Rd[s], -mul_rn(b[s + 1], Rd[s + 1])));
                else if (s > P - d) v = mul_rn(T(d), mul_rn(b[s], Rd[s]));
                else v = mul_rn(T(d), -mul_rn(b[s + 1], Rd[s + 1]));
            } else {
                if (s > P - d && s < P)
                    v = fma(b[s], mul_rn(sub_rn(x, kn[s]), Rd[s]), mul_rn(b[s + 1], mul_rn(sub_rn(kn[s + d + 1], x), Rd[s + 1])));
                else if (s > P - d) v = mul_rn(b[s], mul_rn(sub_rn(x, kn[s]), Rd[s]));
                else v = mul_rn(b[s + 1], mul_rn(sub_rn(kn[s + d + 1], x), Rd[s + 1]));
            }
            b[s] = v;
        }
    }
#pragma unroll
    for (int s = 0; s <= P; ++s) w[s] = b[s];
}

// ---- one dimension's stencil: first node, weights, flags ------------------------------------
// cached_idx: the reference's 1-based idx_eval value, or 0 to search here.
// Returns the reference's 1-based idx through idx1 (for the idx cache kernels).
template <typename T>
__device__ __forceinline__ void dim_stencil(const DimParams<T>& D, T x, int order, int cached_idx,
                                            int& i0, T* w, bool& on_knot, int& idx1) {
    using K = typename KeyOf<T>::type;
    const K kx = to_key(x);
    const int n = D.n_knots;
    on_knot = false;
    if (D.kind == LINEAR) {
        // clamp(searchsortedfirst(t, x) - 1, 1, n-1): intervals (t[i], t[i+1]]
        const int idx = cached_idx ? cached_idx : clampi(dim_count<T, false>(D, x, kx), 1, n - 1);
        idx1 = idx;
        const int c = idx - 1;
        const T t1 = D.knots[c], t2 = D.knots[c + 1], r = D.recip[c];
        i0 = c;
        if (order == 0) { w[0] = (t2 - x) * r; w[1] = (x - t1) * r; }       // interpolation_methods.jl:24-30
        else if (order == 1) { w[0] = -r; w[1] = r; }
        else { w[0] = T(0); w[1] = T(0); }                                  // :10
    } else if (D.kind == CONSTANT) {
        // clamp(searchsortedlast(t, x), 1, n-1): intervals [t[i], t[i+1])
        const int cnt = dim_count<T, true>(D, x, kx);
        const int idx = cached_idx ? cached_idx : clampi(cnt, 1, n - 1);
        idx1 = idx;
        i0 = (x >= D.knots[n - 1]) ? n - 1 : idx - 1;                       // :57-59
        w[0] = T(1);
        // :50 !isempty(searchsorted(t, x)); only consulted when some Constant dimension has a
        // derivative order > 0 (the reference then tests every dimension, not only that one)
        on_knot = dim_count<T, false>(D, x, kx) < cnt;
    } else {
        const int p = D.degree;
        const int idx = cached_idx ? cached_idx : clampi(dim_count<T, true>(D, x, kx), p + 1, n - p - 1);
        idx1 = idx;
        i0 = idx - p - 1;                                                   // :87-89
        bspline_basis_rt<T>(D.knots, D.recip, n, p, idx - 1, x, order, w);
    }
}

// One staging record (see plan_record_stride) with the widest store it allows; 32-byte records leave as
// one full-sector 256-bit store (STG.256, sm_100).
template <typename T, int C>
__device__ __forceinline__ void store_record(T* dst, const T (&v)[C]) {
    constexpr int BYTES = C * (int)sizeof(T);
    constexpr int PER16 = 16 / (int)sizeof(T);
    if constexpr (sizeof(T) == 4 && C == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    } else if constexpr (sizeof(T) == 4 && BYTES <= 16) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], C > 3 ? v[C - 1] : 0.f);
    } else if constexpr (sizeof(T) == 8 && C == 2) {
        *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
    } else if constexpr (sizeof(T) == 8 && BYTES <= 32) {
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(v[0]), "d"(v[1]), "d"(v[2]),
                     "d"(C > 3 ? v[C - 1] : 0.0)
                     : "memory");
    } else if constexpr (sizeof(T) == 8 && BYTES % 32 == 0) {
#pragma unroll
        for (int i = 0; i < C; i += 4)
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "d"(v[i]), "d"(v[i + 1]), "d"(v[i + 2]),
                         "d"(v[i + 3])
                         : "memory");
    } else if constexpr (BYTES > 16) {   // 16-byte chunks, scalar tail (the record stride is a multiple of 16 bytes)
#pragma unroll
        for (int i = 0; i + PER16 <= C; i += PER16) {
            if constexpr (sizeof(T) == 8) *reinterpret_cast<double2*>(dst + i) = make_double2(v[i], v[i + 1]);
            else *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
#pragma unroll
        for (int i = C - C % PER16; i < C; ++i) dst[i] = v[i];
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = v[c];
    }
}

// ---- component-fastest ("AoS") copy of u: node i holds its n_comp values in aos_node_stride(n_comp)
// consecutive elements (dind_fast_unstructured_aos.cu, dind_fast_gradient.cu)
template <typename T, int C>
struct Vec {
    T v[C];
};

// all components of one node (aos_node_stride(C) elements, 8 / 16 / 32 bytes, aligned): ONE vector load —
// 32-byte nodes (3 or 4 doubles) through the 256-bit load of sm_100 (LDG.E.256)
// — unless the points are known to be warp-coherent (cell-sorted plan): lanes then read the SAME node, a
// 64-bit load of one address is one L1 wavefront and the 256-bit load is slower (measured, config 3).
template <typename T, int C, bool WIDE>
__device__ __forceinline__ Vec<T, C> load_node(const T* __restrict__ p) {
    constexpr int BYTES = (int)sizeof(T) * (C == 3 ? 4 : C);
    static_assert(BYTES == 8 || BYTES == 16 || BYTES == 32, "node size");
    Vec<T, C> r;
    if constexpr (BYTES == 32 && !WIDE) {   // (a 16-byte + 8-byte split is slower than three 8-byte loads here, measured)
#pragma unroll
        for (int c = 0; c < C; ++c) r.v[c] = __ldg(p + c);
    } else if constexpr (BYTES == 16) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p));
        memcpy(&r, &q, sizeof(T) * C);
    } else if constexpr (BYTES == 8) {
        const float2 q = __ldg(reinterpret_cast<const float2*>(p));
        memcpy(&r, &q, 8);
    } else {
        double q[4];
        asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(q[0]), "=d"(q[1]), "=d"(q[2]), "=d"(q[3]) : "l"(p));
#pragma unroll
        for (int c = 0; c < C; ++c) r.v[c] = (T)q[c];
    }
    return r;
}

// NURBS quotient num / den (src/interpolation_methods.jl:134-138) for the grid kernels, where it is paid per
// output point.  The IEEE FP64 division is a ~30-instruction routine (ncu: half of all instructions of the
// config-4 kernels); for a denominator in the normal range this is the hardware reciprocal seed, two Newton
// steps and one residual correction — within 1 ulp of the IEEE quotient.  Zero, subnormal, huge, Inf and
// NaN denominators take the IEEE division, so the special values are the reference's.
template <typename T>
__device__ __forceinline__ bool nurbs_den_is_normal(T den) {
    const T a = fabs(den);
    return sizeof(T) == 8 ? (a > T(1e-290) && a < T(1e290)) : (a > T(1e-30) && a < T(1e30));
}
// quotient for a denominator that passed nurbs_den_is_normal
template <typename T>
__device__ __forceinline__ T nurbs_div_normal(T num, T den) {
    return num / den;
}
template <>
__device__ __forceinline__ double nurbs_div_normal<double>(double num, double den) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    const double q = num * r;
    return fma(fma(-den, q, num), r, q);
}
template <typename T>
__device__ __forceinline__ T nurbs_div(T num, T den) {
    if (sizeof(T) == 8 && nurbs_den_is_normal(den)) return nurbs_div_normal(num, den);
    return num / den;
}

// First stencil node only (the i0 of dim_stencil, without the weights): key of the cell-sorted plan.
template <typename T>
__device__ __forceinline__ int dim_first_node(const DimParams<T>& D, T x) {
    const typename KeyOf<T>::type kx = to_key(x);
    const int n = D.n_knots;
    if (D.kind == LINEAR) return clampi(dim_count<T, false>(D, x, kx), 1, n - 1) - 1;
    if (D.kind == CONSTANT) return (x >= D.knots[n - 1]) ? n - 1 : clampi(dim_count<T, true>(D, x, kx), 1, n - 1) - 1;
    const int p = D.degree;
    return clampi(dim_count<T, true>(D, x, kx), p + 1, n - p - 1) - p - 1;
}

}  // namespace dind
