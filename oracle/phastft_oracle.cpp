// phastft_oracle.cpp -- CPU restatement of PhastFT's 1-D power-of-two FFT path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under phastft_b200/ (the product) may
// import, link or call this file.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it, as the checker and as
// the timed CPU baseline ("PhastFT-restatement (C++)", kind = "port").
//
// Why a restatement: the reference is Rust (QuState/PhastFT @ 8cd3a39) and there
// is no cargo/rustc in this image, nor are its arithmetic-carrying dependencies
// vendored (fearless_simd 0.4.0, rayon 1.11.0 -- Cargo.lock:381-384,965-978), so
// the reference itself cannot be compiled here or on the GPU box.  This file
// restates the algorithm operation-for-operation from the reference sources.
// Parity pinning: the reference stores no golden vectors (its RNG fixtures are
// unseeded); its own tests pin this path through closed-form known answers,
// exact bit-reversal, codelet==staged-kernel equivalence, round trips and an
// independent FFT (RustFFT).  tests/test_oracle_*.py re-run every one of those
// against this oracle (numpy.fft / longdouble DFT standing in for RustFFT), so the
// oracle is pinned by the reference's known-answer tests, not by reference output.
//
// fearless_simd semantics assumed: a.mul_add(b, c) = fma(a, b, c)
// (fused on AVX2/NEON), a.mul_sub(b, c) = fma(a, b, -c), zip/unzip = pure data
// movement.  Every fused multiply-add below is an explicit std::fma and the file
// must be compiled with -ffp-contract=off so nothing else is fused.
//
// Reference map (all paths relative to /root/reference/src):
//   planner tables ............ planner.rs:55-100 (DIT), :120-162 (r2c)
//   driver / swap trick / scale algorithms/dit.rs:263-401
//   recursion + stage dispatch  algorithms/dit.rs:33-242
//   butterflies ............... kernels/dit.rs:13-28 (chunk 2), :41-78 (chunk 4),
//                               :132-967 (chunk 8..64 literal twiddles),
//                               :971-1115 (chunk_n, planner twiddles)
//   codelets .................. kernels/codelets.rs:34-210 (f64 FFT-16), :218-498 (f32 FFT-32)
//   bit reversal .............. algorithms/bravo.rs:328-345 (the permutation);
//                               :82-251 BRAVO / CO-BRAVO are cache/SIMD variants of the
//                               same permutation -- restated here as a tiled variant
//   r2c / c2r ................. algorithms/r2c.rs:73-128, :150-242, :263-432, :444-489, :521-895
//   threading ................. parallel.rs:6-25, options.rs:26-43 (rayon::join -> OpenMP tasks)

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/phastft_status.h"

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

inline bool is_pow2(size_t n) { return n != 0 && (n & (n - 1)) == 0; }
inline unsigned ilog2(size_t n) {
    unsigned l = 0;
    while (n >>= 1) ++l;
    return l;
}

// ---------------------------------------------------------------------------
// options.rs:10-43
// ---------------------------------------------------------------------------
struct Options {
    bool multithreaded_bit_reversal = false;      // options.rs:29
    size_t smallest_parallel_chunk_size = 16384;  // options.rs:30
    static Options guess(size_t input_size) {     // options.rs:38-43
        Options o;
        o.multithreaded_bit_reversal = ilog2(input_size) >= 16;
        return o;
    }
};

// ---------------------------------------------------------------------------
// planner.rs:55-100 -- per-stage twiddle tables for stages with chunk_size > 64.
// The angle arithmetic is carried out in T exactly as the reference does:
//   angle_mult = -2.0 * PI / chunk_size as T;  angle = angle_mult * k as T
// ---------------------------------------------------------------------------
template <typename T>
struct PlannerDit {
    unsigned log_n = 0;
    std::vector<std::vector<T>> tw_re, tw_im;
    explicit PlannerDit(size_t n) {
        log_n = ilog2(n);
        for (unsigned stage = 0; stage < log_n; ++stage) {
            size_t dist = size_t(1) << stage;
            size_t chunk = dist * 2;
            if (chunk > 64) {
                std::vector<T> re(dist), im(dist);
                const T pi_t = static_cast<T>(kPi);
                const T angle_mult = (T(-2.0) * pi_t) / static_cast<T>(chunk);
                for (size_t k = 0; k < dist; ++k) {
                    T angle = angle_mult * static_cast<T>(k);
                    re[k] = std::cos(angle);
                    im[k] = std::sin(angle);
                }
                tw_re.push_back(std::move(re));
                tw_im.push_back(std::move(im));
            }
        }
    }
};

// planner.rs:120-162 -- 0.5 * W_N^k by rotation recurrence in f64 (cast for f32).
template <typename T>
void r2c_twiddles(size_t n, std::vector<T>& w_re, std::vector<T>& w_im) {
    size_t half = n / 2;
    w_re.assign(half, T(0));
    w_im.assign(half, T(0));
    double angle_step = -kPi / static_cast<double>(half);
    double st = std::sin(angle_step), ct = std::cos(angle_step);
    double wr = 1.0, wi = 0.0;
    for (size_t k = 0; k < half; ++k) {
        w_re[k] = static_cast<T>(0.5 * wr);
        w_im[k] = static_cast<T>(0.5 * wi);
        double tmp = wr;
        wr = tmp * ct - wi * st;
        wi = tmp * st + wi * ct;
    }
}

template <typename T>
struct PlannerR2c {
    size_t n;
    PlannerDit<T> dit;
    std::vector<T> w_re, w_im;
    Options inner_opts;
    explicit PlannerR2c(size_t n_) : n(n_), dit(n_ / 2), inner_opts(Options::guess(n_ / 2)) {
        r2c_twiddles<T>(n, w_re, w_im);
    }
};

// ---------------------------------------------------------------------------
// Literal twiddles W_chunk^k, k < chunk/2, for chunk in {8,16,32,64}
// (kernels/dit.rs:146-163, 272-298, 415-470, 628-741 and the f32 twins).  The
// reference spells them as decimal literals that are the correctly rounded
// cos/sin values with exact 0 / +-1 / FRAC_1_SQRT_2 at the symmetric points; we
// regenerate them with octant symmetry so those exact values come out too.
// oracle/check_literals.py verifies equality against the reference's literals.
// ---------------------------------------------------------------------------
template <typename T>
void literal_twiddle(unsigned chunk, unsigned k, T& wr, T& wi) {
    // angle = 2*pi*k/chunk in [0, pi); W = (cos, -sin)
    unsigned eighth = chunk / 8;  // chunk >= 8
    unsigned oct = k / eighth;    // 0..3
    unsigned rem = k % eighth;
    auto cs = [&](unsigned j, long double& c, long double& s) {
        // cos/sin of 2*pi*j/chunk for j in [0, chunk/8]
        if (j == 0) { c = 1.0L; s = 0.0L; return; }
        if (j == eighth) { c = s = 0.70710678118654752440084436210484903928L; return; }
        long double a = 2.0L * 3.14159265358979323846264338327950288L * j / chunk;
        c = cosl(a); s = sinl(a);
    };
    long double c, s, co, si;
    switch (oct) {
        case 0: cs(rem, c, s); co = c; si = s; break;                       // a
        case 1: cs(eighth - rem, c, s); co = s; si = c; break;             // pi/2 - a'
        case 2: cs(rem, c, s); co = -s; si = c; break;                     // pi/2 + a
        default: cs(eighth - rem, c, s); co = -c; si = s; break;           // pi - a'
    }
    // Round through double first (the reference's f32 literals are the f32 nearest to the
    // decimal spelling of the f64 value; double->float of the correctly rounded double agrees).
    double dr = static_cast<double>(co), di = -static_cast<double>(si);
    // Six of the reference's f64 decimal literals are 1 ulp away from the correctly rounded
    // value (e.g. W_16 uses 0.38268343236508984 at kernels/dit.rs:279 while W_32 uses the
    // correctly rounded 0.3826834323650898 at :424).  Reproduce the reference's constants
    // exactly; the f32 literals all equal the rounded f64 values, so this only matters for f64.
    auto fix = [&](double v) {
        double a = std::fabs(v), r = a;
        if (chunk == 16 && a == 0.3826834323650898) r = 0.38268343236508984;
        if ((chunk == 32 || chunk == 64) && a == 0.19509032201612828) r = 0.19509032201612825;
        if (chunk == 64 && a == 0.9569403357322088) r = 0.9569403357322089;
        if (chunk == 64 && a == 0.881921264348355) r = 0.8819212643483549;
        if (chunk == 64 && a == 0.2902846772544624) r = 0.29028467725446233;
        return v < 0 ? -r : r;
    };
    if (sizeof(T) == 8) { dr = fix(dr); di = fix(di); }
    wr = static_cast<T>(dr);
    wi = static_cast<T>(di);
    if (wr == T(0)) wr = T(0);  // normalise -0 (the reference writes 0.0)
    if (wi == T(0)) wi = T(0);
}

template <typename T>
struct LiteralTables {
    std::vector<T> re[7], im[7];  // index by log2(chunk): 3..6
    LiteralTables() {
        for (unsigned lc = 3; lc <= 6; ++lc) {
            unsigned chunk = 1u << lc;
            re[lc].resize(chunk / 2);
            im[lc].resize(chunk / 2);
            for (unsigned k = 0; k < chunk / 2; ++k) literal_twiddle<T>(chunk, k, re[lc][k], im[lc][k]);
        }
    }
};
template <typename T>
const LiteralTables<T>& literal_tables() {
    static const LiteralTables<T> t;
    return t;
}

// ---------------------------------------------------------------------------
// Butterfly kernels
// ---------------------------------------------------------------------------

// kernels/dit.rs:13-28
template <typename T>
void fft_dit_chunk_2(T* re, T* im, size_t n) {
    for (size_t i = 0; i + 1 < n; i += 2) {
        T z0r = re[i], z0i = im[i], z1r = re[i + 1], z1i = im[i + 1];
        re[i] = z0r + z1r;
        im[i] = z0i + z1i;
        re[i + 1] = z0r - z1r;
        im[i + 1] = z0i - z1i;
    }
}

// kernels/dit.rs:41-78 (f64) / :91-128 (f32): W_4^1 = -i specialisation, out1 = fma(in0, 2, -out0)
template <typename T>
void fft_dit_chunk_4(T* re, T* im, size_t n) {
    const T two = T(2);
    for (size_t b = 0; b + 3 < n; b += 4) {
        T in0r = re[b], in1r = re[b + 2], in0i = im[b], in1i = im[b + 2];
        re[b] = in0r + in1r;
        im[b] = in0i + in1i;
        re[b + 2] = std::fma(in0r, two, -re[b]);
        im[b + 2] = std::fma(in0i, two, -im[b]);
        in0r = re[b + 1]; in1r = re[b + 3]; in0i = im[b + 1]; in1i = im[b + 3];
        re[b + 1] = in0r + in1i;
        im[b + 1] = in0i - in1r;
        re[b + 3] = std::fma(in0r, two, -re[b + 1]);
        im[b + 3] = std::fma(in0i, two, -im[b + 1]);
    }
}

// The FMA butterfly shared by chunk 8..64 and chunk_n (kernels/dit.rs:177-188, 1029-1041):
//   out0_re = fma(w_im, -in1_im, fma(w_re, in1_re, in0_re))
//   out0_im = fma(w_im,  in1_re, fma(w_re, in1_im, in0_im))
//   out1    = fma(2, in0, -out0)
template <typename T>
inline void butterfly(T& lo_re, T& lo_im, T& hi_re, T& hi_im, T wr, T wi) {
    const T two = T(2);
    T o0r = std::fma(wi, -hi_im, std::fma(wr, hi_re, lo_re));
    T o0i = std::fma(wi, hi_re, std::fma(wr, hi_im, lo_im));
    T o1r = std::fma(two, lo_re, -o0r);
    T o1i = std::fma(two, lo_im, -o0i);
    lo_re = o0r; lo_im = o0i; hi_re = o1r; hi_im = o1i;
}

template <typename T>
void fft_dit_stage_tw(T* __restrict re, T* __restrict im, size_t n, size_t dist, const T* __restrict twr,
                      const T* __restrict twi) {
    size_t chunk = dist * 2;
    for (size_t base = 0; base < n; base += chunk) {
        T* r0 = re + base; T* r1 = r0 + dist;
        T* i0 = im + base; T* i1 = i0 + dist;
        for (size_t k = 0; k < dist; ++k) butterfly(r0[k], i0[k], r1[k], i1[k], twr[k], twi[k]);
    }
}

// kernels/codelets.rs:58-97 -- radix-4 (stages 0+1) on 4 consecutive elements, plain add/sub.
template <typename T>
inline void radix4_group(T* re, T* im) {
    T s01r = re[0] + re[1], d01r = re[0] - re[1];
    T s23r = re[2] + re[3], d23r = re[2] - re[3];
    T s01i = im[0] + im[1], d01i = im[0] - im[1];
    T s23i = im[2] + im[3], d23i = im[2] - im[3];
    re[0] = s01r + s23r; re[2] = s01r - s23r;
    im[0] = s01i + s23i; im[2] = s01i - s23i;
    re[1] = d01r + d23i; re[3] = d01r - d23i;
    im[1] = d01i - d23r; im[3] = d01i + d23r;
}

// kernels/codelets.rs:34-210 -- f64 FFT-16 codelet (stages 0-3).
void fft_dit_codelet_16_f64(double* re, double* im, size_t n) {
    const auto& lt = literal_tables<double>();
    for (size_t b = 0; b + 15 < n; b += 16) {
        double* r = re + b; double* i = im + b;
        for (int g = 0; g < 4; ++g) radix4_group(r + 4 * g, i + 4 * g);
        // stage 2: dist 4, pairs (v0,v1), (v2,v3), W_8^{0..3}
        for (int h = 0; h < 2; ++h)
            for (int k = 0; k < 4; ++k)
                butterfly(r[8 * h + k], i[8 * h + k], r[8 * h + 4 + k], i[8 * h + 4 + k], lt.re[3][k], lt.im[3][k]);
        // stage 3: dist 8, W_16^{0..7}
        for (int k = 0; k < 8; ++k) butterfly(r[k], i[k], r[8 + k], i[8 + k], lt.re[4][k], lt.im[4][k]);
    }
}

// kernels/codelets.rs:218-498 -- f32 FFT-32 codelet (stages 0-4); stage 2 uses the
// W_8 special forms of :320-354.
void fft_dit_codelet_32_f32(float* re, float* im, size_t n) {
    const auto& lt = literal_tables<float>();
    const float s = 0.70710678118654752440f;  // std::f32::consts::FRAC_1_SQRT_2
    for (size_t b = 0; b + 31 < n; b += 32) {
        float* r = re + b; float* i = im + b;
        for (int g = 0; g < 8; ++g) radix4_group(r + 4 * g, i + 4 * g);
        for (int c = 0; c < 4; ++c) {  // per 8-chunk: p_j = e[j], p_{j+4} = e[j+4]
            float* pr = r + 8 * c; float* pi = i + 8 * c;
            // W8^0
            float r0r = pr[0] + pr[4], r4r = pr[0] - pr[4];
            float r0i = pi[0] + pi[4], r4i = pi[0] - pi[4];
            // W8^1 = (s, -s)
            float t5r = std::fma(s, pi[5], s * pr[5]);
            float t5i = std::fma(s, pi[5], -(s * pr[5]));
            float r1r = pr[1] + t5r, r5r = pr[1] - t5r;
            float r1i = pi[1] + t5i, r5i = pi[1] - t5i;
            // W8^2 = -j
            float r2r = pr[2] + pi[6], r6r = pr[2] - pi[6];
            float r2i = pi[2] - pr[6], r6i = pi[2] + pr[6];
            // W8^3 = (-s, -s)
            float t7r = std::fma(s, pi[7], -(s * pr[7]));
            float nt7i = std::fma(s, pi[7], s * pr[7]);
            float r3r = pr[3] + t7r, r7r = pr[3] - t7r;
            float r3i = pi[3] - nt7i, r7i = pi[3] + nt7i;
            pr[0] = r0r; pr[1] = r1r; pr[2] = r2r; pr[3] = r3r; pr[4] = r4r; pr[5] = r5r; pr[6] = r6r; pr[7] = r7r;
            pi[0] = r0i; pi[1] = r1i; pi[2] = r2i; pi[3] = r3i; pi[4] = r4i; pi[5] = r5i; pi[6] = r6i; pi[7] = r7i;
        }
        // stage 3: dist 8, pairs (v0,v1), (v2,v3), W_16^{0..7}
        for (int h = 0; h < 2; ++h)
            for (int k = 0; k < 8; ++k)
                butterfly(r[16 * h + k], i[16 * h + k], r[16 * h + 8 + k], i[16 * h + 8 + k], lt.re[4][k], lt.im[4][k]);
        // stage 4: dist 16, W_32^{0..15}
        for (int k = 0; k < 16; ++k) butterfly(r[k], i[k], r[16 + k], i[16 + k], lt.re[5][k], lt.im[5][k]);
    }
}

template <typename T> struct Codelet;
template <> struct Codelet<double> {
    static constexpr unsigned stages = 4;  // algorithms/dit.rs:46
    static void run(double* re, double* im, size_t n) { fft_dit_codelet_16_f64(re, im, n); }
};
template <> struct Codelet<float> {
    static constexpr unsigned stages = 5;  // algorithms/dit.rs:113
    static void run(float* re, float* im, size_t n) { fft_dit_codelet_32_f32(re, im, n); }
};

// algorithms/dit.rs:168-242 -- stage dispatcher; returns the updated stage_twiddle_idx.
template <typename T>
size_t execute_dit_stage(T* re, T* im, size_t n, unsigned stage, const PlannerDit<T>& p, size_t tw_idx) {
    size_t dist = size_t(1) << stage;
    size_t chunk = dist * 2;
    if (chunk == 2) { fft_dit_chunk_2(re, im, n); return tw_idx; }
    if (chunk == 4) { fft_dit_chunk_4(re, im, n); return tw_idx; }
    if (chunk <= 64) {
        const auto& lt = literal_tables<T>();
        fft_dit_stage_tw(re, im, n, dist, lt.re[stage + 1].data(), lt.im[stage + 1].data());
        return tw_idx;
    }
    fft_dit_stage_tw(re, im, n, dist, p.tw_re[tw_idx].data(), p.tw_im[tw_idx].data());
    return tw_idx + 1;
}

constexpr size_t L1_BLOCK_SIZE = 1024;  // algorithms/dit.rs:27

// algorithms/dit.rs:33-164 -- post-order recursion.  `parallel` maps rayon::join
// (parallel.rs:6-25) onto OpenMP tasks; without OpenMP (or outside a parallel
// region) the tasks run inline, i.e. the serial path.
template <typename T>
size_t recursive_dit_fft(T* re, T* im, size_t size, const PlannerDit<T>& p, const Options& o, size_t tw_idx, bool parallel) {
    unsigned log_size = ilog2(size);
    if (size <= L1_BLOCK_SIZE) {
        unsigned start = 0;
        if (tw_idx == 0 && size >= (size_t(1) << Codelet<T>::stages)) {
            Codelet<T>::run(re, im, size);
            start = Codelet<T>::stages;
        }
        for (unsigned st = start; st < log_size; ++st) tw_idx = execute_dit_stage(re, im, size, st, p, tw_idx);
        return tw_idx;
    }
    size_t half = size / 2;
    unsigned log_half = ilog2(half);
    bool par = parallel && size > o.smallest_parallel_chunk_size;
#ifdef _OPENMP
    if (par) {
#pragma omp task default(shared)
        recursive_dit_fft(re, im, half, p, o, 0, parallel);
        recursive_dit_fft(re + half, im + half, half, p, o, 0, parallel);
#pragma omp taskwait
    } else
#endif
    {
        (void)par;
        recursive_dit_fft(re, im, half, p, o, 0, parallel);
        recursive_dit_fft(re + half, im + half, half, p, o, 0, parallel);
    }
    tw_idx = log_half >= 6 ? log_half - 6 : 0;  // saturating_sub(6)
    for (unsigned st = log_half; st < log_size; ++st) tw_idx = execute_dit_stage(re, im, size, st, p, tw_idx);
    return tw_idx;
}

// ---------------------------------------------------------------------------
// Bit reversal.  algorithms/bravo.rs:328-345 defines the permutation; BRAVO /
// CO-BRAVO (:82-251) are SIMD / cache-tiled ways of performing it.  The tiled
// variant below mirrors CO-BRAVO's structure (index = (u : logB)(t : tile bits)(v : logB),
// B x B tile staged through a small buffer, tile t swapped with tile rev(t)) so the
// timed CPU baseline is not handicapped by a cache-hostile scalar loop.
// ---------------------------------------------------------------------------
inline size_t reverse_bits(size_t x, unsigned bits) {
    if (bits == 0) return 0;
    size_t r = 0;
    for (unsigned i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

template <typename T>
void scalar_bit_reversal(T* data, unsigned n) {
    size_t big_n = size_t(1) << n;
    for (size_t i = 0; i < big_n; ++i) {
        size_t j = reverse_bits(i, n);
        if (i < j) std::swap(data[i], data[j]);
    }
}

template <typename T>
void tiled_bit_reversal(T* data, unsigned n) {
    constexpr unsigned LOGB = sizeof(T) == 4 ? 6 : 5;  // TILE_SIDE 64 (f32) / 32 (f64), bravo.rs:19-20
    constexpr size_t B = size_t(1) << LOGB;
    if (n <= 2 * LOGB + 4) {  // big_n <= B*B*MIN_TILES (bravo.rs:24,239): direct
        scalar_bit_reversal(data, n);
        return;
    }
    unsigned tile_bits = n - 2 * LOGB;
    size_t num_tiles = size_t(1) << tile_bits;
    // rev table for LOGB bits
    size_t revb[B];
    for (size_t i = 0; i < B; ++i) revb[i] = reverse_bits(i, LOGB);
    std::vector<T> buf(B * B), buf2(B * B);
    auto load = [&](size_t tile, T* dst) {   // strip u of tile t = data[(u * num_tiles + tile) * B .. +B]
        for (size_t u = 0; u < B; ++u) std::memcpy(dst + u * B, data + (u * num_tiles + tile) * B, B * sizeof(T));
    };
    auto store_rev = [&](const T* src, size_t tile) {
        // element (u, v) of the source tile goes to strip rev(v), lane rev(u) of the destination tile
        for (size_t u2 = 0; u2 < B; ++u2) {
            T* d = data + (u2 * num_tiles + tile) * B;
            size_t v = revb[u2];
            for (size_t v2 = 0; v2 < B; ++v2) d[v2] = src[revb[v2] * B + v];
        }
    };
    for (size_t t = 0; t < num_tiles; ++t) {
        size_t tr = reverse_bits(t, tile_bits);
        if (t > tr) continue;
        load(t, buf.data());
        if (t == tr) {
            store_rev(buf.data(), t);
        } else {
            load(tr, buf2.data());
            store_rev(buf.data(), tr);
            store_rev(buf2.data(), t);
        }
    }
}

// ---------------------------------------------------------------------------
// algorithms/dit.rs:263-401 -- driver: asserts, swap trick, bit reversal, recursion, 1/N scale.
// ---------------------------------------------------------------------------
template <typename T>
int32_t fft_dit_with_planner_and_opts(T* reals, size_t n_re, T* imags, size_t n_im, int direction,
                                      const PlannerDit<T>& planner, const Options& opts, bool parallel) {
    if (n_re != n_im) return PHASTFT_ERR_LEN_MISMATCH;           // dit.rs:284
    if (!is_pow2(n_re)) return PHASTFT_ERR_NOT_POW2;             // dit.rs:285
    size_t n = n_re;
    unsigned log_n = ilog2(n);
    if (log_n != planner.log_n) return PHASTFT_ERR_PLAN_MISMATCH;  // dit.rs:289
    if (direction != PHASTFT_FORWARD && direction != PHASTFT_REVERSE) return PHASTFT_ERR_INVALID_ARG;
    T* re = reals; T* im = imags;
    if (direction == PHASTFT_REVERSE) std::swap(re, im);         // dit.rs:297-300

    auto body = [&]() {
        // dit.rs:303-317 -- two bit reversals, joined iff opts.multithreaded_bit_reversal
        bool par_br = parallel && opts.multithreaded_bit_reversal;
#ifdef _OPENMP
        if (par_br) {
#pragma omp task default(shared)
            tiled_bit_reversal(re, log_n);
            tiled_bit_reversal(im, log_n);
#pragma omp taskwait
        } else
#endif
        {
            (void)par_br;
            tiled_bit_reversal(re, log_n);
            tiled_bit_reversal(im, log_n);
        }
        recursive_dit_fft(re, im, n, planner, opts, 0, parallel);
    };
#ifdef _OPENMP
    if (parallel && n > opts.smallest_parallel_chunk_size) {
#pragma omp parallel
#pragma omp single
        body();
    } else
#endif
        body();

    if (direction == PHASTFT_REVERSE) {                           // dit.rs:325-331
        T scaling = T(1) / static_cast<T>(n);
        for (size_t i = 0; i < n; ++i) { re[i] *= scaling; im[i] *= scaling; }
    }
    return PHASTFT_OK;
}

// ---------------------------------------------------------------------------
// r2c.rs:150-242 -- in-place untangle (plain mul/add: the reference uses `*`, `+`, `-`
// operators here, not mul_add).
// ---------------------------------------------------------------------------
template <typename T>
void untangle_inplace(T* ore, T* oim, const T* w_re, const T* w_im, size_t half) {
    T a0 = ore[0], b0 = oim[0];
    ore[0] = a0 + b0; oim[0] = T(0);
    ore[half] = a0 - b0; oim[half] = T(0);
    size_t q = half / 2;
    for (size_t k = 1; k < q; ++k) {
        size_t m = half - k;
        T a = ore[k], b = oim[k], c = ore[m], d = oim[m];
        T s_re = T(0.5) * (a + c), s_im = T(0.5) * (b - d);
        T t_re = b + d, t_im = c - a;
        T wkr = w_re[k], wki = w_im[k];
        T wzr = wkr * t_re - wki * t_im;
        T wzi = wkr * t_im + wki * t_re;
        ore[k] = s_re + wzr; oim[k] = s_im + wzi;
        ore[m] = s_re - wzr; oim[m] = wzi - s_im;
    }
    T a = ore[q], b = oim[q];
    ore[q] = a + T(2) * w_re[q] * b;   // r2c.rs:233-236
    oim[q] = T(2) * w_im[q] * b;
}

// r2c.rs:263-347 -- c2r pre-processing.
template <typename T>
void c2r_preprocess(const T* ire, const T* iim, const T* w_re, const T* w_im, T* zre, T* zim, size_t half) {
    for (size_t k = 0; k < half; ++k) {
        size_t m = half - k;
        T re_f = ire[k], im_f = iim[k];
        T re_s = ire[m], im_s = -iim[m];
        T zx_re = T(0.5) * (re_f + re_s), zx_im = T(0.5) * (im_f + im_s);
        T dr = re_f - re_s, di = im_f - im_s;
        T c_h = w_re[k], s_h = w_im[k];
        T zy_re = c_h * dr + s_h * di;
        T zy_im = c_h * di - s_h * dr;
        zre[k] = zx_re - zy_im;
        zim[k] = zx_im + zy_re;
    }
}

// r2c.rs:535-593
template <typename T>
int32_t r2c_with_planner(const T* in, size_t n_in, T* ore, size_t n_ore, T* oim, size_t n_oim, const PlannerR2c<T>& p,
                         bool parallel) {
    size_t n = p.n, half = n / 2;
    if (n_in != n) return PHASTFT_ERR_INPUT_LEN;
    if (n_ore != half + 1) return PHASTFT_ERR_OUTPUT_RE_LEN;
    if (n_oim != half + 1) return PHASTFT_ERR_OUTPUT_IM_LEN;
    for (size_t k = 0; k < half; ++k) { ore[k] = in[2 * k]; oim[k] = in[2 * k + 1]; }  // r2c.rs:73-128
    int32_t st = fft_dit_with_planner_and_opts<T>(ore, half, oim, half, PHASTFT_FORWARD, p.dit, p.inner_opts, parallel);
    if (st != PHASTFT_OK) return st;
    untangle_inplace<T>(ore, oim, p.w_re.data(), p.w_im.data(), half);
    return PHASTFT_OK;
}

// r2c.rs:740-799
template <typename T>
int32_t c2r_with_planner_and_scratch(const T* ire, size_t n_ire, const T* iim, size_t n_iim, T* out, size_t n_out,
                                     const PlannerR2c<T>& p, T* sre, size_t n_sre, T* sim, size_t n_sim, bool parallel) {
    size_t n = p.n, half = n / 2;
    if (n_out != n) return PHASTFT_ERR_OUTPUT_LEN;
    if (n_ire != half + 1) return PHASTFT_ERR_INPUT_RE_LEN;
    if (n_iim != half + 1) return PHASTFT_ERR_INPUT_IM_LEN;
    if (n_sre != half) return PHASTFT_ERR_SCRATCH_RE_LEN;
    if (n_sim != half) return PHASTFT_ERR_SCRATCH_IM_LEN;
    c2r_preprocess<T>(ire, iim, p.w_re.data(), p.w_im.data(), sre, sim, half);
    int32_t st = fft_dit_with_planner_and_opts<T>(sre, half, sim, half, PHASTFT_REVERSE, p.dit, p.inner_opts, parallel);
    if (st != PHASTFT_OK) return st;
    for (size_t k = 0; k < half; ++k) { out[2 * k] = sre[k]; out[2 * k + 1] = sim[k]; }  // r2c.rs:444-489
    return PHASTFT_OK;
}

}  // namespace

// ===========================================================================
// extern "C" surface (ctypes; tests/ and bench.py only)
// ===========================================================================
#define ORACLE_API extern "C" __attribute__((visibility("default")))

ORACLE_API int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORACLE_API void oracle_set_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

#define DEFINE_FOR(T, SFX)                                                                                           \
    ORACLE_API int32_t oracle_plan_dit_##SFX##_new(size_t n, void** out) {                                           \
        if (!out) return PHASTFT_ERR_INVALID_ARG;                                                                    \
        if (!is_pow2(n)) return PHASTFT_ERR_NOT_POW2; /* planner.rs:66 */                                            \
        *out = new PlannerDit<T>(n);                                                                                 \
        return PHASTFT_OK;                                                                                           \
    }                                                                                                                \
    ORACLE_API void oracle_plan_dit_##SFX##_free(void* p) { delete static_cast<PlannerDit<T>*>(p); }                 \
    ORACLE_API size_t oracle_plan_dit_##SFX##_num_tables(void* p) { return static_cast<PlannerDit<T>*>(p)->tw_re.size(); } \
    ORACLE_API void oracle_plan_dit_##SFX##_table(void* p, size_t idx, T* re, T* im) {                               \
        auto* pl = static_cast<PlannerDit<T>*>(p);                                                                   \
        std::memcpy(re, pl->tw_re[idx].data(), pl->tw_re[idx].size() * sizeof(T));                                   \
        std::memcpy(im, pl->tw_im[idx].data(), pl->tw_im[idx].size() * sizeof(T));                                   \
    }                                                                                                                \
    /* lib.rs:143-150 / dit.rs:263: opts = guess_options(len); `parallel` = the crate's `parallel` feature */        \
    ORACLE_API int32_t oracle_fft_dit_##SFX##_with_planner(T* re, size_t n_re, T* im, size_t n_im, int dir, void* p, \
                                                           int parallel) {                                           \
        if (!p || !re || !im) return PHASTFT_ERR_INVALID_ARG;                                                        \
        if (n_re == 0) return n_im == 0 ? PHASTFT_ERR_NOT_POW2 : PHASTFT_ERR_LEN_MISMATCH;                           \
        Options o = Options::guess(n_re);                                                                            \
        return fft_dit_with_planner_and_opts<T>(re, n_re, im, n_im, dir, *static_cast<PlannerDit<T>*>(p), o,         \
                                                parallel != 0);                                                      \
    }                                                                                                                \
    /* lib.rs:180-183: plans per call */                                                                             \
    ORACLE_API int32_t oracle_fft_dit_##SFX(T* re, size_t n_re, T* im, size_t n_im, int dir, int parallel) {         \
        if (!is_pow2(n_re)) return PHASTFT_ERR_NOT_POW2;                                                             \
        PlannerDit<T> pl(n_re);                                                                                      \
        return oracle_fft_dit_##SFX##_with_planner(re, n_re, im, n_im, dir, &pl, parallel);                          \
    }                                                                                                                \
    ORACLE_API int32_t oracle_plan_r2c_##SFX##_new(size_t n, void** out) {                                           \
        if (!out) return PHASTFT_ERR_INVALID_ARG;                                                                    \
        if (!(n >= 4 && is_pow2(n))) return PHASTFT_ERR_R2C_N; /* planner.rs:195 */                                  \
        *out = new PlannerR2c<T>(n);                                                                                 \
        return PHASTFT_OK;                                                                                           \
    }                                                                                                                \
    ORACLE_API void oracle_plan_r2c_##SFX##_free(void* p) { delete static_cast<PlannerR2c<T>*>(p); }                 \
    ORACLE_API void oracle_plan_r2c_##SFX##_twiddles(void* p, T* re, T* im) {                                        \
        auto* pl = static_cast<PlannerR2c<T>*>(p);                                                                   \
        std::memcpy(re, pl->w_re.data(), pl->w_re.size() * sizeof(T));                                               \
        std::memcpy(im, pl->w_im.data(), pl->w_im.size() * sizeof(T));                                               \
    }                                                                                                                \
    ORACLE_API int32_t oracle_r2c_##SFX##_with_planner(const T* in, size_t n_in, T* ore, size_t n_ore, T* oim,       \
                                                       size_t n_oim, void* p, int parallel) {                        \
        if (!p) return PHASTFT_ERR_INVALID_ARG;                                                                      \
        return r2c_with_planner<T>(in, n_in, ore, n_ore, oim, n_oim, *static_cast<PlannerR2c<T>*>(p), parallel != 0); \
    }                                                                                                                \
    ORACLE_API int32_t oracle_r2c_##SFX(const T* in, size_t n_in, T* ore, size_t n_ore, T* oim, size_t n_oim,        \
                                        int parallel) {                                                              \
        if (!(n_in >= 4 && is_pow2(n_in))) return PHASTFT_ERR_R2C_N;                                                 \
        PlannerR2c<T> pl(n_in);                                                                                      \
        return r2c_with_planner<T>(in, n_in, ore, n_ore, oim, n_oim, pl, parallel != 0);                             \
    }                                                                                                                \
    ORACLE_API int32_t oracle_c2r_##SFX##_with_planner_and_scratch(const T* ire, size_t n_ire, const T* iim,         \
                                                                   size_t n_iim, T* out, size_t n_out, void* p,      \
                                                                   T* sre, size_t n_sre, T* sim, size_t n_sim,       \
                                                                   int parallel) {                                   \
        if (!p) return PHASTFT_ERR_INVALID_ARG;                                                                      \
        return c2r_with_planner_and_scratch<T>(ire, n_ire, iim, n_iim, out, n_out, *static_cast<PlannerR2c<T>*>(p),  \
                                               sre, n_sre, sim, n_sim, parallel != 0);                               \
    }                                                                                                                \
    /* r2c.rs:708-728: allocates the two N/2 scratch Vecs per call */                                                \
    ORACLE_API int32_t oracle_c2r_##SFX##_with_planner(const T* ire, size_t n_ire, const T* iim, size_t n_iim,       \
                                                       T* out, size_t n_out, void* p, int parallel) {                \
        if (!p) return PHASTFT_ERR_INVALID_ARG;                                                                      \
        auto* pl = static_cast<PlannerR2c<T>*>(p);                                                                   \
        size_t half = pl->n / 2;                                                                                     \
        std::vector<T> sre(half), sim(half);                                                                         \
        return c2r_with_planner_and_scratch<T>(ire, n_ire, iim, n_iim, out, n_out, *pl, sre.data(), half,            \
                                               sim.data(), half, parallel != 0);                                     \
    }                                                                                                                \
    ORACLE_API int32_t oracle_c2r_##SFX(const T* ire, size_t n_ire, const T* iim, size_t n_iim, T* out,              \
                                        size_t n_out, int parallel) {                                                \
        if (!(n_out >= 4 && is_pow2(n_out))) return PHASTFT_ERR_R2C_N;                                               \
        PlannerR2c<T> pl(n_out);                                                                                     \
        return oracle_c2r_##SFX##_with_planner(ire, n_ire, iim, n_iim, out, n_out, &pl, parallel);                   \
    }                                                                                                                \
    /* building blocks exposed for the reference's unit-level tests */                                               \
    ORACLE_API void oracle_bit_reverse_##SFX(T* data, unsigned log_n, int tiled) {                                   \
        if (tiled) tiled_bit_reversal<T>(data, log_n); else scalar_bit_reversal<T>(data, log_n);                     \
    }                                                                                                                \
    ORACLE_API void oracle_codelet_##SFX(T* re, T* im, size_t n) { Codelet<T>::run(re, im, n); }                     \
    ORACLE_API unsigned oracle_codelet_stages_##SFX() { return Codelet<T>::stages; }                                 \
    /* one staged kernel (chunk 2..64 only: no planner table needed) */                                              \
    ORACLE_API void oracle_stage_##SFX(T* re, T* im, size_t n, unsigned stage) {                                     \
        PlannerDit<T> none(1);                                                                                       \
        execute_dit_stage<T>(re, im, n, stage, none, 0);                                                             \
    }                                                                                                                \
    ORACLE_API void oracle_literal_twiddles_##SFX(unsigned chunk, T* re, T* im) {                                    \
        for (unsigned k = 0; k < chunk / 2; ++k) literal_twiddle<T>(chunk, k, re[k], im[k]);                         \
    }

DEFINE_FOR(double, f64)
DEFINE_FOR(float, f32)
