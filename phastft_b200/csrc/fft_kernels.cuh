// fft_kernels.cuh -- the three sm_100a pass kernels of the B200 FFT.
//
// One launch = one "pass" over the whole signal: every CTA owns an R x C tile (R-point
// sub-transform x C adjacent columns/rows), runs all log2(R) butterfly stages on it with
// the tile held in shared memory between register-resident radix-4/8/16 stages, and
// touches HBM exactly once for the read and once for the write.  A transform of
// N = R_1 * R_2 [* R_3] points is 1, 2 or 3 such launches.
//
// What this replaces in the reference (all CPU, one sweep over memory per radix-2 stage):
//   algorithms/dit.rs:33-164   recursive cache-blocked DIT driver      -> the pass decomposition
//   algorithms/bravo.rs:82-251 BRAVO / CO-BRAVO bit-reversal           -> folded into addressing:
//        digit-reversed placement into shared memory on load (stage 1) and the transposed
//        store of the last pass; no separate permutation pass exists on the GPU
//   kernels/dit.rs:971-1115    fft_dit_chunk_n (planner-twiddle stage) -> stage twiddles from the
//        per-pass W_R table, inter-pass twiddles from a two-level W_N table built by the planner
//   algorithms/dit.rs:325-331  inverse 1/N scaling loop                -> fused into the last store
//
// Kernel kinds
//   KIND_COL   first / middle pass: tile rows are strided (stride B), columns contiguous.
//              load [r][c] -> store [k][c], same addresses (layout preserving, in-place safe).
//   KIND_TRANS last pass of a multi-pass plan: tile rows are contiguous sub-sequences, the C rows
//              of a tile are C adjacent values of the slowest input digit; the store is the
//              digit-reversing transpose, written as C-element contiguous runs.
//   KIND_ROW   whole transform in one CTA (N <= 4096 f64 / 8192 f32): rows contiguous in and out.
#pragma once
#include <type_traits>

#include "fft_device.cuh"

namespace phast {

enum { KIND_COL = 0, KIND_TRANS = 1, KIND_ROW = 2 };

template <typename T>
struct PassParams {
    const T* in_re;                // planar input (or interleaved (re,im) pairs if in_interleaved)
    const T* in_im;                // (no __restrict__: in-place passes alias in and out)
    T* out_re;                     // planar output (or interleaved if out_interleaved)
    T* out_im;
    long long in_bstride;          // elements between consecutive transforms of the batch
    long long out_bstride;
    int batch;                     // number of transforms (KIND_ROW: may not be a multiple of C)
    int in_interleaved;            // 1: in_re points at N (re,im) pairs (r2c first pass); 2: pairs, re/im swapped
    int out_interleaved;           // 1: out_re points at N pairs; 2: pairs with re/im swapped (c2r)
    // geometry
    int log2A;                     // A = number of sub-transform groups before this digit
    int log2B;                     // B = stride of this pass's digit = product of later pass sizes
    int log2R1;                    // KIND_TRANS: size of the first pass digit (tile columns run over it)
    int log2Rprev;                 // size of the previous pass digit (0 if no inter-pass twiddle)
    int blk_offset;                // KIND_COL: first linear tile index of this launch (L2-blocked sub-range of `a`)
    int kt_base;                   // KIND_TRANS: first k1 tile of this launch ...
    int log2_ktn;                  // ... and log2 of the number of k1 tiles it covers
    int pdl;                       // launched with programmatic stream serialization
    int has_tw;                    // apply inter-pass twiddle on load
    int tw_shift;                  // exponent scale: e_N = e_L << tw_shift   (N / L)
    Tw2 tw2;                       // two-level W_N table
    const cx<T>* __restrict__ tw_stage;  // W_R^e, e < R           (stage twiddles)
    const cx<T>* __restrict__ tw_wc;     // KIND_TRANS, 2-pass plans: W_L^(c*m), [c][m] layout
    T scale;                       // multiplied into the stored result (1/N for the inverse)
    // cluster exchange (XCH = 1 producer only): the consuming pass's tile is [CB rows][P2 points]
    int xch_log2P2;                // log2 of the consumer's row length (= its R)
    int xch_log2CB;                // log2 of the consumer's rows per CTA (= its C)
};

// Shared-memory tile addressing (units: complex elements).
//   CFAST kernels (COL, TRANS): [pos][c'] with c' = c ^ swz(pos)  (swz == 0 for COL)
//   ROW kernel: [c][pos'] with pos' = pos ^ swz(pos)
// swz(pos) takes the bits of `pos` in which the low bits of the *memory-order* index of stage 1
// land after digit reversal, so that the stage-1 scatter of a coalesced global read is
// bank-conflict free; later stages address whole rows / aligned runs and are unaffected.
template <class RL, int C, int KIND, typename T>
struct TileAddr {
    static constexpr int R = RL::R();
    static constexpr int LOG2R = ilog2_c(R);
    static constexpr int RS_LAST = RL::rad(RL::S - 1);
    static constexpr int SWZ_SHIFT = LOG2R - ilog2_c(RS_LAST);
    // 16-byte (f64) / 8-byte (f32) complex elements: a 128-byte wavefront holds 8 / 16 of them
    static constexpr int LANES_PER_WF = 128 / (2 * (int)sizeof(T));
    static constexpr int SWZ_MASK_WANT = LANES_PER_WF - 1;
    static constexpr int SWZ_MASK_TRANS = (C - 1) < SWZ_MASK_WANT ? (C - 1) : SWZ_MASK_WANT;
    static constexpr int SWZ_MASK_ROW = (R >= 4 * LANES_PER_WF) ? SWZ_MASK_WANT : 0;
    static __device__ __forceinline__ int at(int pos, int c) {
        if constexpr (KIND == KIND_COL) {
            return pos * C + c;
        } else if constexpr (KIND == KIND_TRANS) {
            return pos * C + (c ^ ((pos >> SWZ_SHIFT) & SWZ_MASK_TRANS));
        } else {
            return c * R + (pos ^ ((pos >> SWZ_SHIFT) & SWZ_MASK_ROW));
        }
    }
};

// ---------------------------------------------------------------------------------------------
// The pass kernel.
// ---------------------------------------------------------------------------------------------
// Cluster exchange helpers (thread-block clusters + distributed shared memory, sm_90+):
//   cluster_sync()        barrier.cluster arrive(release) + wait(acquire) by every thread of every CTA of the cluster
//   st_cluster(addr, v)   store into the shared memory of CTA `rank` of the cluster (mapa + st.shared::cluster)
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned map_to_rank(unsigned smem_addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(unsigned addr, const double2& v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(unsigned addr, const float2& v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// XCH (cluster exchange role of this pass inside fft_cluster2_kernel, see there):
//   0  plain pass: global memory in, global memory out
//   1  producer (KIND_COL): the LAST stage does not store to global memory; every thread keeps its results in
//      registers across a cluster barrier and then scatters them into the shared-memory tiles of the CTAs that own
//      the rows in the next pass, in the layout [row c'][t] that pass's global loads would have read
//   2  consumer (KIND_TRANS): stage 1 reads that tile instead of global memory (all threads first, then a
//      __syncthreads, then the tile is overwritten in the pass's own layout)
template <typename T, class RL, int C, int NT, int KIND, int XCH = 0, int VARIANT = 0>
struct PassKernel {
    static constexpr int S = RL::S;
    static constexpr int R = RL::R();
    static constexpr int LOG2R = ilog2_c(R);
    static constexpr int R1 = RL::rad(0);          // first-stage radix
    static constexpr int M = R / R1;               // stage-1 tasks per column
    static constexpr int LOG2C = ilog2_c(C);
    using Addr = TileAddr<RL, C, KIND, T>;
    // shared memory: tile (only if S >= 2) + Um[M] + G[C][R1]
    static constexpr int TILE_ELEMS = (S >= 2) ? R * C : 0;
    static constexpr int G_ELEMS = (KIND == KIND_TRANS) ? C * R1 : R1;
    static constexpr size_t SMEM_BYTES = sizeof(cx<T>) * (size_t)(TILE_ELEMS + M + G_ELEMS);
    static_assert(XCH == 0 || S >= 2, "an exchanging pass needs a shared-memory tile");
    static_assert(XCH != 1 || KIND == KIND_COL, "the producer of a cluster exchange is a COL pass");
    static_assert(XCH != 2 || KIND == KIND_TRANS, "the consumer of a cluster exchange is a TRANS pass");

    // ---- global element access -------------------------------------------------------------
    static __device__ __forceinline__ void gload(const PassParams<T>& p, long long idx, T& re, T& im) {
        if (p.in_interleaved) {
            cx<T> v = reinterpret_cast<const cx<T>*>(p.in_re)[idx];
            if (p.in_interleaved == 1) { re = v.x; im = v.y; } else { re = v.y; im = v.x; }
        } else {
            re = p.in_re[idx];
            im = p.in_im[idx];
        }
    }
    static __device__ __forceinline__ void gstore(const PassParams<T>& p, long long idx, T re, T im) {
        if (p.scale != T(1)) { re *= p.scale; im *= p.scale; }
        if (p.out_interleaved == 0) {
            p.out_re[idx] = re;
            p.out_im[idx] = im;
        } else if (p.out_interleaved == 1) {
            reinterpret_cast<cx<T>*>(p.out_re)[idx] = make_cx<T>(re, im);
        } else {
            reinterpret_cast<cx<T>*>(p.out_re)[idx] = make_cx<T>(im, re);
        }
    }

    // N strided elements at once.  The layout test is hoisted out of the unrolled loop so each layout gets
    // its own straight-line block behind a uniform branch: with the test inside the loop the compiler
    // predicates both layouts into one stream (every launch then issues the other layout's dead loads and
    // selects, and the selects of the interleaved path split its loads into two dependent batches).
    // Layout class: 0 planar, 1 interleaved (re, im), 2 interleaved swapped (im, re), -1 = test at run time.
    template <int N, int IL = -1>
    static __device__ __forceinline__ void gload_n(const PassParams<T>& p, long long a0, long long step, T (&re)[N], T (&im)[N]) {
        if constexpr (IL < 0) {
            if (p.in_interleaved == 0) gload_n<N, 0>(p, a0, step, re, im);
            else if (p.in_interleaved == 1) gload_n<N, 1>(p, a0, step, re, im);
            else gload_n<N, 2>(p, a0, step, re, im);
        } else if constexpr (IL == 0) {
            const T* pr = p.in_re + a0;
            const T* pi = p.in_im + a0;
#pragma unroll
            for (int i = 0; i < N; ++i) { re[i] = pr[(long long)i * step]; im[i] = pi[(long long)i * step]; }
        } else {
            const cx<T>* pc = reinterpret_cast<const cx<T>*>(p.in_re) + a0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const cx<T> v = pc[(long long)i * step];
                if constexpr (IL == 1) { re[i] = v.x; im[i] = v.y; } else { re[i] = v.y; im[i] = v.x; }
            }
        }
    }
    template <int N, int OL = -1>
    static __device__ __forceinline__ void gstore_n(const PassParams<T>& p, long long a0, long long step, T (&re)[N], T (&im)[N]) {
        if constexpr (OL < 0) {
            if (p.out_interleaved == 0) gstore_n<N, 0>(p, a0, step, re, im);
            else if (p.out_interleaved == 1) gstore_n<N, 1>(p, a0, step, re, im);
            else gstore_n<N, 2>(p, a0, step, re, im);
        } else {
            if (p.scale != T(1)) {
#pragma unroll
                for (int i = 0; i < N; ++i) { re[i] *= p.scale; im[i] *= p.scale; }
            }
            if constexpr (OL == 0) {
                T* pr = p.out_re + a0;
                T* pi = p.out_im + a0;
#pragma unroll
                for (int i = 0; i < N; ++i) { pr[(long long)i * step] = re[i]; pi[(long long)i * step] = im[i]; }
            } else {
                cx<T>* pc = reinterpret_cast<cx<T>*>(p.out_re) + a0;
#pragma unroll
                for (int i = 0; i < N; ++i) pc[(long long)i * step] = (OL == 1) ? make_cx<T>(re[i], im[i]) : make_cx<T>(im[i], re[i]);
            }
        }
    }

    // OL: output layout class of the LAST stage's stores (see gload_n); the test is made once
    // per stage (run_stages), outside the task loop, so each layout is a straight-line loop.
    template <int s, int OL = -1>
    static __device__ __forceinline__ void stage_from_tile(const PassParams<T>& p, cx<T>* tile, long long out_base,
                                                           long long out_kstride, int tile_rows_valid, int tid) {
        constexpr int RAD = RL::rad(s);
        constexpr int NS = RL::Ns(s);
        constexpr int J = R / RAD;                      // tasks per column
        constexpr int TW_SHIFT = LOG2R - ilog2_c(NS * RAD);
        constexpr bool LAST = (s == S - 1);
        constexpr int NTASK = J * C;
        constexpr int TRIPS = (NTASK + NT - 1) / NT;
#pragma unroll((VARIANT & 1) ? TRIPS : 1)
        for (int t = tid; t < NTASK; t += NT) {
            int c, j;
            if constexpr (KIND == KIND_ROW) { j = t % J; c = t / J; } else { c = t % C; j = t / C; }
            const int m = j & (NS - 1);
            const int g = j / NS;
            const int base = g * NS * RAD + m;
            T xr[RAD], xi[RAD];
#pragma unroll
            for (int i = 0; i < RAD; ++i) {
                cx<T> v = tile[Addr::at(base + i * NS, c)];
                xr[i] = v.x; xi[i] = v.y;
            }
            if constexpr ((VARIANT & 2) && RAD == 8) {
                // 3 table loads (W^m, W^2m, W^4m), the other four twiddles by complex products
                cx<T> w1 = __ldg(p.tw_stage + ((m * 1) << TW_SHIFT));
                cx<T> w2 = __ldg(p.tw_stage + ((m * 2) << TW_SHIFT));
                cx<T> w4 = __ldg(p.tw_stage + ((m * 4) << TW_SHIFT));
                cx<T> w3 = cmul<T>(w1, w2), w5 = cmul<T>(w1, w4), w6 = cmul<T>(w2, w4);
                cx<T> w7 = cmul<T>(w3, w4);
                const cx<T> ws[8] = {w1, w1, w2, w3, w4, w5, w6, w7};
#pragma unroll
                for (int i = 1; i < RAD; ++i) {
                    T a = xr[i], b = xi[i];
                    xr[i] = fma_t(-b, ws[i].y, a * ws[i].x);
                    xi[i] = fma_t(b, ws[i].x, a * ws[i].y);
                }
            } else {
#pragma unroll
                for (int i = 1; i < RAD; ++i) {
                    cx<T> w = __ldg(p.tw_stage + ((m * i) << TW_SHIFT));
                    T a = xr[i], b = xi[i];
                    xr[i] = fma_t(-b, w.y, a * w.x);
                    xi[i] = fma_t(b, w.x, a * w.y);
                }
            }
            Dft<T, RAD>::run(xr, xi);
            if constexpr (!LAST) {
#pragma unroll
                for (int k = 0; k < RAD; ++k) tile[Addr::at(base + k * NS, c)] = make_cx<T>(xr[k], xi[k]);
            } else if constexpr (XCH == 1) {
                // last stage of the producing pass of a cluster exchange.  Output row h = m + k*NS of this CTA's column
                // tcol belongs, in the next pass, to CTA h / CB of the cluster, which wants it at [h % CB][tcol] of its tile
                // (the layout its global loads would have read from the workspace).  All tiles of the cluster are still
                // being read by this very stage, so: results stay in registers, cluster barrier, then the scatter.
                // Lanes run along c, so one store instruction writes 32 consecutive elements of one destination row.
                static_assert(TRIPS == 1 && NTASK == NT, "the exchanging stage must be a single trip: one task per thread");
                cluster_sync();
                const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
                const unsigned tcol = (unsigned)out_base + (unsigned)c;     // body() passes the tile's first column in out_base
                const unsigned cb_mask = (1u << p.xch_log2CB) - 1u;
#pragma unroll
                for (int k = 0; k < RAD; ++k) {
                    const unsigned h = (unsigned)(m + k * NS);
                    const unsigned off = ((h & cb_mask) << p.xch_log2P2) + tcol;
                    st_cluster(map_to_rank(tile_s + off * (unsigned)sizeof(cx<T>), h >> p.xch_log2CB), make_cx<T>(xr[k], xi[k]));
                }
            } else {
                // last stage: g == 0, natural-order outputs kr = m + k*NS
                if (KIND == KIND_ROW && c >= tile_rows_valid) continue;
                if constexpr (KIND == KIND_ROW) gstore_n<RAD, OL>(p, out_base + (long long)c * p.out_bstride + m, (long long)NS, xr, xi);
                else gstore_n<RAD, OL>(p, out_base + (long long)m * out_kstride + c, (long long)NS * out_kstride, xr, xi);
            }
        }
    }

    template <int s>
    static __device__ __forceinline__ void run_stages(const PassParams<T>& p, cx<T>* tile, long long out_base,
                                                      long long out_kstride, int rows_valid, int tid) {
        if constexpr (s < S) {
            __syncthreads();
            if constexpr (s == S - 1) {
                if (p.out_interleaved == 0) stage_from_tile<s, 0>(p, tile, out_base, out_kstride, rows_valid, tid);
                else if (p.out_interleaved == 1) stage_from_tile<s, 1>(p, tile, out_base, out_kstride, rows_valid, tid);
                else stage_from_tile<s, 2>(p, tile, out_base, out_kstride, rows_valid, tid);
            } else {
                stage_from_tile<s>(p, tile, out_base, out_kstride, rows_valid, tid);
            }
            run_stages<s + 1>(p, tile, out_base, out_kstride, rows_valid, tid);
        }
    }

    // ---- the kernel body ------------------------------------------------------------------------
    // `tile_index` is the linear tile id (blockIdx.x for a plain launch; a fused launch loops over tiles).
    // Threads with threadIdx.x >= NT (a fused launch whose other pass needs more threads) take part in
    // the barriers only: their task index starts beyond every task count.
    static __device__ __forceinline__ void body(const PassParams<T>& p, unsigned tile_index) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        cx<T>* tile = reinterpret_cast<cx<T>*>(smem_raw);
        cx<T>* s_um = tile + TILE_ELEMS;   // [M]       per-CTA  W_L^(kp*B*m')
        cx<T>* s_g = s_um + M;             // [C][R1] or [R1]     W_L^(kp(c)*M*B*i)

        const int tid = (threadIdx.x < NT) ? (int)threadIdx.x : (1 << 28);
        // Programmatic dependent launch (sm_90+): this grid may have been scheduled while the previous
        // pass is still draining; wait for its memory to be visible before touching global data, and
        // let the next pass's CTAs be scheduled as soon as every CTA of this grid is resident.
        if (p.pdl) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            asm volatile("griddepcontrol.launch_dependents;");
        }
        long long in_base, out_base, out_kstride;
        long long in_rstride;          // element stride of the tile row index r (COL) / 1 (ROW, TRANS)
        long long in_cstride;          // element stride between tile columns
        int rows_valid = C;
        uint32_t kp0 = 0;              // previous-pass output digit of column c = 0
        uint32_t kp_cstep = 0;         // ... and its increment per column
        uint32_t bcol0 = 0;            // flat index of the remaining digits for column 0 (COL only)

        if constexpr (KIND == KIND_COL) {
            const int tilesB = 1 << (p.log2B - LOG2C);
            const unsigned blk = tile_index + (unsigned)p.blk_offset;
            const int bt = blk & (tilesB - 1);
            const int rest = blk >> (p.log2B - LOG2C);
            const int a = rest & ((1 << p.log2A) - 1);
            const int batch = rest >> p.log2A;
            const long long off = ((long long)a << (LOG2R + p.log2B)) + ((long long)bt << LOG2C);
            in_base = (long long)batch * p.in_bstride + off;
            out_base = (long long)batch * p.out_bstride + off;
            in_rstride = 1LL << p.log2B;
            in_cstride = 1;
            out_kstride = in_rstride;
            kp0 = a & ((1u << p.log2Rprev) - 1u);
            bcol0 = (uint32_t)bt << LOG2C;
            if constexpr (XCH == 1) out_base = (long long)bcol0;   // no global output: the last stage wants the tile's first column
        } else if constexpr (KIND == KIND_TRANS) {
            // rows of the tile: a(c) = (k0 + c) * rest_n + rest, rest_n = A / R1
            const int log2restn = p.log2A - p.log2R1;
            const int tilesK = 1 << p.log2_ktn;
            const int kt = p.kt_base + (tile_index & (tilesK - 1));
            const int tmp = tile_index >> p.log2_ktn;
            const int rest = tmp & ((1 << log2restn) - 1);
            const int batch = tmp >> log2restn;
            const int k0 = kt << LOG2C;
            in_base = (long long)batch * p.in_bstride + ((((long long)k0 << log2restn) + rest) << LOG2R);
            in_rstride = 1;
            in_cstride = 1LL << (log2restn + LOG2R);
            // out[(k0 + c) + R1*rest + A*kr]
            out_base = (long long)batch * p.out_bstride + k0 + ((long long)rest << p.log2R1);
            out_kstride = 1LL << p.log2A;
            const uint32_t rprev_mask = (1u << p.log2Rprev) - 1u;
            kp0 = (uint32_t)(((long long)k0 << log2restn) + rest) & rprev_mask;
            kp_cstep = (log2restn == 0) ? 1u : 0u;   // 2-pass plan: kp = k1 = k0 + c ; 3-pass: kp = k2
        } else {
            const long long first = (long long)tile_index * C;
            rows_valid = (int)min((long long)C, (long long)p.batch - first);
            in_base = first * p.in_bstride;
            out_base = first * p.out_bstride;
            in_rstride = 1;
            in_cstride = p.in_bstride;
            out_kstride = 1;
        }

        // ---- PRELOAD: when every thread owns exactly one stage-1 task, issue its global loads NOW so
        // they are in flight while the twiddle tables below are built (their two-level lookups are two
        // dependent L2 round trips that would otherwise sit in front of the first data load).
        constexpr bool PRELOAD = (M * C <= NT);
        static_assert(XCH != 2 || PRELOAD, "the consumer's stage 1 must be a single trip (every thread holds its task's inputs across the barrier)");
        T pre_r[PRELOAD ? R1 : 1], pre_i[PRELOAD ? R1 : 1];
        if constexpr (PRELOAD) {
            const int t = tid;
            if (t < M * C) {
                int c, mp;
                if constexpr (KIND == KIND_COL) { c = t % C; mp = t / C; } else { mp = t % M; c = t / M; }
                if constexpr (XCH == 2) {
                    // the previous pass's CTAs left this tile as [c][t], t = mp + i*M (what the global loads would have read)
                    const cx<T>* src = tile + c * R + mp;
#pragma unroll
                    for (int i = 0; i < R1; ++i) { const cx<T> v = src[i * M]; pre_r[i] = v.x; pre_i[i] = v.y; }
                } else if ((KIND != KIND_ROW) || (c < rows_valid)) {
                    const long long a0 = in_base + (long long)c * in_cstride + (long long)mp * in_rstride;
                    gload_n<R1>(p, a0, (long long)M * in_rstride, pre_r, pre_i);
                } else {
#pragma unroll
                    for (int i = 0; i < R1; ++i) { pre_r[i] = T(0); pre_i[i] = T(0); }
                }
            }
        }

        // ---- per-CTA inter-pass twiddle factors (two-level lookups, f64, once per CTA) ----------
        // tw(r, c) = W_L^( kp(c) * (r*B + bcol(c)) ),  r = m' + i*M
        //          = W_L^(kp*B*m') * W_L^(kp*bcol) * W_L^(kp*M*B*i)  =  Um[m'] * V[c] * G[c][i]
        cx<T> vreg = make_cx<T>(T(1), T(0));
        const bool has_tw = (KIND != KIND_ROW) && p.has_tw;
        if (has_tw) {
            const int log2B = (KIND == KIND_COL) ? p.log2B : 0;
            for (int mp = tid; mp < M; mp += NT) {
                uint32_t e = (kp0 * (uint32_t)mp) << log2B;
                s_um[mp] = to_cx<T>(p.tw2.get(e << p.tw_shift));
            }
            constexpr int LOG2M = ilog2_c(M);
            for (int q = tid; q < G_ELEMS; q += NT) {
                const int i = q % R1;
                const int c = q / R1;
                uint32_t kp = kp0 + kp_cstep * (uint32_t)c;
                uint32_t e = (kp * (uint32_t)i) << (LOG2M + log2B);
                s_g[q] = to_cx<T>(p.tw2.get(e << p.tw_shift));
            }
            if constexpr (KIND == KIND_COL) {
                const int c = tid % C;   // NT % C == 0: a thread keeps its column for the whole kernel
                uint32_t e = kp0 * (bcol0 + (uint32_t)c);
                vreg = to_cx<T>(p.tw2.get(e << p.tw_shift));
            }
            __syncthreads();
        } else if constexpr (XCH == 2) {
            __syncthreads();     // every thread has read its inputs out of the exchanged tile before stage 1 overwrites it
        }

        // ---- stage 1: global -> registers -> (twiddle, DFT) -> tile (or global when S == 1) ------
        // The input layout is tested once, outside the task loop (see gload_n).
        auto stage1 = [&](auto il_tag) {
            [[maybe_unused]] constexpr int IL = decltype(il_tag)::value;
            constexpr int NTASK = M * C;
            constexpr int TRIPS1 = (NTASK + NT - 1) / NT;
#pragma unroll((VARIANT & 4) ? TRIPS1 : 1)
            for (int t = tid; t < NTASK; t += NT) {
                int c, mp;
                if constexpr (KIND == KIND_COL) { c = t % C; mp = t / C; } else { mp = t % M; c = t / M; }
                T xr[R1], xi[R1];
                [[maybe_unused]] const bool valid = (KIND != KIND_ROW) || (c < rows_valid);
                if constexpr (PRELOAD) {
#pragma unroll
                    for (int i = 0; i < R1; ++i) { xr[i] = pre_r[i]; xi[i] = pre_i[i]; }
                } else if (valid) {
                    const long long a0 = in_base + (long long)c * in_cstride + (long long)mp * in_rstride;
                    gload_n<R1, IL>(p, a0, (long long)M * in_rstride, xr, xi);
                } else {
#pragma unroll
                    for (int i = 0; i < R1; ++i) { xr[i] = T(0); xi[i] = T(0); }
                }
                if (has_tw) {
                    cx<T> pt = s_um[mp];
                    if constexpr (KIND == KIND_COL) {
                        pt = cmul<T>(pt, vreg);
                    } else {
                        if (kp_cstep) pt = cmul<T>(pt, __ldg(p.tw_wc + c * M + mp));
                    }
                    const cx<T>* g = (KIND == KIND_TRANS) ? (s_g + c * R1) : s_g;
                    {
                        T a = xr[0], b = xi[0];
                        xr[0] = fma_t(-b, pt.y, a * pt.x);
                        xi[0] = fma_t(b, pt.x, a * pt.y);
                    }
#pragma unroll
                    for (int i = 1; i < R1; ++i) {
                        cx<T> w = cmul<T>(pt, g[i]);
                        T a = xr[i], b = xi[i];
                        xr[i] = fma_t(-b, w.y, a * w.x);
                        xi[i] = fma_t(b, w.x, a * w.y);
                    }
                }
                Dft<T, R1>::run(xr, xi);
                if constexpr (S >= 2) {
                    const int j = rev_tail<RL>(mp);
#pragma unroll
                    for (int k = 0; k < R1; ++k) tile[Addr::at(j * R1 + k, c)] = make_cx<T>(xr[k], xi[k]);
                } else {
                    if (valid) {
                        if constexpr (KIND == KIND_ROW) gstore_n<R1>(p, out_base + (long long)c * p.out_bstride, 1LL, xr, xi);
                        else gstore_n<R1>(p, out_base + c, out_kstride, xr, xi);
                    }
                }
            }
        };
        if constexpr (PRELOAD) stage1(std::integral_constant<int, 0>{});
        else if (p.in_interleaved == 0) stage1(std::integral_constant<int, 0>{});
        else if (p.in_interleaved == 1) stage1(std::integral_constant<int, 1>{});
        else stage1(std::integral_constant<int, 2>{});
        // ---- stages 2..S -------------------------------------------------------------------------
        run_stages<1>(p, tile, out_base, out_kstride, rows_valid, tid);
    }

};

template <typename T, class RL, int C, int NT, int KIND, int VARIANT = 0, int MINB = 0>
__global__ void __launch_bounds__(NT, MINB) fft_pass_kernel(const __grid_constant__ PassParams<T> p) {
    PassKernel<T, RL, C, NT, KIND, 0, VARIANT>::body(p, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// One HBM pass for transforms that fit the shared memory of a thread-block cluster (2^13..2^17 points): the K CTAs of a
// cluster run pass 1 (PK1, KIND_COL, each CTA P2/K adjacent columns of the P1 x P2 view of the signal) out of global
// memory, exchange the intermediate through distributed shared memory (every CTA scatters its results into the tiles of
// the CTAs that own those rows in pass 2 -- the transpose that a two-launch plan does through a global workspace), and
// run pass 2 (PK2, KIND_TRANS, P1/K rows each) out of shared memory into global memory.  HBM sees the signal once in
// and once out; grid = transforms x K, cluster dimension K (launch attribute).
// Replaces, for these sizes, the L1-resident leaf of the reference's recursion (algorithms/dit.rs:27-93).
// ---------------------------------------------------------------------------------------------
template <class PK1, class PK2, typename T, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) fft_cluster2_kernel(const __grid_constant__ PassParams<T> p1,
                                                                const __grid_constant__ PassParams<T> p2) {
    PK1::body(p1, blockIdx.x);      // ends with: cluster barrier, scatter into the cluster's tiles
    cluster_sync();                 // every CTA's scatter has landed; nobody writes another CTA's tile after this
    PK2::body(p2, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// Fused two-pass launch for signals that live in L2 (N <= 2^20): ONE cooperative grid runs the tiles
// of pass 1, meets at a grid-wide barrier, then runs the tiles of pass 2.  At these sizes a pass is a
// single latency-bound wave and a launch costs ~2.5 us of a ~10 us pass, so removing one launch (and
// the drain/fill between the passes) is worth 10-30 %.
// bar[0] = arrival count, bar[1] = generation; self-resetting, so CUDA-graph replays can reuse it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        volatile unsigned* vgen = bar + 1;
        const unsigned gen = *vgen;
        const unsigned prev = atomicAdd(bar, 1u);
        if (prev == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*vgen == gen) __nanosleep(32);
        }
        __threadfence();
    }
    __syncthreads();
}

template <class PK1, class PK2, typename T, int NTF, int MINB>
__global__ void __launch_bounds__(NTF, MINB) fft_fused2_kernel(const __grid_constant__ PassParams<T> p1,
                                                              const __grid_constant__ PassParams<T> p2,
                                                              unsigned tiles1, unsigned tiles2, unsigned* bar) {
    for (unsigned t = blockIdx.x; t < tiles1; t += gridDim.x) {
        PK1::body(p1, t);
        __syncthreads();
    }
    grid_barrier(bar, gridDim.x);
    for (unsigned t = blockIdx.x; t < tiles2; t += gridDim.x) {
        PK2::body(p2, t);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// r2c post-processing (reference: simd_untangle_inplace_*, algorithms/r2c.rs:150-242) and c2r
// pre-processing (simd_c2r_preprocess_*, r2c.rs:263-432).  Elementwise over bin pairs (k, half-k).
// The planner's w[k] = 0.5 * W_N^k table (planner.rs:120-162, an O(N) rotation recurrence on the
// CPU) is replaced by a two-level lookup: consecutive k hit consecutive `lo` entries (coalesced)
// and one broadcast `hi` entry.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct RealParams {
    T* __restrict__ re;         // untangle: in/out (length half+1).  preprocess: z_re out (length half)
    T* __restrict__ im;
    const T* __restrict__ in_re;  // preprocess only: spectrum (length half+1)
    const T* __restrict__ in_im;
    long long bstride;          // elements between batch members in re/im
    long long in_bstride;
    int log2half;
    Tw2 tw2;                    // two-level table for W_N, N = 2*half
};

template <typename T>
__global__ void __launch_bounds__(256) r2c_untangle_kernel(const __grid_constant__ RealParams<T> p) {
    const long long half = 1LL << p.log2half;
    const long long q = half >> 1;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    T* re = p.re + (long long)blockIdx.y * p.bstride;
    T* im = p.im + (long long)blockIdx.y * p.bstride;
    if (k > q) return;
    if (k == 0) {
        // r2c.rs:161-166
        T a0 = re[0], b0 = im[0];
        re[0] = a0 + b0; im[0] = T(0);
        re[half] = a0 - b0; im[half] = T(0);
        return;
    }
    double2 wd = p.tw2.get((uint32_t)k);
    const T wkr = T(0.5 * wd.x), wki = T(0.5 * wd.y);
    if (k == q) {
        // r2c.rs:233-236 (self pair)
        T a = re[q], b = im[q];
        re[q] = a + T(2) * wkr * b;
        im[q] = T(2) * wki * b;
        return;
    }
    const long long m = half - k;
    T a = re[k], b = im[k], c = re[m], d = im[m];
    T s_re = T(0.5) * (a + c), s_im = T(0.5) * (b - d);
    T t_re = b + d, t_im = c - a;
    T wzr = wkr * t_re - wki * t_im;
    T wzi = wkr * t_im + wki * t_re;
    re[k] = s_re + wzr; im[k] = s_im + wzi;
    re[m] = s_re - wzr; im[m] = wzi - s_im;
}

template <typename T>
__global__ void __launch_bounds__(256) c2r_preprocess_kernel(const __grid_constant__ RealParams<T> p) {
    const long long half = 1LL << p.log2half;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= half) return;
    const T* ire = p.in_re + (long long)blockIdx.y * p.in_bstride;
    const T* iim = p.in_im + (long long)blockIdx.y * p.in_bstride;
    T* zre = p.re + (long long)blockIdx.y * p.bstride;
    T* zim = p.im + (long long)blockIdx.y * p.bstride;
    const long long m = half - k;
    double2 wd = p.tw2.get((uint32_t)k);
    const T c_h = T(0.5 * wd.x), s_h = T(0.5 * wd.y);
    // r2c.rs:263-347
    T re_f = ire[k], im_f = iim[k];
    T re_s = ire[m], im_s = -iim[m];
    T zx_re = T(0.5) * (re_f + re_s), zx_im = T(0.5) * (im_f + im_s);
    T dr = re_f - re_s, di = im_f - im_s;
    T zy_re = c_h * dr + s_h * di;
    T zy_im = c_h * di - s_h * dr;
    zre[k] = zx_re - zy_im;
    zim[k] = zx_im + zy_re;
}

}  // namespace phast
