// Microbenchmarks behind the cluster / DSMEM tile design:
//   1. distributed-shared-memory exchange bandwidth per SM: st.shared::cluster, ld.shared::cluster and
//      cp.async.bulk (shared::cta -> shared::cluster), cluster sizes 2/4/8/16, 8- and 16-byte elements
//   2. whether DSMEM traffic and an HBM stream overlap on the same SM
//   3. cluster.sync cost
//   4. the HBM access pattern of an 8192-row tile spread over a cluster (2^26 = 8192 x 8192 two-pass plan)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dsmem_bench tools/dsmem_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } \
    } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v2f64(unsigned addr, double a, double b) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v2f32(unsigned addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v4f32(unsigned addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ double2 ld_cluster_v2f64(unsigned addr) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// ---------------------------------------------------------------------------------------------
// 1a. st.shared::cluster: every thread owns 16 elements per trip (a radix-16 task's outputs); output k goes to
// CTA (k * K / 16) like the exchange stage of the cluster FFT; lanes write consecutive elements.
// ELEM = 8 (float2) or 16 (double2).  MODE 0: all destinations remote-or-local as the FFT does; 1: local only.
// ---------------------------------------------------------------------------------------------
template <int K, int ELEM, int NT, int MODE>
__global__ void __launch_bounds__(NT) k_dsmem_st(int iters, long long* cycles_out, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned rank = cluster_rank();
    const unsigned base = smem_u32(smem);
    constexpr int BUF_ELEMS = NT * 16;   // per CTA: NT tasks x 16 outputs
    cluster_sync_all();
    long long t0 = clock64();
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned dst_rank = MODE == 1 ? rank : (unsigned)(k * K / 16);
            // slot inside the destination buffer: [src_rank-local k][tid]
            const unsigned slot = ((unsigned)((k % (16 / K)) + (16 / K) * (MODE == 1 ? 0 : rank)) % 16) * NT + threadIdx.x;
            const unsigned a = mapa(base + slot * ELEM, dst_rank);
            if (ELEM == 16) st_cluster_v2f64(a, (double)it, (double)k);
            else st_cluster_v2f32(a, (float)it, (float)k);
        }
    }
    cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (sink && threadIdx.x == 0) sink[blockIdx.x] = acc + reinterpret_cast<double*>(smem)[threadIdx.x % BUF_ELEMS];
}

// 1b. ld.shared::cluster, 16-byte elements (pull model)
template <int K, int NT>
__global__ void __launch_bounds__(NT) k_dsmem_ld(int iters, long long* cycles_out, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned rank = cluster_rank();
    const unsigned base = smem_u32(smem);
    for (int i = threadIdx.x; i < NT * 16 * 2; i += NT) reinterpret_cast<double*>(smem)[i] = i;
    cluster_sync_all();
    long long t0 = clock64();
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        double2 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned src_rank = (unsigned)(k * K / 16);
            const unsigned slot = ((unsigned)((k % (16 / K)) + (16 / K) * rank) % 16) * NT + threadIdx.x;
            v[k] = ld_cluster_v2f64(mapa(base + slot * 16, src_rank));
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += v[k].x + v[k].y;
    }
    cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (sink) sink[blockIdx.x * NT + threadIdx.x] = acc;
}

// 1c. cp.async.bulk shared::cta -> shared::cluster with mbarrier complete_tx on the destination CTA.
// Every CTA sends CHUNK-byte pieces of its 64 KB send buffer round-robin to all K CTAs (itself included).
template <int K, int NT>
__global__ void __launch_bounds__(NT) k_dsmem_bulk(int iters, int chunk, long long* cycles_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BUF = 64 * 1024;
    unsigned char* send = smem;
    unsigned char* recv = smem + BUF;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(smem + 2 * BUF);
    const unsigned rank = cluster_rank();
    const unsigned mb = smem_u32(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    long long t0 = clock64();
    const int nchunks = BUF / chunk;
    unsigned phase = 0;
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == 0) {
            // this CTA will receive BUF bytes in total (BUF/K from each of K senders)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)BUF) : "memory");
        }
        // issue: warp w lane 0 sends chunks w, w+nwarps, ...
        const int warp = threadIdx.x / 32, nw = NT / 32;
        if ((threadIdx.x & 31) == 0) {
            for (int c = warp; c < nchunks; c += nw) {
                const unsigned dst_rank = (unsigned)(c % K);
                // destination offset: sender `rank` owns the slice [rank*BUF/K, (rank+1)*BUF/K) of every receiver
                const unsigned dst_off = rank * (BUF / K) + (unsigned)(c / K) * chunk;
                const unsigned src = smem_u32(send + (size_t)c * chunk);
                const unsigned dst = mapa(smem_u32(recv + dst_off), dst_rank);
                const unsigned rmb = mapa(mb, dst_rank);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "r"(src), "r"((unsigned)chunk), "r"(rmb) : "memory");
            }
        }
        // wait for this CTA's receive buffer to fill
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mb), "r"(phase) : "memory");
        }
        phase ^= 1;
        // before reusing anyone's receive buffer all CTAs must have seen theirs complete
        cluster_sync_all();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
}

// 3. cluster.sync cost
template <int K>
__global__ void __launch_bounds__(256) k_csync(int iters, long long* cycles_out) {
    cluster_sync_all();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
}

// 2. overlap: warps [0, NT/2) stream a big global array (HBM), warps [NT/2, NT) do DSMEM stores.
// which: 1 = stream only, 2 = dsmem only, 3 = both.
template <int K, int NT>
__global__ void __launch_bounds__(NT) k_overlap(int which, int iters, const double2* __restrict__ g, long long n2, long long* cycles_out,
                                                double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned rank = cluster_rank();
    const unsigned base = smem_u32(smem);
    cluster_sync_all();
    long long t0 = clock64();
    double acc = 0;
    const int half = NT / 2;
    if (threadIdx.x < half) {
        if (which & 1) {
            // each CTA streams iters * half * 16 elements of 16 bytes
            long long idx = ((long long)blockIdx.x * iters) * half * 16 + threadIdx.x;
            for (int it = 0; it < iters; ++it) {
                double2 v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __ldcs(g + ((idx + (long long)k * half) % n2));
#pragma unroll
                for (int k = 0; k < 16; ++k) acc += v[k].x + v[k].y;
                idx += (long long)half * 16;
            }
        }
    } else {
        if (which & 2) {
            const int t = threadIdx.x - half;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const unsigned dst_rank = (unsigned)(k * K / 16);
                    const unsigned slot = ((unsigned)((k % (16 / K)) + (16 / K) * rank) % 16) * half + t;
                    st_cluster_v2f64(mapa(base + slot * 16, dst_rank), (double)it, (double)k);
                }
            }
        }
    }
    cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (sink) sink[(long long)blockIdx.x * NT + threadIdx.x] = acc;
}

// 4. tile copy: a (RCTA*K)-row x C-column tile of 8-byte (planar re + im) or 16-byte (interleaved) elements spread over a
// K-CTA cluster by rows; row stride B elements; adjacent clusters take adjacent C-runs.  out may use the same pattern
// (OUT_BLOCKED = 0) or write each CTA's RCTA x C piece contiguously (OUT_BLOCKED = 1: the "blocked workspace").
template <int RCTA, int C, int K, int NT, int ELEM, int OUT_BLOCKED>
__global__ void __launch_bounds__(NT) k_tile_copy(const double* __restrict__ in_re, const double* __restrict__ in_im, double* __restrict__ out_re,
                                                  double* __restrict__ out_im, int log2B) {
    const long long B = 1LL << log2B;
    const long long tilesB = B / C;
    const unsigned cl = blockIdx.x / K, q = blockIdx.x % K;
    const long long bt = cl % tilesB, a = cl / tilesB;
    const long long base = a * (long long)(RCTA * K) * B + (long long)q * RCTA * B + bt * C;
    constexpr int VEC = (ELEM == 16) ? 1 : 2;                  // 16 bytes per thread access either way
    constexpr int CV = C / VEC;
    constexpr int ROWS_PER_IT = NT / CV;
    constexpr int ITERS = RCTA / ROWS_PER_IT;
    const int c = (threadIdx.x % CV) * VEC, r0 = threadIdx.x / CV;
    constexpr int U = ITERS < 8 ? ITERS : 8;
    const double2* in2 = reinterpret_cast<const double2*>(in_re);
    double2* out2 = reinterpret_cast<double2*>(out_re);
    for (int it0 = 0; it0 < ITERS; it0 += U) {
        double2 vr[U], vi[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long idx = base + (long long)(r0 + (it0 + u) * ROWS_PER_IT) * B + c;
            if (ELEM == 16) vr[u] = in2[idx];
            else { vr[u] = *reinterpret_cast<const double2*>(in_re + idx); vi[u] = *reinterpret_cast<const double2*>(in_im + idx); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long idx = base + (long long)(r0 + (it0 + u) * ROWS_PER_IT) * B + c;
            if (OUT_BLOCKED) idx = (long long)blockIdx.x * RCTA * C + (long long)(r0 + (it0 + u) * ROWS_PER_IT) * C + c;
            if (ELEM == 16) out2[idx] = vr[u];
            else { *reinterpret_cast<double2*>(out_re + idx) = vr[u]; *reinterpret_cast<double2*>(out_im + idx) = vi[u]; }
        }
    }
}


// 1d. the same exchange through global memory (L2-resident scratch): st.global to the destination CTA's region,
// cluster barrier, ld.global.cg of the own region.  Counts the bytes once (like the DSMEM rows).
template <int K, int NT>
__global__ void __launch_bounds__(NT) k_l2_exchange(int iters, double2* scratch, long long* cycles_out, double* sink) {
    const unsigned rank = cluster_rank();
    const unsigned cl = blockIdx.x / K;
    double2* mine = scratch + ((size_t)cl * K + rank) * NT * 16;
    cluster_sync_all();
    long long t0 = clock64();
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const unsigned dst_rank = (unsigned)(k * K / 16);
            const unsigned slot = ((unsigned)((k % (16 / K)) + (16 / K) * rank) % 16) * NT + threadIdx.x;
            __stcg(scratch + ((size_t)cl * K + dst_rank) * NT * 16 + slot, make_double2((double)it, (double)k));
        }
        cluster_sync_all();
        double2 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = __ldcg(mine + k * NT + threadIdx.x);
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += v[k].x + v[k].y;
        cluster_sync_all();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (sink) sink[(long long)blockIdx.x * NT + threadIdx.x] = acc;
}


// 5. L2 bandwidth: grid-stride copy / read of a buffer that fits the 126 MB L2, repeated (first repetition warms L2)
__global__ void __launch_bounds__(512) k_l2_copy(const double2* __restrict__ in, double2* __restrict__ out, long long n2, int reps) {
    for (int r = 0; r < reps; ++r)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) out[i] = __ldcg(in + i);
}
__global__ void __launch_bounds__(512) k_l2_read(const double2* __restrict__ in, double* sink, long long n2, int reps) {
    double acc = 0;
    for (int r = 0; r < reps; ++r) {
        long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        for (; i + 3LL * gridDim.x * blockDim.x < n2; i += 4LL * gridDim.x * blockDim.x) {
            double2 a = __ldcg(in + i), b = __ldcg(in + i + (long long)gridDim.x * blockDim.x), c = __ldcg(in + i + 2LL * gridDim.x * blockDim.x),
                    d = __ldcg(in + i + 3LL * gridDim.x * blockDim.x);
            acc += a.x + b.y + c.x + d.y;
        }
    }
    if (acc == 1.2345) sink[0] = acc;
}
void bench_l2(size_t mib) {
    const long long n2 = (long long)(mib << 20) / 16;
    double2 *a, *b; double* sink;
    CK(cudaMalloc(&a, n2 * 16)); CK(cudaMalloc(&b, n2 * 16)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(a, 0, n2 * 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    for (int mode = 0; mode < 2; ++mode) {
        for (int w = 0; w < 2; ++w) { if (mode) k_l2_read<<<148 * 4, 512>>>(a, sink, n2, 2); else k_l2_copy<<<148 * 4, 512>>>(a, b, n2, 2); }
        cudaEventRecord(e0);
        if (mode) k_l2_read<<<148 * 4, 512>>>(a, sink, n2, reps); else k_l2_copy<<<148 * 4, 512>>>(a, b, n2, reps);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)reps * n2 * 16 * (mode ? 1 : 2);
        printf("L2 %s of %4zu MiB %s: %7.2f TB/s\n", mode ? "read" : "copy", mib, mode ? "            " : "(read+write)", bytes / (ms * 1e-3) / 1e12);
    }
    CK(cudaFree(a)); CK(cudaFree(b)); CK(cudaFree(sink));
}


// 6. Does an L2-sized ring of intermediates save the HBM round trip?  Persistent CTAs; work item i: copy tile i of `in` into
// ring slot (i % ring_tiles) of `ws`, and copy the tile written `delay` items ago from the ring to `out`.  With ring = the whole
// array this is two full HBM copies (what a two-launch plan moves); with an L2-sized ring the middle write+read should stay in L2.
// hint: 0 plain ld/st, 1 streaming (.cs) for in/out and .cg for the ring
__global__ void __launch_bounds__(256) k_ring(const double2* __restrict__ in, double2* __restrict__ out, double2* ws, long long tiles, int tile_elems,
                                              long long ring_tiles, int delay, int hint, unsigned long long* ticket) {
    __shared__ long long s_item;
    for (;;) {
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(ticket, 1ULL);
        __syncthreads();
        const long long item = s_item;
        __syncthreads();
        if (item >= tiles + delay) break;
        if (item < tiles) {
            const double2* src = in + item * tile_elems;
            double2* dst = ws + (item % ring_tiles) * tile_elems;
            for (int e = threadIdx.x; e < tile_elems; e += 4 * blockDim.x) {
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = hint ? __ldcs(src + e + u * blockDim.x) : src[e + u * blockDim.x];
#pragma unroll
                for (int u = 0; u < 4; ++u) { if (hint) __stcg(dst + e + u * blockDim.x, v[u]); else dst[e + u * blockDim.x] = v[u]; }
            }
        }
        if (item >= delay) {
            const long long j = item - delay;
            const double2* src = ws + (j % ring_tiles) * tile_elems;
            double2* dst = out + j * tile_elems;
            for (int e = threadIdx.x; e < tile_elems; e += 4 * blockDim.x) {
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = hint ? __ldcg(src + e + u * blockDim.x) : src[e + u * blockDim.x];
#pragma unroll
                for (int u = 0; u < 4; ++u) { if (hint) __stcs(dst + e + u * blockDim.x, v[u]); else dst[e + u * blockDim.x] = v[u]; }
            }
        }
    }
}
void bench_ring() {
    const long long total = 2048LL << 20;                 // 2 GiB in, 2 GiB out
    const int tile_bytes = 32 * 1024, tile_elems = tile_bytes / 16;
    const long long tiles = total / tile_bytes;
    double2 *in, *out, *ws; unsigned long long* ticket;
    CK(cudaMalloc(&in, total)); CK(cudaMalloc(&out, total)); CK(cudaMalloc(&ws, total)); CK(cudaMalloc(&ticket, 8));
    CK(cudaMemset(in, 0, total));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int hint = 0; hint < 2; ++hint)
        for (long long ring_mb : {2048LL, 96LL, 64LL, 48LL, 32LL, 16LL}) {
            const long long ring_tiles = (ring_mb << 20) / tile_bytes;
            const int delay = (int)std::min<long long>(ring_tiles / 2, tiles / 2);     // read back half a ring later
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(ticket, 0, 8));
                cudaEventRecord(e0);
                k_ring<<<148 * 4, 256>>>(in, out, ws, tiles, tile_elems, ring_tiles, delay, hint, ticket);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                best = std::min(best, ms);
            }
            printf("ring pipeline hint=%d ring=%5lld MiB delay=%6d tiles: %8.1f us -> %5.2f TB/s of compulsory traffic (2 GiB in + 2 GiB out)\n", hint, ring_mb, delay,
                   best * 1e3, 2.0 * total / (best * 1e-3) / 1e12);
        }
    CK(cudaFree(in)); CK(cudaFree(out)); CK(cudaFree(ws)); CK(cudaFree(ticket));
}

template <typename F>
void launch_cluster(F kernel, int grid, int nt, size_t smem, int K, void** args) {
    if (smem > 227 * 1024) { printf("  (skipped: %zu bytes of shared memory)\n", smem); return; }
    CK(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (K > 8) CK(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = K; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelExC(&cfg, (const void*)kernel, args));
}

long long* d_cycles;
double* d_sink;
double max_cycles(int grid) {
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long m = 0;
    for (auto v : h) m = v > m ? v : m;
    return (double)m;
}

template <int K, int ELEM, int NT, int MODE>
void bench_st(int clusters) {
    int iters = 200;
    double* nul = nullptr;
    void* args[] = {&iters, &d_cycles, &nul};
    const size_t smem = (size_t)NT * 16 * ELEM;
    const int grid = clusters * K;
    launch_cluster(k_dsmem_st<K, ELEM, NT, MODE>, grid, NT, smem, K, args);
    launch_cluster(k_dsmem_st<K, ELEM, NT, MODE>, grid, NT, smem, K, args);
    CK(cudaDeviceSynchronize());
    const double cyc = max_cycles(grid);
    const double bytes = (double)iters * NT * 16 * ELEM;
    printf("st.shared::cluster  K=%2d elem=%2dB NT=%4d %s grid=%4d: %7.1f B/clk/SM stored (%.0f%% remote)\n", K, ELEM, NT,
           MODE == 1 ? "LOCAL " : "fftmap", grid, bytes / cyc, MODE == 1 ? 0.0 : 100.0 * (K - 1) / K);
}
template <int K, int NT>
void bench_ld(int clusters) {
    int iters = 200;
    double* nul = nullptr;
    void* args[] = {&iters, &d_cycles, &nul};
    const size_t smem = (size_t)NT * 16 * 16 * 2;
    const int grid = clusters * K;
    launch_cluster(k_dsmem_ld<K, NT>, grid, NT, smem, K, args);
    launch_cluster(k_dsmem_ld<K, NT>, grid, NT, smem, K, args);
    CK(cudaDeviceSynchronize());
    const double cyc = max_cycles(grid);
    printf("ld.shared::cluster  K=%2d elem=16B NT=%4d fftmap grid=%4d: %7.1f B/clk/SM loaded\n", K, NT, grid, (double)iters * NT * 16 * 16 / cyc);
}
template <int K, int NT>
void bench_bulk(int clusters, int chunk) {
    int iters = 50;
    void* args[] = {&iters, &chunk, &d_cycles};
    const size_t smem = 2 * 64 * 1024 + 64;
    const int grid = clusters * K;
    launch_cluster(k_dsmem_bulk<K, NT>, grid, NT, smem, K, args);
    launch_cluster(k_dsmem_bulk<K, NT>, grid, NT, smem, K, args);
    CK(cudaDeviceSynchronize());
    const double cyc = max_cycles(grid);
    printf("cp.async.bulk->DSMEM K=%2d chunk=%5dB NT=%4d grid=%4d: %7.1f B/clk/SM sent (incl. 1 cluster.sync per 64 KB)\n", K, chunk, NT, grid,
           (double)iters * 64 * 1024 / cyc);
}

template <int K, int NT>
void bench_l2x(int clusters) {
    int iters = 100;
    double* nul = nullptr;
    const int grid = clusters * K;
    double2* scratch;
    CK(cudaMalloc(&scratch, (size_t)grid * NT * 16 * sizeof(double2)));
    void* args[] = {&iters, &scratch, &d_cycles, &nul};
    launch_cluster(k_l2_exchange<K, NT>, grid, NT, 0, K, args);
    launch_cluster(k_l2_exchange<K, NT>, grid, NT, 0, K, args);
    CK(cudaDeviceSynchronize());
    const double cyc = max_cycles(grid);
    printf("L2 exchange (st.cg + 2 cluster.sync + ld.cg) K=%2d NT=%4d grid=%4d: %7.1f B/clk/SM exchanged (each byte written once, read once)\n", K, NT, grid,
           (double)iters * NT * 16 * 16 / cyc);
    CK(cudaFree(scratch));
}

template <int K>
void bench_csync(int clusters) {
    int iters = 200;
    void* args[] = {&iters, &d_cycles};
    const int grid = clusters * K;
    launch_cluster(k_csync<K>, grid, 256, 0, K, args);
    launch_cluster(k_csync<K>, grid, 256, 0, K, args);
    CK(cudaDeviceSynchronize());
    printf("cluster.sync        K=%2d grid=%4d: %7.1f cycles each\n", K, grid, max_cycles(grid) / iters);
}
template <int K, int NT>
void bench_overlap(int clusters, const double2* g, long long n2) {
    const int grid = clusters * K;
    double res[4] = {0, 0, 0, 0};
    for (int which = 1; which <= 3; ++which) {
        int iters = 100;
        double* nul = nullptr;
        void* args[] = {&which, &iters, &g, &n2, &d_cycles, &nul};
        const size_t smem = (size_t)(NT / 2) * 16 * 16;
        launch_cluster(k_overlap<K, NT>, grid, NT, smem, K, args);
        launch_cluster(k_overlap<K, NT>, grid, NT, smem, K, args);
        CK(cudaDeviceSynchronize());
        res[which] = max_cycles(grid);
    }
    const double bytes = 100.0 * (NT / 2) * 16 * 16;
    printf("overlap K=%2d NT=%4d grid=%4d: HBM stream alone %6.1f B/clk/SM, DSMEM alone %6.1f, together %6.1f + %6.1f (cycles %0.f / %0.f / %0.f)\n", K, NT,
           grid, bytes / res[1], bytes / res[2], bytes / res[3], bytes / res[3], res[1], res[2], res[3]);
}

template <int RCTA, int C, int K, int NT, int ELEM, int OUT_BLOCKED>
void bench_tile(const char* name, int log2n, double* a, double* b, double* c, double* d) {
    const long long n = 1LL << log2n;   // complex points
    int log2R = 0;
    while ((1 << log2R) < RCTA * K) ++log2R;
    const int log2B = log2n - log2R;
    const long long blocks = n / RCTA / C;
    int lb = log2B;
    void* args[] = {&a, &b, &c, &d, &lb};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto kern = k_tile_copy<RCTA, C, K, NT, ELEM, OUT_BLOCKED>;
    for (int w = 0; w < 2; ++w) launch_cluster(kern, (int)blocks, NT, 0, K, args);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int w = 0; w < reps; ++w) launch_cluster(kern, (int)blocks, NT, 0, K, args);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    printf("tile copy %-22s rows=%5d (K=%2d x %4d) C=%2d elem=%2dB run=%4dB out=%s: %8.1f us %6.2f TB/s\n", name, RCTA * K, K, RCTA, C, ELEM, C * ELEM,
           OUT_BLOCKED ? "blocked" : "same   ", ms * 1e3, 32.0 * n / (ms * 1e-3) / 1e12);
}

int main(int argc, char** argv) {
    CK(cudaMalloc(&d_cycles, 4096 * sizeof(long long)));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("SMs %d, clock %d kHz\n", sms, clk);

    if (argc > 1 && std::string(argv[1]) == "ring") { bench_ring(); return 0; }
    if (argc > 1 && std::string(argv[1]) == "l2") { for (size_t m : {8, 16, 32, 48, 64, 256, 1024}) bench_l2(m); return 0; }
    // ---- 3. cluster.sync ----
    bench_csync<2>(64); bench_csync<4>(32); bench_csync<8>(16); bench_csync<16>(8);

    // ---- 1. DSMEM bandwidth: one cluster alone, then the chip full of clusters ----
    bench_st<4, 16, 256, 1>(1);
    bench_st<2, 16, 256, 0>(1); bench_st<4, 16, 256, 0>(1); bench_st<8, 16, 256, 0>(1); bench_st<16, 16, 256, 0>(1);
    bench_st<4, 16, 512, 0>(1); bench_st<8, 16, 512, 0>(1);
    bench_st<4, 8, 512, 0>(1); bench_st<8, 8, 512, 0>(1); bench_st<4, 8, 1024, 0>(1);
    bench_st<2, 16, 256, 0>(74); bench_st<4, 16, 256, 0>(37); bench_st<8, 16, 256, 0>(18); bench_st<16, 16, 256, 0>(8);
    bench_st<4, 16, 512, 0>(37); bench_st<8, 16, 512, 0>(18);
    bench_st<4, 8, 512, 0>(37); bench_st<8, 8, 512, 0>(18); bench_st<4, 8, 1024, 0>(37);
    bench_st<4, 16, 768, 0>(37); bench_st<8, 16, 768, 0>(18);
    bench_l2x<4, 256>(37); bench_l2x<8, 256>(18); bench_l2x<8, 512>(18); bench_l2x<8, 1024>(18);
    for (int chunk : {1024, 4096}) { bench_bulk<4, 128>(37, chunk); bench_bulk<8, 128>(18, chunk); }
    bench_bulk<2, 128>(74, 4096); bench_bulk<16, 128>(8, 4096);

    // ---- 2. overlap with an HBM stream ----
    {
        const long long n2 = 1LL << 27;   // 2 GiB of double2
        double2* g;
        CK(cudaMalloc(&g, n2 * sizeof(double2)));
        CK(cudaMemset(g, 0, n2 * sizeof(double2)));
        bench_overlap<4, 512>(37, g, n2);
        bench_overlap<8, 512>(18, g, n2);
        bench_overlap<4, 1024>(37, g, n2);
        CK(cudaFree(g));
    }

    // ---- 4. 8192-row tiles over a cluster, 2^26 complex f64 points (1 GiB in, 1 GiB out) ----
    {
        const int log2n = 26;
        const long long n = 1LL << log2n;
        double *a, *b, *c, *d;
        CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8)); CK(cudaMalloc(&c, n * 8)); CK(cudaMalloc(&d, n * 8));
        CK(cudaMemset(a, 0, n * 8)); CK(cudaMemset(b, 0, n * 8));
        // planar (8-byte elements, re + im): 64 B, 128 B runs
        bench_tile<1024, 8, 8, 256, 8, 0>("planar 8192x8", log2n, a, b, c, d);
        bench_tile<1024, 8, 8, 256, 8, 1>("planar 8192x8", log2n, a, b, c, d);
        bench_tile<512, 16, 16, 256, 8, 0>("planar 8192x16", log2n, a, b, c, d);
        bench_tile<512, 16, 16, 256, 8, 1>("planar 8192x16", log2n, a, b, c, d);
        bench_tile<2048, 4, 4, 256, 8, 1>("planar 8192x4", log2n, a, b, c, d);
        bench_tile<512, 8, 8, 256, 8, 1>("planar 4096x8", log2n, a, b, c, d);
        bench_tile<256, 16, 16, 256, 8, 1>("planar 4096x16", log2n, a, b, c, d);
        bench_tile<512, 16, 8, 256, 8, 1>("planar 4096x16", log2n, a, b, c, d);
        // interleaved (16-byte elements), a and c viewed as n double2 each (uses a..b and c..d contiguous? no: separate) -> use n/2 points
        bench_tile<1024, 4, 8, 256, 16, 0>("interleaved 8192x4", log2n - 1, a, b, c, d);
        bench_tile<1024, 4, 8, 256, 16, 1>("interleaved 8192x4", log2n - 1, a, b, c, d);
        bench_tile<1024, 8, 8, 256, 16, 1>("interleaved 8192x8", log2n - 1, a, b, c, d);
        // one CTA per tile, for comparison (the round-1 geometry): 256 rows x 16 columns, rows 2^18 elements apart
        bench_tile<256, 16, 1, 256, 8, 0>("planar 256x16 (r1)", log2n, a, b, c, d);
        bench_tile<1024, 8, 1, 256, 8, 0>("planar 1024x8 (r1)", log2n, a, b, c, d);
    }
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
