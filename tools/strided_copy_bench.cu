// Microbenchmark: the HBM access pattern of a COL pass (R rows x C doubles per tile, row stride B
// elements, adjacent CTAs take adjacent C-runs) as a pure copy, to find the pattern's own ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/strided_copy tools/strided_copy_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int R, int C, int NT, int VEC>
__global__ void __launch_bounds__(NT) copy_tile(const double* __restrict__ in_re, const double* __restrict__ in_im,
                                                double* __restrict__ out_re, double* __restrict__ out_im, int log2B) {
    // tile = all R rows (stride B) x C columns; VEC doubles per thread-load (1 or 2)
    const long long B = 1LL << log2B;
    const long long tilesB = B / C;
    const long long bt = blockIdx.x % tilesB, a = blockIdx.x / tilesB;
    const long long base = a * R * B + bt * C;
    constexpr int CV = C / VEC;
    constexpr int ROWS_PER_IT = NT / CV;
    constexpr int ITERS = R / ROWS_PER_IT;
    const int c = (threadIdx.x % CV) * VEC, r0 = threadIdx.x / CV;
    constexpr int U = ITERS < 8 ? ITERS : 8;
    for (int it0 = 0; it0 < ITERS; it0 += U) {
        double vr[U][VEC], vi[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long idx = base + (long long)(r0 + (it0 + u) * ROWS_PER_IT) * B + c;
            if (VEC == 2) {
                double2 x = *reinterpret_cast<const double2*>(in_re + idx), y = *reinterpret_cast<const double2*>(in_im + idx);
                vr[u][0] = x.x; vr[u][VEC - 1] = x.y; vi[u][0] = y.x; vi[u][VEC - 1] = y.y;
            } else { vr[u][0] = in_re[idx]; vi[u][0] = in_im[idx]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            long long idx = base + (long long)(r0 + (it0 + u) * ROWS_PER_IT) * B + c;
            if (VEC == 2) {
                *reinterpret_cast<double2*>(out_re + idx) = make_double2(vr[u][0], vr[u][VEC - 1]);
                *reinterpret_cast<double2*>(out_im + idx) = make_double2(vi[u][0], vi[u][VEC - 1]);
            } else { out_re[idx] = vr[u][0]; out_im[idx] = vi[u][0]; }
        }
    }
}

template <int R, int C, int NT, int VEC>
void run(const char* name, int log2n, double* a, double* b, double* c, double* d) {
    const long long n = 1LL << log2n;
    int log2R = 0; while ((1 << log2R) < R) ++log2R;
    for (int log2A = 0; log2A <= 9; log2A += 9) {   // A = 1 (first pass) and A = 512 (middle pass)
        int log2B = log2n - log2R - log2A;
        if (log2B < 5) continue;
        long long blocks = n / R / C;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int w = 0; w < 2; ++w) copy_tile<R, C, NT, VEC><<<(unsigned)blocks, NT>>>(a, b, c, d, log2B);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int w = 0; w < reps; ++w) copy_tile<R, C, NT, VEC><<<(unsigned)blocks, NT>>>(a, b, c, d, log2B);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("%-28s R=%4d C=%2d (%3d B runs) NT=%3d log2B=%2d: %8.1f us  %6.2f TB/s  (%s)\n", name, R, C, C * 8, NT, log2B, ms * 1e3,
               4.0 * n * 8 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    }
}

__global__ void plain_copy(const double4* __restrict__ in, double4* __restrict__ out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
    const int log2n = 26;
    const long long n = 1LL << log2n;
    double *a, *b, *c, *d;
    cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8); cudaMalloc(&c, n * 8); cudaMalloc(&d, n * 8);
    cudaMemset(a, 0, n * 8); cudaMemset(b, 0, n * 8);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        plain_copy<<<148 * 8, 512>>>((double4*)a, (double4*)c, n / 4);
        cudaEventRecord(e0);
        for (int w = 0; w < 5; ++w) { plain_copy<<<148 * 8, 512>>>((double4*)a, (double4*)c, n / 4); plain_copy<<<148 * 8, 512>>>((double4*)b, (double4*)d, n / 4); }
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("plain contiguous copy of the same 2 GiB: %8.1f us %6.2f TB/s\n", ms * 1e3, 4.0 * n * 8 / (ms * 1e-3) / 1e12);
    }
    run<512, 4, 256, 1>("col 32B", log2n, a, b, c, d);
    run<512, 8, 256, 1>("col 64B", log2n, a, b, c, d);
    run<512, 8, 256, 2>("col 64B vec2", log2n, a, b, c, d);
    run<512, 16, 256, 1>("col 128B", log2n, a, b, c, d);
    run<512, 16, 256, 2>("col 128B vec2", log2n, a, b, c, d);
    run<512, 32, 256, 2>("col 256B vec2", log2n, a, b, c, d);
    run<256, 8, 256, 1>("col 64B", log2n, a, b, c, d);
    run<256, 16, 256, 1>("col 128B", log2n, a, b, c, d);
    run<256, 32, 256, 2>("col 256B vec2", log2n, a, b, c, d);
    run<1024, 8, 256, 1>("col 64B", log2n, a, b, c, d);
    run<1024, 16, 256, 2>("col 128B vec2", log2n, a, b, c, d);
    run<64, 8, 64, 1>("col 64B small", log2n, a, b, c, d);
    run<64, 16, 128, 1>("col 128B small", log2n, a, b, c, d);
    run<64, 64, 256, 2>("col 512B small", log2n, a, b, c, d);
    return 0;
}
