// Micro-benchmarks that ground the design of the pencil sweep (B200, sm_100a): dependent fp64 op latencies,
// shared-memory round trips, CTA barrier cost, L2 load latency and the SM->L2->SM hand-off latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat lat.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_dadd(double *out, double a, double b, int n, long long *cyc)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = x - b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dmul(double *out, double a, double b, int n, long long *cyc)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = x * b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dmuladd(double *out, double a, double b, double c, int n, long long *cyc)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = c - b * x;   // -fmad=false: DMUL then DADD
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma(double *out, double a, double b, double c, int n, long long *cyc)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = fma(b, x, c);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ddiv(double *out, double a, double b, int n, long long *cyc)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = x / b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// Markstein-style division with a precomputed correctly rounded reciprocal
__global__ void k_ddiv_mark(double *out, double a, double b, int n, long long *cyc)
{
    double x = a;
    const double y = 1.0 / b;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            double q = x * y;
            double r = fma(-b, q, x);
            q = fma(r, y, q);
            r = fma(-b, q, x);
            x = fma(r, y, q);
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// shared memory: dependent LDS chain (pointer chasing)
__global__ void k_lds(int *out, int n, long long *cyc)
{
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) p = s[p];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p; if (threadIdx.x == 0) *cyc = t1 - t0;
}
// STS -> bar.sync -> LDS -> (dependent) loop: what one wavefront step costs without arithmetic
__global__ void k_step(double *out, int n, long long *cyc)
{
    __shared__ double s[2][1024];
    const int t = threadIdx.x, T = blockDim.x;
    s[0][t] = t; s[1][t] = 0;
    __syncthreads();
    double x = 1.0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        const double o = s[i & 1][(t + T - 1) % T];
        x = x - 0.5 * o;
        s[(i + 1) & 1][t] = x;
        __syncthreads();
    }
    long long t1 = clock64();
    out[t] = x; if (t == 0) *cyc = t1 - t0;
}
// the same with three operands and three dependent subtractions (an ILU(0) row) and optional division
template <int DIV>
__global__ void k_step3(double *out, int n, double d, long long *cyc)
{
    __shared__ double s[2][1024];
    const int t = threadIdx.x, T = blockDim.x;
    s[0][t] = t; s[1][t] = 0;
    __syncthreads();
    double x = 1.0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        const double a = s[i & 1][(t + T - 16) % T], b = s[i & 1][(t + T - 1) % T], c = s[i & 1][t];
        x = 1.0 - 0.25 * a;
        x = x - 0.25 * b;
        x = x - 0.25 * c;
        if (DIV) x = x / d;
        s[(i + 1) & 1][t] = x;
        __syncthreads();
    }
    long long t1 = clock64();
    out[t] = x; if (t == 0) *cyc = t1 - t0;
}
// L2 latency: dependent ld.relaxed.gpu chain over a buffer
__global__ void k_l2(const int *buf, int *out, int n, long long *cyc)
{
    int p = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        int v;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(buf + p) : "memory");
        p = v;
    }
    long long t1 = clock64();
    *out = p; *cyc = t1 - t0;
}
// hand-off between two CTAs on different SMs through L2: ping-pong of a relaxed 8-byte value
__global__ void k_pingpong(volatile unsigned long long *a, volatile unsigned long long *b, int n, long long *cyc, int fence)
{
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 1; i <= n; i++) {
            if (fence) __threadfence();
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(a), "l"((unsigned long long)i) : "memory");
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(b) : "memory"); } while (v != (unsigned long long)i);
        }
        long long t1 = clock64();
        *cyc = t1 - t0;
    }
    else if (blockIdx.x == gridDim.x - 1) {
        for (int i = 1; i <= n; i++) {
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory"); } while (v != (unsigned long long)i);
            if (fence) __threadfence();
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(b), "l"((unsigned long long)i) : "memory");
        }
    }
}
__global__ void k_fence(double *g, int n, long long *cyc)
{
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        g[threadIdx.x + 1024 * (i & 7)] = i;
        __threadfence();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double *d_out; long long *d_cyc, h;
    int *d_i;
    CK(cudaMalloc(&d_out, 1 << 20)); CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_i, 1 << 20));
    const int n = 1000;
#define RUN(name, per, ...) do { __VA_ARGS__; CK(cudaDeviceSynchronize()); __VA_ARGS__; CK(cudaDeviceSynchronize()); \
        CK(cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost)); printf("%-44s %8.1f cycles\n", name, (double)h / (per)); } while (0)
    RUN("dependent DADD (1 warp)", n * 16.0, k_dadd<<<1, 32>>>(d_out, 1.0, 1e-9, n, d_cyc));
    RUN("dependent DADD (8 warps/CTA)", n * 16.0, k_dadd<<<1, 256>>>(d_out, 1.0, 1e-9, n, d_cyc));
    RUN("dependent DADD (16 warps on the SM)", n * 16.0, k_dadd<<<1, 512>>>(d_out, 1.0, 1e-9, n, d_cyc));
    RUN("dependent DMUL (1 warp)", n * 16.0, k_dmul<<<1, 32>>>(d_out, 1.0, 1.0000001, n, d_cyc));
    RUN("dependent DMUL+DADD (1 warp)", n * 16.0, k_dmuladd<<<1, 32>>>(d_out, 1.0, 0.5, 1.0, n, d_cyc));
    RUN("dependent DMUL+DADD (8 warps)", n * 16.0, k_dmuladd<<<1, 256>>>(d_out, 1.0, 0.5, 1.0, n, d_cyc));
    RUN("dependent DFMA (1 warp)", n * 16.0, k_dfma<<<1, 32>>>(d_out, 1.0, 0.5, 1.0, n, d_cyc));
    RUN("dependent DDIV (1 warp)", n * 16.0, k_ddiv<<<1, 32>>>(d_out, 1.0, 1.0000001, n, d_cyc));
    RUN("dependent DDIV (8 warps)", n * 16.0, k_ddiv<<<1, 256>>>(d_out, 1.0, 1.0000001, n, d_cyc));
    RUN("dependent Markstein div (1 warp)", n * 16.0, k_ddiv_mark<<<1, 32>>>(d_out, 1.0, 1.0000001, n, d_cyc));
    RUN("dependent Markstein div (8 warps)", n * 16.0, k_ddiv_mark<<<1, 256>>>(d_out, 1.0, 1.0000001, n, d_cyc));
    RUN("dependent LDS (1 warp)", n * 16.0, k_lds<<<1, 32>>>(d_i, n, d_cyc));
    RUN("STS+bar+LDS+1 sub step, 64 thr", (double)n, k_step<<<1, 64>>>(d_out, n, d_cyc));
    RUN("STS+bar+LDS+1 sub step, 128 thr", (double)n, k_step<<<1, 128>>>(d_out, n, d_cyc));
    RUN("STS+bar+LDS+1 sub step, 256 thr", (double)n, k_step<<<1, 256>>>(d_out, n, d_cyc));
    RUN("STS+bar+LDS+1 sub step, 512 thr", (double)n, k_step<<<1, 512>>>(d_out, n, d_cyc));
    RUN("ILU(0) row step (3 sub), 256 thr", (double)n, k_step3<0><<<1, 256>>>(d_out, n, 1.5, d_cyc));
    RUN("ILU(0) row step (3 sub + div), 256 thr", (double)n, k_step3<1><<<1, 256>>>(d_out, n, 1.5, d_cyc));
    RUN("ILU(0) row step (3 sub), 128 thr", (double)n, k_step3<0><<<1, 128>>>(d_out, n, 1.5, d_cyc));
    RUN("ILU(0) row step (3 sub + div), 128 thr", (double)n, k_step3<1><<<1, 128>>>(d_out, n, 1.5, d_cyc));
    RUN("ILU(0) row step (3 sub), 2 CTAs x 256 on 1 SM?", (double)n, k_step3<0><<<296, 256>>>(d_out, n, 1.5, d_cyc));
    RUN("ILU(0) row step (3 sub + div), 296 CTAs x 256", (double)n, k_step3<1><<<296, 256>>>(d_out, n, 1.5, d_cyc));
    {   // L2 latency: stride chain of 64 MB
        const int N = 16 << 20;
        int *hb = (int *)malloc(N * 4), *db;
        for (int i = 0; i < N; i++) hb[i] = (int)(((long long)i + 1048583) % N);
        CK(cudaMalloc(&db, N * 4)); CK(cudaMemcpy(db, hb, N * 4, cudaMemcpyHostToDevice));
        RUN("dependent ld.relaxed.gpu, 64 MB chain (L2/HBM)", 2000.0, k_l2<<<1, 1>>>(db, d_i, 2000, d_cyc));
        for (int i = 0; i < N; i++) hb[i] = (i + 64) % 4096;
        CK(cudaMemcpy(db, hb, N * 4, cudaMemcpyHostToDevice));
        RUN("dependent ld.relaxed.gpu, 16 KB chain (L2 hit)", 2000.0, k_l2<<<1, 1>>>(db, d_i, 2000, d_cyc));
    }
    {
        unsigned long long *f;
        CK(cudaMalloc(&f, 1024)); 
        for (int g : {2, 74, 148}) {
            char nm[96];
            CK(cudaMemset(f, 0, 1024));
            snprintf(nm, 96, "ping-pong round trip, CTA 0 <-> CTA %d", g - 1);
            RUN(nm, 1000.0, (cudaMemset(f, 0, 1024), k_pingpong<<<g, 32>>>(f, f + 32, 1000, d_cyc, 0)));
            snprintf(nm, 96, "ping-pong with __threadfence, CTA 0 <-> %d", g - 1);
            RUN(nm, 1000.0, (cudaMemset(f, 0, 1024), k_pingpong<<<g, 32>>>(f, f + 32, 1000, d_cyc, 1)));
        }
    }
    RUN("store + __threadfence (256 thr)", 1000.0, k_fence<<<1, 256>>>(d_out, 1000, d_cyc));
    return 0;
}
