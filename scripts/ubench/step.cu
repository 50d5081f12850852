// Micro-benchmark of the pencil kernel's compute step: which ingredient makes a step cost far more than the
// 117 cycles of the plain  LDS x3 -> 3 subtractions -> STS -> __syncthreads  loop?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o step step.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ int ld_flag(const int *p)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned int)__cvta_generic_to_shared(p)) : "memory");
    return v;
}

// FEAT bits: 1 named barrier among the first T threads of a bigger CTA (extra warps sleep-spin on a flag)
//            2 ring with wrap logic (16 rows) instead of 2 rows
//            4 values from a shared-memory ring (3 LDS) instead of constants
//            8 rhs from a tile + second STS (out tile)
//           16 flag poll every 2 steps
//           32 st.relaxed.gpu by lanes 15 and 31 (mailbox)
//           64 extra warps spin WITHOUT sleeping
template <int FEAT>
__global__ void k_step(double *out, double *mail, int n, int T, long long *cyc)
{
    extern __shared__ double sm[];
    __shared__ int flag[4];
    const int t = threadIdx.x;
    const int RS = T + 48, RD = 16;
    double *ring = sm, *vr = ring + RD * RS, *rt = vr + 8 * 3 * T, *ot = rt + T * 9;
    for (int i = t; i < RD * RS + 8 * 3 * T + 2 * T * 9; i += blockDim.x) sm[i] = 0.25;
    if (t == 0) { flag[0] = 1 << 30; flag[1] = 0; }
    __syncthreads();
    if (t >= T) {
        if (FEAT & 1) {
            // extra warps: wait for the compute warps
            while (ld_flag(&flag[1]) < n) { if (!(FEAT & 64)) __nanosleep(100); }
        }
        return;
    }
    const unsigned int RS8 = RS * 8, ring_bytes = RD * RS8;
    char *ringb = (char *)ring;
    unsigned int rop[3] = {15 * RS8 + ((t + T - 16) % T) * 8u, 15 * RS8 + ((t + T - 1) % T) * 8u, 15 * RS8 + t * 8u};
    unsigned int rown = t * 8u, voff = 0;
    const unsigned int vstep = 3 * T * 8, vbytes = 8 * vstep;
    const char *vb = (const char *)(vr + t);
    double x = 1.0;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if ((FEAT & 16) && (i & 1) == 0) while (ld_flag(&flag[0]) < i) {}
        double acc = (FEAT & 8) ? rt[t * 9 + (i & 7)] : 1.0;
        double av[3] = {0.25, 0.25, 0.25};
        if (FEAT & 4) {
#pragma unroll
            for (int w = 0; w < 3; w++) av[w] = *(const double *)(vb + voff + w * T * 8);
            voff += vstep;
            if (voff >= vbytes) voff = 0;
        }
        if (FEAT & 2) {
#pragma unroll
            for (int w = 0; w < 3; w++) {
                const double xv = *(const double *)(ringb + rop[w]);
                rop[w] += RS8;
                if (rop[w] >= ring_bytes) rop[w] -= ring_bytes;
                acc = acc - av[w] * xv;
            }
            *(double *)(ringb + rown) = acc;
            rown += RS8;
            if (rown >= ring_bytes) rown -= ring_bytes;
        }
        else {
            const double a = ring[(i & 1) * RS + (t + T - 16) % T], b = ring[(i & 1) * RS + (t + T - 1) % T], c = ring[(i & 1) * RS + t];
            acc = acc - av[0] * a;
            acc = acc - av[1] * b;
            acc = acc - av[2] * c;
            ring[((i + 1) & 1) * RS + t] = acc;
        }
        if (FEAT & 8) ot[t * 9 + (i & 7)] = acc;
        if ((FEAT & 32) && (t & 15) == 15) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(mail + (size_t)t * 4096 + i), "d"(acc) : "memory");
        x = acc;
        if (FEAT & 1) asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory");
        else __syncthreads();
        if ((FEAT & 1) && (i & 1) && t == 0) asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(&flag[1])), "r"(i + 1) : "memory");
    }
    long long t1 = clock64();
    out[t] = x;
    if (t == 0) { *cyc = t1 - t0; asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(&flag[1])), "r"(n + 1) : "memory"); }
}

template <int FEAT>
void run(const char *name, int T, int block, double *d_out, double *d_mail, long long *d_cyc, int grid = 1)
{
    const int n = 1024;
    const size_t smem = sizeof(double) * (16 * (T + 48) + 8 * 3 * T + 2 * T * 9);
    CK(cudaFuncSetAttribute(k_step<FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long h;
    for (int rep = 0; rep < 2; rep++) {
        k_step<FEAT><<<grid, block, smem>>>(d_out, d_mail, n, T, d_cyc);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("%-70s %7.1f cycles/step\n", name, (double)h / n);
}

int main()
{
    double *d_out, *d_mail; long long *d_cyc;
    CK(cudaMalloc(&d_out, 1 << 20)); CK(cudaMalloc(&d_mail, 256 * 4096 * 8 + (1 << 20))); CK(cudaMalloc(&d_cyc, 8));
    run<0>("T=256: plain (2-row ring, __syncthreads)", 256, 256, d_out, d_mail, d_cyc);
    run<1>("T=256 of 480: named barrier, extra warps sleep-spin", 256, 480, d_out, d_mail, d_cyc);
    run<1 | 64>("T=256 of 480: named barrier, extra warps spin hard", 256, 480, d_out, d_mail, d_cyc);
    run<2>("T=256: 16-row ring with wrap logic", 256, 256, d_out, d_mail, d_cyc);
    run<2 | 4>("T=256: + values from smem ring", 256, 256, d_out, d_mail, d_cyc);
    run<2 | 4 | 8>("T=256: + rhs tile, out tile", 256, 256, d_out, d_mail, d_cyc);
    run<2 | 4 | 8 | 16>("T=256: + flag poll every 2 steps", 256, 256, d_out, d_mail, d_cyc);
    run<2 | 4 | 8 | 16 | 32>("T=256: + mailbox stores (2 lanes per warp)", 256, 256, d_out, d_mail, d_cyc);
    run<1 | 2 | 4 | 8 | 16 | 32>("T=256 of 480: everything, extra warps sleep-spin", 256, 480, d_out, d_mail, d_cyc);
    run<1 | 2 | 4 | 8 | 16 | 32>("T=128 of 352: everything", 128, 352, d_out, d_mail, d_cyc);
    run<1 | 2 | 4 | 8 | 16 | 32>("T=256 of 480: everything, 148 CTAs", 256, 480, d_out, d_mail, d_cyc, 148);
    return 0;
}
