// tools/ubench.cu — issue rate of the instruction classes the solve kernel is made of (sm_100a), per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/_build/ubench tools/ubench.cu && tools/_build/ubench
// Every loop trip issues 8 independent chains x 8 links of the op under test (64 instructions + ~4 of loop overhead);
// one CTA on one SM with 4 / 8 / 12 warps = 1 / 2 / 3 warps per sub-partition.  Printed: cycles per instruction per
// SUB-PARTITION (all its warps together) — the reciprocal is the instructions per cycle that class can reach.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 512
#define CH 8
#define LINKS 8
template <int OP>
__global__ void k(long long* out, float seed) {
    float2 f[CH];
    int iv[CH];
    double d[CH];
    for (int c = 0; c < CH; c++) { f[c] = make_float2(seed + c, seed * c); iv[c] = (int)(seed * 1000) + c * 77; d[c] = seed + c; }
    const float2 m = make_float2(seed * 0.999f, seed * 1.001f);
    const double dm = seed * 0.999, dk = seed * 1.5;
    const int im = (int)(seed * 123457);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int l = 0; l < LINKS; l++)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                if (OP == 0) f[c] = __ffma2_rn(f[c], m, m);
                if (OP == 1) iv[c] = min(iv[c], im + c + l);                       // VIMNMX
                if (OP == 2) d[c] = fma(d[c], dm, dm);                             // DFMA
                if (OP == 3) f[c].x = fmaf(f[c].x, m.x, m.y);                      // FFMA
                if (OP == 4) d[c] = d[c] + dm;                                     // DADD
                if (OP == 5) iv[c] = (iv[c] & im) | (c + l);                       // LOP3
                if (OP == 6) { d[c] = fma(d[c], dm, dm); f[c].x = fmaf(f[c].x, m.x, m.y); }   // DFMA + FFMA
                if (OP == 7) { d[c] = fma(d[c], dm, dm); iv[c] = min(iv[c], im + c + l); }    // DFMA + VIMNMX
                if (OP == 8) { d[c] = fma(d[c], dm, dm); iv[c] = (iv[c] & im) | (c + l); }    // DFMA + LOP3
                if (OP == 9) { d[c] = (d[c] > dk) ? dm : d[c]; }                              // DSETP + 2 FSEL
                if (OP == 10) { d[c] = fma(d[c], dm, dm); f[c] = __ffma2_rn(f[c], m, m); }    // DFMA + FFMA2
                if (OP == 11) { f[c].x = fmaf(f[c].x, m.x, m.y); iv[c] = min(iv[c], im + c + l); }  // FFMA + VIMNMX
            }
    }
    long long t1 = clock64();
    float acc = 0; for (int c = 0; c < CH; c++) acc += f[c].x + f[c].y + iv[c] + (float)d[c];
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)acc; }
}
template <int OP> void run(const char* name, int ops_per, long long* dout) {
    printf("%-16s", name);
    for (int nw : {4, 8, 12}) {
        k<OP><<<1, 32 * nw>>>(dout, 1.0001f);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
        printf("  %d warp/SMSP: %.2f", nw / 4, (double)h[0] / ((double)ITER * CH * LINKS * ops_per * (nw / 4)));
    }
    printf("   (cycles per instruction per sub-partition)\n");
}
int main() {
    long long* dout; cudaMalloc(&dout, 16);
    run<0>("FFMA2", 1, dout); run<1>("VIMNMX", 1, dout); run<2>("DFMA", 1, dout); run<3>("FFMA", 1, dout); run<4>("DADD", 1, dout);
    run<5>("LOP3", 1, dout); run<6>("DFMA+FFMA", 2, dout); run<7>("DFMA+VIMNMX", 2, dout); run<8>("DFMA+LOP3", 2, dout);
    run<9>("DSETP+2FSEL", 3, dout); run<10>("DFMA+FFMA2", 2, dout); run<11>("FFMA+VIMNMX", 2, dout);
    return 0;
}
