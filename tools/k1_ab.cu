// A/B timing harness for the fused step kernel: compile against different revisions of sb_kernels.cuh
// (-I<dir> -DKV=<1|2|3>) and compare on the same GPU.  Usage: k1_ab <nx> <ny> <nz> <steps> [rj] [chunk]
#include "sb_kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace sb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
int main(int argc, char **argv)
{
    int nx = argc > 1 ? atoi(argv[1]) : 512, ny = argc > 2 ? atoi(argv[2]) : 512, nz = argc > 3 ? atoi(argv[3]) : 512;
    int steps = argc > 4 ? atoi(argv[4]) : 20, rj = argc > 5 ? atoi(argv[5]) : 2, chunk = argc > 6 ? atoi(argv[6]) : 64;
    int var = argc > 7 ? atoi(argv[7]) : 0;      // KV>=4: 0 UNI lean, 1 tables (non-UNI), 2 UNI+PEER, 3 UNI+FUSE, 4 tables + solids (config 4), 5 UNI + solids, 6 / 7 = 5 with the ADE box left to the list kernels (BOXM 3) / skipped (BOXM 1), 8 = config 3's sphere: no solids, BOXM 3, 9 = config 3's slab: no solids, BOXM 1 (the last quarter of the planes skipped)
    int wj = argc > 8 ? atoi(argv[8]) : 8, wk = argc > 9 ? atoi(argv[9]) : 1;
    int pitch = (nz + 7) / 8 * 8;
    long long plane = (long long)ny * pitch, elems = (long long)(nx + 2) * plane;
    float *buf[8];
    for (int q = 0; q < 8; q++) { CK(cudaMalloc(&buf[q], elems * 4)); CK(cudaMemset(buf[q], 0, elems * 4)); }
    auto table = [&](int n, float v) { std::vector<float> h(n + 8, v); float *d; CK(cudaMalloc(&d, (n + 8) * 4));
                                       CK(cudaMemcpy(d, h.data(), (n + 8) * 4, cudaMemcpyHostToDevice)); return d; };
    StepParams P;
    memset(&P, 0, sizeof P);
    P.cvx = table(nx + 2, -1e-3f) + 1; P.cvy = table(ny, -1e-3f); P.cvz = table(pitch, -1e-3f);
    P.n_sponge = 1; P.decx[0] = table(nx + 2, 0.999f) + 1; P.decy[0] = table(ny, 0.999f); P.decz[0] = table(pitch, 0.999f);
    P.cp = -100.f; P.nx = nx; P.ny = ny; P.nz = nz; P.pitch = pitch; P.plane = plane;
    P.i_begin = 0; P.i_end = nx; P.chunk_i = chunk;
#if KV >= 4
    P.cv_uni = -1e-3f;
    if (var == 1 || var == 4) { P.icx = table(nx + 2, 1.0f) + 1; P.icy = table(ny, 1.0f); P.icz = table(pitch, 1.0f); }
    if (var >= 4 && var <= 7) {                  // all open except a solid box in the middle (faces closed around it)
        std::vector<uint8_t> m((size_t)elems, 0x0F);
        for (int i = nx / 3; i < nx / 2; i++) for (int j = ny / 3; j < ny / 2; j++) for (int k = nz / 3; k < nz / 2; k++)
            m[(size_t)(i + 1) * plane + (size_t)j * pitch + k] = 0;
        uint8_t *d; CK(cudaMalloc(&d, elems)); CK(cudaMemcpy(d, m.data(), elems, cudaMemcpyHostToDevice));
        P.mask = d + plane;
    }
    if (var == 8) { uint8_t *d; CK(cudaMalloc(&d, elems)); CK(cudaMemset(d, 0x0F, elems)); P.ade_mask = d + plane; }
    if (var >= 6 && var <= 9) {                  // config 3's sphere: the middle fifth of every axis
        P.box_mode = (var == 7 || var == 9) ? 1 : 3;
        P.bi0 = (2 * nx / 5) / chunk * chunk; P.bi1 = (3 * nx / 5 / chunk + 1) * chunk; P.bj0 = 2 * ny / 5; P.bj1 = 3 * ny / 5;
        P.bk0 = (2 * nz / 5) / 4 * 4; P.bk1 = (3 * nz / 5) / 4 * 4;
        if (var < 8) P.ade_mask = P.mask;
        if (var == 9) { P.bi0 = 3 * nx / 4 / chunk * chunk; P.bi1 = nx; P.bj0 = 0; P.bj1 = ny; P.bk0 = 0; P.bk1 = (nz + 3) / 4 * 4; }   // config 3's slab
    }
#endif
#if KV >= 2
    int *ctr; CK(cudaMalloc(&ctr, 8)); CK(cudaMemset(ctr, 0, 8));
    P.step_global = ctr; P.err_flag = ctr + 1;
#endif
    dim3 blk(32 * wk, wj), grd((nz + 128 * wk - 1) / (128 * wk), (ny + rj * wj - 1) / (rj * wj), (nx + chunk - 1) / chunk);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, sum = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        for (int s = 0; s < steps; s++) {
            int in = s & 1;
            P.p_in = buf[in * 4] + plane; P.vx_in = buf[in * 4 + 1] + plane; P.vy_in = buf[in * 4 + 2] + plane; P.vz_in = buf[in * 4 + 3] + plane;
            int o = 1 - in;
            P.p_out = buf[o * 4] + plane; P.vx_out = buf[o * 4 + 1] + plane; P.vy_out = buf[o * 4 + 2] + plane; P.vz_out = buf[o * 4 + 3] + plane;
#if KV >= 4
            if (rj == 1) {
                if (var == 0) k1_step_march<1, false, true, false, false><<<grd, blk>>>(P);
                else if (var == 1) k1_step_march<1, false, false, false, false><<<grd, blk>>>(P);
                else if (var == 2) k1_step_march<1, false, true, true, false><<<grd, blk>>>(P);
                else if (var == 4) k1_step_march<1, true, false, false, false><<<grd, blk>>>(P);
                else if (var == 5) k1_step_march<1, true, true, false, false><<<grd, blk>>>(P);
                else if (var == 6) k1_step_march<1, true, true, false, false, false, 3><<<grd, blk>>>(P);
                else if (var == 7) k1_step_march<1, true, true, false, false, false, 1><<<grd, blk>>>(P);
                else if (var == 8) k1_step_march<1, false, true, false, false, false, 3><<<grd, blk>>>(P);
                else if (var == 9) k1_step_march<1, false, true, false, false, false, 1><<<grd, blk>>>(P);
                else k1_step_march<1, false, true, false, true><<<grd, blk>>>(P);
            } else {
                if (var == 0) k1_step_march<2, false, true, false, false><<<grd, blk>>>(P);
                else if (var == 1) k1_step_march<2, false, false, false, false><<<grd, blk>>>(P);
                else if (var == 2) k1_step_march<2, false, true, true, false><<<grd, blk>>>(P);
                else if (var == 4) k1_step_march<2, true, false, false, false><<<grd, blk>>>(P);
                else if (var == 5) k1_step_march<2, true, true, false, false><<<grd, blk>>>(P);
                else if (var == 6) k1_step_march<2, true, true, false, false, false, 3><<<grd, blk>>>(P);
                else if (var == 7) k1_step_march<2, true, true, false, false, false, 1><<<grd, blk>>>(P);
                else if (var == 8) k1_step_march<2, false, true, false, false, false, 3><<<grd, blk>>>(P);
                else if (var == 9) k1_step_march<2, false, true, false, false, false, 1><<<grd, blk>>>(P);
                else k1_step_march<2, false, true, false, true><<<grd, blk>>>(P);
            }
#else
            if (rj == 1) k1_step_march<1, false><<<grd, blk>>>(P);
            else if (rj == 2) k1_step_march<2, false><<<grd, blk>>>(P);
            else k1_step_march<4, false><<<grd, blk>>>(P);
#endif
        }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= steps;
        if (rep) { best = ms < best ? ms : best; sum += ms; }
    }
    double cells = (double)nx * ny * nz;
    printf("KV=%d var=%d %dx%dx%d rj=%d chunk=%d  best %.4f ms  mean %.4f ms  %.1f Gcell/s  %.3f of 6540.8 GB/s\n", KV, var, nx, ny, nz, rj, chunk,
           best, sum / 4, cells / best / 1e6, cells * 32 / (best * 1e-3) / 6540.8e9);
    return 0;
}
