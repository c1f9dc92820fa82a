// sb_kernels.cuh -- sm_100a kernels for the strata-fdtd time-stepping hot path.
//
// Arithmetic contract: every fp32 operation is separately rounded, in the order the
// reference's C++ backend performs it (this file is compiled with --fmad=false and
// without --use_fast_math), so results are bit-identical to
// /root/reference/src/strata_fdtd/_kernels/src/{fdtd_step,boundaries,pml,ade,microphones}.cpp.
//
// Layout: fields are [i][j][k] with k contiguous, row pitch `pitch` floats (multiple of 4),
// plane stride `plane` = ny*pitch, and the pointers handed to the kernels address local
// plane 0, so plane -1 (lower ghost) and plane nx (upper ghost) are valid addresses.
//
// One time step reads the "in" set {p,vx,vy,vz} and writes the "out" set (ping-pong), which
// makes every cell independent: 4 reads + 4 writes of fp32 = 32 B per cell update.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// helpers shared with the host-side lockstep emulation of the resident kernel (tests/emu): plain arithmetic only
#define SB_HD __host__ __device__ __forceinline__

namespace sb {

constexpr int MAX_SPONGES = 4;
constexpr uint8_t M_AIR = 1, M_XOPEN = 2, M_YOPEN = 4, M_ZOPEN = 8;
// upper nibble (fused ADE path): the cell carries a material with poles; its +x / +y / +z neighbour carries the SAME one
constexpr uint8_t M_XSAME = 0x10, M_YSAME = 0x20, M_ZSAME = 0x40, M_ADE = 0x80;

// point-source entries (cell, source, weight) a step / chunk kernel applies itself instead of leaving them to K3: a
// phased array of a few dozen elements still fits (one bit each in a thread's 32-bit ownership mask)
constexpr int SB_MAX_INLINE = 32;

struct StepParams {
    const float *p_in, *vx_in, *vy_in, *vz_in;
    float *p_out, *vx_out, *vy_out, *vz_out;
    const uint8_t *mask;                 // nullptr = all air, nothing rigid
    const float *cvx, *cvy, *cvz;        // per-face velocity coefficients (cvx valid from index -1)
    const float *icx, *icy, *icz;        // per-cell inverse spacing, or nullptr (uniform grid)
    const float *decx[MAX_SPONGES], *decy[MAX_SPONGES], *decz[MAX_SPONGES];   // decx valid from -1
    int n_sponge;
    float cp;
    int nx, ny, nz, pitch;
    long long plane;
    int has_lower, has_upper;
    int i_begin, i_end, chunk_i;
    // peer-to-peer halo (multi-GPU, NVLink): K1 stores its first / last p plane straight into the neighbour's
    // ghost plane and the blocks that touch a cut wait for the neighbour's "step done" flag first.
    float *peer_lo_p, *peer_hi_p;        // neighbour ghost planes of the set being written, or nullptr
    const int *flag_lo, *flag_hi;        // my flags: number of steps the lower / upper neighbour has completed
    const int *step_global;              // index of the step being executed (device counter)
    int *err_flag;
    int two_range;                       // PEER launch over the two cut ranges [0,chunk_i) and [nx-chunk_i,nx)
    float cv_uni;                        // UNI kernels: the one velocity coefficient of a uniform grid
    // single-kernel step ("fused K3"): a handful of point sources are injected by the thread that owns the cell
    // right before it stores p, and the probes / microphones of the PREVIOUS step are recorded from the input set
    // by the first warp of block 0 (they are final there); a tail launch records the last step of a chunk.
    int n_inline;                        // 0 = off
    int inl_i[SB_MAX_INLINE], inl_j[SB_MAX_INLINE], inl_k[SB_MAX_INLINE], inl_src[SB_MAX_INLINE];
    double inl_weight[SB_MAX_INLINE];
    const double *src_row;               // this step's waveform samples [n_sources]
    int rec_prev;                        // 1 = record the previous step's slots in this launch
    int n_probes, n_mics;
    const long long *probe_off, *mic_off8;
    const int *mic_field;
    const float *mic_w8;
    float *rec_row;                      // record slots of the previous step [n_probes + n_mics]
    // A box of cells [bi0,bi1) x [bj0,bj1) x [bk0,bk1) that ANOTHER launch of the same step computes (the ADE variant
    // of the kernel, which covers the bounding box of the dispersive materials): threads of this launch still load and
    // compute there -- their neighbours' shuffles need the values -- but store nothing, and warps that lie entirely
    // inside leave at once.  box_mode (K1: the template parameter BOXM) 0 = no box, 1 = this launch skips the box,
    // 2 = this launch owns only the box (K1-ADE), 3 = inside the box this launch leaves out, cell by cell, what the concurrently running ADE list
    // kernels (K2a / K2b) write: p of every material cell and the + faces they correct (mask bits M_ADE, M_?SAME).
    int box_mode, bi0, bi1, bj0, bj1, bk0, bk1;
    const uint8_t *ade_mask;             // mask bytes carrying the ADE bits (box_mode 3; == mask when GEOM)
    int bx_off, by_off, bz_off;          // block index offsets (a launch restricted to the tiles that meet the box)
};

// 8-point weighted gather of one field (trilinear microphone sample): sum = 0; sum += w[c]*f[idx[c]], fp32,
// corner order of microphones.hpp:30-32 (microphones.cpp:82-116; Python path core/solver.py:1027-1034, 1085-1099)
struct FieldPtrs { const float *f[4]; };
__device__ __forceinline__ float gather8(const FieldPtrs &F, const int *mic_field, const long long *mic_off8,
                                         const float *mic_w8, int m)
{
    const float *f = F.f[mic_field ? mic_field[m] : 0];
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; c++) sum = sum + mic_w8[8 * m + c] * f[mic_off8[8 * m + c]];
    return sum;
}

__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// one lane per warp polls; bounded so that a dead neighbour cannot hang the GPU
__device__ __forceinline__ void wait_neighbour(const int *flag, int target, int *err)
{
    if ((threadIdx.x & 31) == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) < target) {
            __nanosleep(100);
            if (clock64() - t0 > (6LL << 30)) { atomicExch(err, 1); break; }      // ~3 s
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// K0: one thread per cell, every face velocity recomputed from the "in" set.  Reference
// implementation of the fused step on the device; also the cross-check for the marching kernel.
//   velocity  fdtd_step.cpp:34-80 / 235-307      rigid     boundaries.cpp:66-89
//   pressure  fdtd_step.cpp:109-211 / 309-439    sponge    pml.cpp:47-149
// ------------------------------------------------------------------------------------------
template <bool GEOM>
__device__ __forceinline__ float face_v(const StepParams &P, const float *v_in, const float *cv_tab,
                                        long long c, long long c_next, int m_idx, bool update,
                                        uint8_t open_bit)
{
    float v = v_in[c];
    if (update) {
        v = v + cv_tab[m_idx] * (P.p_in[c_next] - P.p_in[c]);
        if (GEOM && !(P.mask[c] & open_bit)) v = 0.0f;
    }
    return v;
}

template <bool GEOM>
__global__ void __launch_bounds__(256) k0_step_naive(StepParams P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int i = P.i_begin + (int)blockIdx.z;       // i_begin may be -1 (ghost vx maintenance)
    if (k >= P.nz || i >= P.i_end) return;
    const long long c = (long long)i * P.plane + (long long)j * P.pitch + k;
    const bool upd_x = (i < P.nx - 1) || P.has_upper;
    const float vxn = face_v<GEOM>(P, P.vx_in, P.cvx, c, c + P.plane, i, upd_x, M_XOPEN);
    if (i < 0) {                                      // ghost plane: only vx is maintained here
        float v = vxn;
        for (int s = 0; s < P.n_sponge; s++) v = v * P.decx[s][i];
        P.vx_out[c] = v;
        return;
    }
    const float vyn = face_v<GEOM>(P, P.vy_in, P.cvy, c, c + P.pitch, j, j < P.ny - 1, M_YOPEN);
    const float vzn = face_v<GEOM>(P, P.vz_in, P.cvz, c, c + 1, k, k < P.nz - 1, M_ZOPEN);
    float ddx = vxn, ddy = vyn, ddz = vzn;
    if (i > 0 || P.has_lower)
        ddx = vxn - face_v<GEOM>(P, P.vx_in, P.cvx, c - P.plane, c, i - 1, true, M_XOPEN);
    if (j > 0)
        ddy = vyn - face_v<GEOM>(P, P.vy_in, P.cvy, c - P.pitch, c, j - 1, true, M_YOPEN);
    if (k > 0)
        ddz = vzn - face_v<GEOM>(P, P.vz_in, P.cvz, c - 1, c, k - 1, true, M_ZOPEN);
    if (P.icx) { ddx = P.icx[i] * ddx; ddy = P.icy[j] * ddy; ddz = P.icz[k] * ddz; }
    float pn = P.p_in[c] + P.cp * ((ddx + ddy) + ddz);
    if (GEOM && !(P.mask[c] & M_AIR)) pn = 0.0f;
    float ox = vxn, oy = vyn, oz = vzn;
    for (int s = 0; s < P.n_sponge; s++) {
        const float dx = P.decx[s][i], dy = P.decy[s][j], dz = P.decz[s][k];
        ox = ox * dx; oy = oy * dy; oz = oz * dz;
        pn = ((pn * dx) * dy) * dz;
    }
    P.p_out[c] = pn; P.vx_out[c] = ox; P.vy_out[c] = oy; P.vz_out[c] = oz;
}

// ------------------------------------------------------------------------------------------
// K1: 2.5-D marching kernel.  A warp owns a strip of 128 cells along k (float4 per lane) by RJ
// rows along j and marches along i over one chunk, keeping the p planes i / i+1 and the
// undamped vx of plane i-1 in registers.  Neighbours along k come from warp shuffles, the two
// strip-edge values from scalar loads that hit lines the neighbouring strip streams anyway.
// No shared memory, no block barrier: warps are independent; the block only groups strips
// that are adjacent in j so that their halo rows hit in L1.
// ------------------------------------------------------------------------------------------
SB_HD float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
SB_HD void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
SB_HD float4 f4(float a) { return make_float4(a, a, a, a); }
SB_HD float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
SB_HD float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
SB_HD float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
SB_HD float4 mul4s(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
SB_HD float4 sel4(bool cx, bool cy, bool cz, bool cw, float4 a, float4 b)
{ return make_float4(cx ? a.x : b.x, cy ? a.y : b.y, cz ? a.z : b.z, cw ? a.w : b.w); }

// keep element e of v iff byte e of the packed mask word has `bit` set, else +0.0f (boundaries.cpp:66-89)
SB_HD float4 keep4(float4 v, unsigned w, unsigned bit)
{
    return make_float4((w & bit) ? v.x : 0.0f, (w & (bit << 8)) ? v.y : 0.0f,
                       (w & (bit << 16)) ? v.z : 0.0f, (w & (bit << 24)) ? v.w : 0.0f);
}
constexpr unsigned ALL_OPEN = 0x0F0F0F0Fu;      // four cells: air, all three faces open

// Compile-time variants keep the common path lean (the kernel is register-bound at 128 regs / 2 blocks per SM):
//   UNI  uniform grid: one scalar velocity coefficient, no inverse-cell multiplies
//   PEER multi-GPU: neighbour flags + peer stores (launched only over the chunks that touch a cut)
//   FUSE single-kernel step: inline point-source injection + deferred probe recording (small grids)
// The tile body is a device function so that the step-pipelined kernel (sb_pipeline.cuh) can run the same code
// for a tile of any step; F names the input and output sets of that step.
struct FieldSet { const float *p_in, *vx_in, *vy_in, *vz_in; float *p_out, *vx_out, *vy_out, *vz_out; };

// FLAT = how the (j, k) plane is dealt to the warps.  Strip mode (false): a warp owns 128 consecutive cells of one
// row group; rows whose length is not a multiple of 128 leave lanes idle (nz = 200: 28 of 64), and K1 is latency-bound
// per warp iteration, so an under-filled warp costs as much as a full one.  Flat mode (true): the float4 groups of
// the whole plane -- (row group, k) in memory order -- are dealt to the lanes consecutively, so a warp may hold the end
// of one row group and the start of the next; the only idle lanes are the row padding.  A lane's k-neighbours are
// still its neighbouring lanes wherever a neighbour exists (nothing crosses a row end), so the shuffles stay valid.
// Strip mode keeps whole blocks adjacent in j (halo rows hit in L1) and is used for rows that fill their strips.
// One more such steering (see k1_min_blocks below): the variant with solids AND spacing tables, one row per thread --
// config 4 -- keeps the run-time form of the box test, which is never true there (box_mode is 0 whenever BOXM is 0).  With
// it ptxas emits the schedule that runs at 199.1 Gcell-updates/s instead of 187.9 (tools/k1_ab.cu, variant 4).
#if defined(SB_K1_RTBOX_ALL)
#define SB_K1_RTBOX(RJ, GEOM, UNI) true
#elif defined(SB_K1_RTBOX_NONE)
#define SB_K1_RTBOX(RJ, GEOM, UNI) false
#else
#define SB_K1_RTBOX(RJ, GEOM, UNI) ((RJ) == 1 && (GEOM) && !(UNI))
#endif
template <int RJ, bool GEOM, bool UNI, bool PEER, bool FUSE, bool FLAT = false, int BOXM = 0>
__device__ __forceinline__ void k1_tile(const StepParams &P, const FieldSet &F, int bx, int by, int bz)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int k0, j0;
    bool warp_out;                                             // the whole warp lies outside the plane
    if (FLAT) {
        const int P4 = P.pitch >> 2, warps_x = blockDim.x >> 5;
        const long long g0 = ((long long)bx * (warps_x * blockDim.y) + threadIdx.y * warps_x + (threadIdx.x >> 5)) * 32;
        const long long groups = (long long)((P.ny + RJ - 1) / RJ) * P4;
        warp_out = g0 >= groups;
        const long long g = g0 + lane;
        const int q = (int)(g / P4);
        k0 = 4 * (int)(g - (long long)q * P4);
        j0 = q * RJ;                                           // lanes past the last row group have j0 >= ny: predicated off
    } else {
        const int warp_k = threadIdx.x >> 5;                   // blockDim.x = 32 * WK
        const int strip_k0 = (bx * (blockDim.x >> 5) + warp_k) * 128;
        k0 = strip_k0 + lane * 4;
        j0 = (by * blockDim.y + threadIdx.y) * RJ;
        warp_out = strip_k0 >= P.nz || j0 >= P.ny;
    }
    int ib = P.i_begin + bz * P.chunk_i;
    if (PEER && P.two_range) ib = bz ? P.nx - P.chunk_i : 0;
    const int ie = min(ib + P.chunk_i, P.i_end);
    if (warp_out || ib >= ie) return;                           // warp-uniform exit
    if (PEER) {
        if (P.flag_lo && ib == 0) wait_neighbour(P.flag_lo, *P.step_global, P.err_flag);
        if (P.flag_hi && ie == P.nx) wait_neighbour(P.flag_hi, *P.step_global, P.err_flag);
    }
    if (FUSE && P.rec_prev && bx == 0 && by == 0 && bz == 0 && threadIdx.y == 0 && threadIdx.x < 32) {
        // probes / microphones of the previous step: its output set is this launch's (read-only) input set
        const FieldPtrs FP{{F.p_in, F.vx_in, F.vy_in, F.vz_in}};
        for (int t = lane; t < P.n_probes + P.n_mics; t += 32)
            P.rec_row[t] = t < P.n_probes ? F.p_in[P.probe_off[t]]
                                          : gather8(FP, P.mic_field, P.mic_off8, P.mic_w8, t - P.n_probes);
    }
    unsigned inl_mask = 0;                                       // inline point sources inside this thread's column
    for (int q = 0; FUSE && q < P.n_inline; q++)
        if (P.inl_j[q] >= j0 && P.inl_j[q] < j0 + RJ && P.inl_k[q] >= k0 && P.inl_k[q] < k0 + 4 &&
            P.inl_i[q] >= ib && P.inl_i[q] < ie) inl_mask |= 1u << q;
    const bool lane_ok = k0 < P.nz;
    const int nz = P.nz, ny = P.ny;
    // cells another launch of this step owns (box_mode 1): computed here as far as the neighbours need them, never stored
    unsigned skip_rows = 0;
    constexpr bool RTBOX = BOXM == 0 && !PEER && !FUSE && !FLAT && SB_K1_RTBOX(RJ, GEOM, UNI);
    if ((BOXM == 1 || (RTBOX && P.box_mode == 1)) && ib >= P.bi0 && ib < P.bi1) {   // the box is aligned to this launch's chunks of planes
        bool nothing_left = true;
#pragma unroll
        for (int r = 0; r < RJ; r++) {
            const bool in = k0 >= P.bk0 && k0 < P.bk1 && j0 + r >= P.bj0 && j0 + r < P.bj1;
            skip_rows |= in ? (1u << r) : 0u;
            nothing_left = nothing_left && (in || !lane_ok || j0 + r >= ny);
        }
        if (__all_sync(FULL, nothing_left)) return;
    }
    // per-element validity and "z face is updated" flags
    const bool e0 = k0 < nz, e1 = k0 + 1 < nz, e2 = k0 + 2 < nz, e3 = k0 + 3 < nz;
    const bool u0 = k0 < nz - 1, u1 = k0 + 1 < nz - 1, u2 = k0 + 2 < nz - 1, u3 = k0 + 3 < nz - 1;
    // ADEX: rows of this thread inside the box of the dispersive materials (their ADE cells belong to K2a / K2b)
    unsigned adex_rows = 0;
    constexpr bool ADEX = BOXM == 3;
    if (ADEX && ib < P.bi1 && ie > P.bi0 && k0 >= P.bk0 && k0 < P.bk1) {
#pragma unroll
        for (int r = 0; r < RJ; r++) adex_rows |= (j0 + r >= P.bj0 && j0 + r < P.bj1) ? (1u << r) : 0u;
    }
    const unsigned uw = (u0 ? 0x80u : 0u) | (u1 ? 0x8000u : 0u) | (u2 ? 0x800000u : 0u) | (u3 ? 0x80000000u : 0u);
    const bool edge_hi = (lane == 31) && (k0 + 4 < nz);        // needs p[k0+4] from the next warp's cells
    const bool edge_lo = (lane == 0) && (k0 > 0);              // needs the z face k0-1 of the previous warp's cells
    const float4 z4 = f4(0.0f);

    // k tables (hoisted)
    const float4 cvz4 = UNI ? f4(P.cv_uni) : (lane_ok ? ld4(P.cvz + k0) : z4);
    const float4 icz4 = (!UNI && lane_ok) ? ld4(P.icz + k0) : f4(1.0f);
    const float4 dz0 = (P.n_sponge > 0 && lane_ok) ? ld4(P.decz[0] + k0) : f4(1.0f);
    const float cvz_lo = UNI ? P.cv_uni : (edge_lo ? P.cvz[k0 - 1] : 0.0f);
    // j tables (hoisted); row index r = -1 .. RJ-1 stored at [r+1]
    float cvy[RJ + 1], icy[RJ], dy0[RJ];
    bool row_ok[RJ + 2];                                       // rows -1 .. RJ
#pragma unroll
    for (int r = -1; r <= RJ; r++) row_ok[r + 1] = (j0 + r >= 0) && (j0 + r < ny);
#pragma unroll
    for (int r = -1; r < RJ; r++) cvy[r + 1] = UNI ? P.cv_uni : ((row_ok[r + 1] && j0 + r < ny - 1) ? P.cvy[j0 + r] : 0.0f);
#pragma unroll
    for (int r = 0; r < RJ; r++) {
        icy[r] = (!UNI && row_ok[r + 1]) ? P.icy[j0 + r] : 1.0f;
        dy0[r] = (P.n_sponge > 0 && row_ok[r + 1]) ? P.decy[0][j0 + r] : 1.0f;
    }
    const long long col = (long long)j0 * P.pitch + k0;         // offset of (row 0, k0) inside a plane

    // ---- prologue: p rows of plane ib, undamped vx of plane ib-1 --------------------------
    float4 pc[RJ + 2];
    float4 vxp[RJ];
    {
        const long long base = (long long)ib * P.plane + col;
#pragma unroll
        for (int r = 0; r < RJ; r++)
            pc[r + 1] = (row_ok[r + 1] && lane_ok) ? ld4(F.p_in + base + (long long)r * P.pitch) : z4;
        const bool have_prev = (ib > 0) || P.has_lower;
        const float cx = UNI ? P.cv_uni : (have_prev ? P.cvx[ib - 1] : 0.0f);
#pragma unroll
        for (int r = 0; r < RJ; r++) {
            float4 v = z4;
            if (have_prev && row_ok[r + 1] && lane_ok) {
                const long long cm = base - P.plane + (long long)r * P.pitch;
                v = add4(ld4(F.vx_in + cm), mul4s(sub4(pc[r + 1], ld4(F.p_in + cm)), cx));
                if (GEOM) v = keep4(v, *reinterpret_cast<const unsigned *>(P.mask + cm), M_XOPEN);
                if (ib == 0) {                                   // maintain the lower ghost plane of vx
                    float4 o = v;
                    for (int s = 0; s < P.n_sponge; s++) o = mul4s(o, P.decx[s][-1]);
                    st4(F.vx_out + cm, sel4(e0, e1, e2, e3, o, z4));
                }
            }
            vxp[r] = v;
        }
    }

    // ---- march ----------------------------------------------------------------------------
    for (int i = ib; i < ie; i++) {
        const long long base = (long long)i * P.plane + col;
        const bool upd_x = (i < P.nx - 1) || P.has_upper;
        const float cx = UNI ? P.cv_uni : (upd_x ? P.cvx[i] : 0.0f);
        const float icx = UNI ? 1.0f : P.icx[i];
        const float dx0 = (P.n_sponge > 0) ? P.decx[0][i] : 1.0f;

        // loads (all issued before use)
        float4 pn[RJ], vx[RJ], vy[RJ + 1], vz[RJ];
        unsigned mk[RJ + 1];
        unsigned ma[RJ];                                         // ADEX: mask word carrying the ADE bits of the row's four cells
        float p_hi[RJ], p_lo[RJ], vz_lo[RJ];
        uint8_t m_lo[RJ];
#pragma unroll
        for (int r = 0; r < RJ; r++) {
            const bool ok = row_ok[r + 1] && lane_ok;
            const long long c = base + (long long)r * P.pitch;
            if (ADEX) ma[r] = (ok && ((adex_rows >> r) & 1u) && i >= P.bi0 && i < P.bi1) ? *reinterpret_cast<const unsigned *>(P.ade_mask + c) : 0u;
            pn[r] = (ok && upd_x) ? ld4(F.p_in + c + P.plane) : z4;
            vx[r] = ok ? ld4(F.vx_in + c) : z4;
            vz[r] = ok ? ld4(F.vz_in + c) : z4;
            p_hi[r] = (edge_hi && row_ok[r + 1]) ? F.p_in[c + 4] : 0.0f;
            p_lo[r] = (edge_lo && row_ok[r + 1]) ? F.p_in[c - 1] : 0.0f;
            vz_lo[r] = (edge_lo && row_ok[r + 1]) ? F.vz_in[c - 1] : 0.0f;
            if (GEOM) m_lo[r] = (edge_lo && row_ok[r + 1]) ? P.mask[c - 1] : (uint8_t)0x0F;
        }
#pragma unroll
        for (int r = -1; r < RJ; r++) {
            const bool ok = row_ok[r + 1] && lane_ok;
            const long long c = base + (long long)r * P.pitch;
            vy[r + 1] = ok ? ld4(F.vy_in + c) : z4;
            if (GEOM) mk[r + 1] = ok ? *reinterpret_cast<const unsigned *>(P.mask + c) : ALL_OPEN;
        }
        // warp-uniform fast path: nothing solid or rigid in this warp's cells of this plane
        bool masked = false;
        if (GEOM) {
            bool open = true;
#pragma unroll
            for (int r = -1; r < RJ; r++) open = open && ((mk[r + 1] & ALL_OPEN) == ALL_OPEN);      // upper nibble: ADE bits
#pragma unroll
            for (int r = 0; r < RJ; r++) open = open && ((m_lo[r] & 0x0F) == 0x0F);
            masked = !__all_sync(FULL, open);
        }
        pc[0] = (row_ok[0] && lane_ok) ? ld4(F.p_in + base - P.pitch) : z4;
        pc[RJ + 1] = (row_ok[RJ + 1] && lane_ok) ? ld4(F.p_in + base + (long long)RJ * P.pitch) : z4;

        // undamped, rigid-masked y faces for rows -1 .. RJ-1
        float4 vyn[RJ + 1];
#pragma unroll
        for (int r = -1; r < RJ; r++) {
            float4 v = vy[r + 1];
            if (row_ok[r + 1] && (j0 + r < ny - 1)) {
                v = add4(v, mul4s(sub4(pc[r + 2], pc[r + 1]), cvy[r + 1]));
                if (GEOM && masked) v = keep4(v, mk[r + 1], M_YOPEN);
            }
            vyn[r + 1] = v;                                      // row -1 outside the grid stays 0
        }

#pragma unroll
        for (int r = 0; r < RJ; r++) {
            const float4 p = pc[r + 1];
            // x face
            float4 vxn = vx[r];
            if (upd_x) {
                vxn = add4(vxn, mul4s(sub4(pn[r], p), cx));
                if (GEOM && masked) vxn = keep4(vxn, mk[r + 1], M_XOPEN);
            }
            // z faces: p[k+1] from the next lane (or the next strip)
            float p_next = __shfl_down_sync(FULL, p.x, 1);
            if (lane == 31) p_next = p_hi[r];
            const float4 pk1 = make_float4(p.y, p.z, p.w, p_next);
            float4 vzn = vz[r];
            {
                const float4 upd = add4(vzn, mul4(sub4(pk1, p), cvz4));
                vzn = sel4(u0, u1, u2, u3, upd, vzn);
                if (GEOM && masked) {      // only updated faces are zeroed; the last face keeps its value
                    const unsigned m = mk[r + 1] | (u0 ? 0u : 0x08u) | (u1 ? 0u : 0x0800u) | (u2 ? 0u : 0x080000u) |
                                       (u3 ? 0u : 0x08000000u);
                    vzn = keep4(vzn, m, M_ZOPEN);
                }
            }
            // z face k0-1: previous lane's .w, or recomputed from the previous strip's values
            float vz_prev = __shfl_up_sync(FULL, vzn.w, 1);
            if (lane == 0 || (FLAT && k0 == 0)) {                // first lane of the warp, or (flat) first cells of a row
                vz_prev = 0.0f;
                if (edge_lo) {
                    vz_prev = vz_lo[r] + cvz_lo * (p.x - p_lo[r]);
                    if (GEOM && masked && !(m_lo[r] & M_ZOPEN)) vz_prev = 0.0f;
                }
            }
            const float4 vzm = make_float4(vz_prev, vzn.x, vzn.y, vzn.z);
            // divergence and pressure
            float4 ddx = sub4(vxn, vxp[r]);
            float4 ddy = sub4(vyn[r + 1], vyn[r]);
            float4 ddz = sub4(vzn, vzm);
            if (!UNI) { ddx = mul4s(ddx, icx); ddy = mul4s(ddy, icy[r]); ddz = mul4(ddz, icz4); }
            float4 pnew = add4(p, mul4s(add4(add4(ddx, ddy), ddz), P.cp));
            if (GEOM && masked) pnew = keep4(pnew, mk[r + 1], M_AIR);
            // sponge
            float4 ox = vxn, oy = vyn[r + 1], oz = vzn;
            if (P.n_sponge > 0) {
                ox = mul4s(ox, dx0); oy = mul4s(oy, dy0[r]); oz = mul4(oz, dz0);
                pnew = mul4(mul4s(mul4s(pnew, dx0), dy0[r]), dz0);
                _Pragma("unroll 1")                              // rare (several sponge objects): keep it out of the hot code
                for (int s = 1; s < P.n_sponge; s++) {
                    const float dxs = P.decx[s][i];
                    const float dys = row_ok[r + 1] ? P.decy[s][j0 + r] : 1.0f;
                    const float4 dzs = lane_ok ? ld4(P.decz[s] + k0) : f4(1.0f);
                    ox = mul4s(ox, dxs); oy = mul4s(oy, dys); oz = mul4(oz, dzs);
                    pnew = mul4(mul4s(mul4s(pnew, dxs), dys), dzs);
                }
            }
            if (row_ok[r + 1] && lane_ok && !((BOXM == 1 || RTBOX) && ((skip_rows >> r) & 1u))) {
                const long long c = base + (long long)r * P.pitch;
                float4 pst = sel4(e0, e1, e2, e3, pnew, z4);
                if (FUSE && inl_mask) {                          // float64 add, fp32 store (solver.py:2421), list order
                    for (int q = 0; q < P.n_inline; q++)
                        if (((inl_mask >> q) & 1u) && P.inl_i[q] == i && P.inl_j[q] == j0 + r) {
                            const double w = __dmul_rn(P.src_row[P.inl_src[q]], P.inl_weight[q]);
                            const int e = P.inl_k[q] - k0;
                            if (e == 0) pst.x = (float)((double)pst.x + w);
                            else if (e == 1) pst.y = (float)((double)pst.y + w);
                            else if (e == 2) pst.z = (float)((double)pst.z + w);
                            else pst.w = (float)((double)pst.w + w);
                        }
                }
                if (ADEX && (ma[r] & 0x80808080u)) {
                    // some of the four cells carry a dispersive material: their p, and the + faces that K2b corrects, are
                    // written by the list kernels running beside this launch -- store the rest element by element
                    const unsigned a = ma[r] & 0x80808080u;
                    const unsigned sx = upd_x ? (a & ((ma[r] & 0x10101010u) << 3)) : 0u;
                    const unsigned sy = (j0 + r < ny - 1) ? (a & ((ma[r] & 0x20202020u) << 2)) : 0u;
                    const unsigned sz = a & ((ma[r] & 0x40404040u) << 1) & uw;
                    const float4 sxv = sel4(e0, e1, e2, e3, ox, z4), syv = sel4(e0, e1, e2, e3, oy, z4), szv = sel4(e0, e1, e2, e3, oz, z4);
                    const float pe[4] = {pst.x, pst.y, pst.z, pst.w}, xe[4] = {sxv.x, sxv.y, sxv.z, sxv.w};
                    const float ye[4] = {syv.x, syv.y, syv.z, syv.w}, ze[4] = {szv.x, szv.y, szv.z, szv.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const unsigned bit = 0x80u << (8 * e);
                        if (!(a & bit)) F.p_out[c + e] = pe[e];
                        if (!(sx & bit)) F.vx_out[c + e] = xe[e];
                        if (!(sy & bit)) F.vy_out[c + e] = ye[e];
                        if (!(sz & bit)) F.vz_out[c + e] = ze[e];
                    }
                } else {
                st4(F.p_out + c, pst);
                if (PEER && P.peer_lo_p && i == 0) st4(P.peer_lo_p + (c - base) + col, pst);           // NVLink peer store
                if (PEER && P.peer_hi_p && i == P.nx - 1) st4(P.peer_hi_p + (c - base) + col, pst);
                st4(F.vx_out + c, sel4(e0, e1, e2, e3, ox, z4));
                st4(F.vy_out + c, sel4(e0, e1, e2, e3, oy, z4));
                st4(F.vz_out + c, sel4(e0, e1, e2, e3, oz, z4));
                }
            }
            vxp[r] = vxn;
            pc[r + 1] = pn[r];
        }
    }
}

// Register target per variant.  The kernel is bound by memory latency x bytes in flight, and how ptxas orders the ~20 loads
// of a plane changes with the register budget it aims at -- by up to 6 % either way, differently for every variant, with
// identical instruction counts.  The table below is measured, variant by variant, with tools/k1_ab.cu on B200 (CUDA 12.9;
// 1024x512x512 and 256x2048x2048, Gcell-updates/s, unspecified / 2 / 3 blocks per SM):
//   plain uniform              203.0 / 205.9 / 202.3      uniform + solids          188.6 / 198.4 / 164.3
//   spacing tables             200.9 / 194.2 / 205.1      tables + solids (config 4) 187.9 / 166.8 / 187.9  (199.1 with RTBOX)
//   uniform + peer stores      199.3 / 206.2 / 203.9      2 rows, peer stores        186.9 / 200.6 / 182.3
// 0 = leave it to the compiler (every variant not listed, among them the box modes with solids: 188.9 / 152.8 / 161.1).
template <int RJ, bool GEOM, bool UNI, bool PEER, bool FUSE, bool FLAT, int BOXM>
constexpr int k1_min_blocks()
{
#ifdef SB_K1_MINB
    return SB_K1_MINB;                                           // experiments (tools/k1_ab.cu)
#else
    if (BOXM != 0 && RJ == 1 && !GEOM && UNI && !PEER && !FUSE && !FLAT) return 2;   // config 3: beside the list kernels (BOXM 3)
                                                                                       // 202.4 / 204.7 / 202.6, beside K1-ADE (BOXM 1) 197.9 / 205.2 / 204.1
    if (FLAT || BOXM != 0 || FUSE) return 0;
    if (RJ == 1) return (GEOM && !UNI) ? 0 : (!GEOM && !UNI) ? (PEER ? 0 : 3) : ((GEOM && PEER) ? 0 : 2);
    return (PEER && UNI && !GEOM) ? 2 : 0;
#endif
}

template <int RJ, bool GEOM, bool UNI, bool PEER, bool FUSE, bool FLAT = false, int BOXM = 0>
__global__ void __launch_bounds__(256, k1_min_blocks<RJ, GEOM, UNI, PEER, FUSE, FLAT, BOXM>()) k1_step_march(StepParams P)
{
    const FieldSet F{P.p_in, P.vx_in, P.vy_in, P.vz_in, P.p_out, P.vx_out, P.vy_out, P.vz_out};
    k1_tile<RJ, GEOM, UNI, PEER, FUSE, FLAT, BOXM>(P, F, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z);
}

// ------------------------------------------------------------------------------------------
// Face-mask derivation (replaces precompute_boundary_cells, boundaries.cpp:13-64).
// geom: dense uint8 with the slab's live ghost planes, plane index gi = i + has_lower.
// ------------------------------------------------------------------------------------------
__global__ void k_build_mask(const uint8_t *geom, uint8_t *mask, int nx, int ny, int nz, int pitch,
                             long long plane, int has_lower, int has_upper, int rigid)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int i = (int)blockIdx.z - 1;                            // -1 .. nx
    if (k >= nz) return;
    const long long c = (long long)i * plane + (long long)j * pitch + k;
    const uint8_t keep = mask[c] & 0xF0;                          // ADE bits live in the upper nibble (k_ade_bits)
    if ((i < 0 && !has_lower) || (i >= nx && !has_upper)) { mask[c] = keep; return; }
    auto G = [&](int ii, int jj, int kk) -> bool {                // geom == nullptr: all air
        return !geom || geom[((long long)(ii + has_lower) * ny + jj) * nz + kk] != 0;
    };
    const bool air = G(i, j, k);
    uint8_t m = air ? M_AIR : 0;
    const bool have_xn = (i + 1 < nx) || (i + 1 == nx && has_upper);
    if (!rigid || (air && (!have_xn || G(i + 1, j, k)))) m |= M_XOPEN;
    if (!rigid || (air && (j + 1 >= ny || G(i, j + 1, k)))) m |= M_YOPEN;
    if (!rigid || (air && (k + 1 >= nz || G(i, j, k + 1)))) m |= M_ZOPEN;
    mask[c] = m | keep;
}

// what K3 needs to keep the neighbours' ghosts and flags current (all nullptr on a single GPU)
struct PeerLink {
    float *peer_lo_p, *peer_hi_p;        // neighbour ghost planes of the set just written
    int *sig_lo, *sig_hi;                // neighbours' flags to bump once this step is complete
    int *step_global;                    // device counter of completed steps
    int nx; long long plane;
};
__device__ __forceinline__ void mirror_injection(const PeerLink &L, int field, long long off, float v)
{
    if (field != 0) return;
    const long long i = off / L.plane, rem = off - i * L.plane;
    if (L.peer_lo_p && i == 0) L.peer_lo_p[rem] = v;
    if (L.peer_hi_p && i == L.nx - 1) L.peer_hi_p[rem] = v;
}
__device__ __forceinline__ void signal_step_done(const PeerLink &L)
{
    const int done = *L.step_global + 1;
    __threadfence_system();
    if (L.sig_lo) st_release_sys(L.sig_lo, done);
    if (L.sig_hi) st_release_sys(L.sig_hi, done);
    *L.step_global = done;
}

// ------------------------------------------------------------------------------------------
// K4: one-plane boundary updates applied to p after the sponge, in boundary-list order:
//   first-order Mur ABC        boundaries/_boundaries.py:476-513   p_b = prev + c*(p_i - p_b); prev = p_i
//   RadiationImpedance         boundaries/_boundaries.py:700-760   abc as above; p_b = R*p_i + (1-R)*abc
// Mixed precision exactly as NumPy evaluates the reference expressions: the difference is fp32, the
// Mur coefficient and (1-R) are float64, R*p_i is fp32 when R is a Python float ("weak" scalar) and
// float64 when it is a NumPy float64; the sum is rounded to fp32 once on store.
// ------------------------------------------------------------------------------------------
constexpr int SB_MAX_PLANE_OPS = 8;          // planes the chunk kernels (K5, K6) apply themselves

struct PlaneOp {
    int axis, side;              // axis 0/1/2, side 0 = low face, 1 = high face
    int kind;                    // 0 = Mur, 1 = radiation impedance
    int weak_r;                  // R is a Python float -> fp32 product
    double mur, R, one_minus_R;
    float r32;
    float *prev;                 // previous interior-neighbour plane, n_a * n_b floats
};

// the new face value from the previous interior neighbour, the face cell `pb` and its interior neighbour `pi` after the
// pressure update (boundaries/_boundaries.py:476-513 Mur, :700-760 radiation impedance; float64 intermediates)
__device__ __forceinline__ float plane_op_value(const PlaneOp &op, float prev, float pb, float pi)
{
    const double abc = __dadd_rn((double)prev, __dmul_rn(op.mur, (double)(pi - pb)));
    if (op.kind == 0) return (float)abc;
    const double rigid = op.weak_r ? (double)(op.r32 * pi) : __dmul_rn(op.R, (double)pi);
    return (float)__dadd_rn(rigid, __dmul_rn(op.one_minus_R, abc));
}

__global__ void k4_plane_op(PlaneOp op, float *p, int nx, int ny, int nz, int pitch, long long plane, PeerLink L)
{
    const int n[3] = {nx, ny, nz};
    const int a_ax = op.axis == 0 ? 1 : 0, b_ax = op.axis == 2 ? 1 : 2;      // the two in-plane axes
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= n[b_ax] || a >= n[a_ax]) return;
    int c[3];
    c[a_ax] = a; c[b_ax] = b;
    c[op.axis] = op.side ? n[op.axis] - 1 : 0;
    const long long ib = (long long)c[0] * plane + (long long)c[1] * pitch + c[2];
    c[op.axis] = op.side ? n[op.axis] - 2 : 1;
    const long long ii = (long long)c[0] * plane + (long long)c[1] * pitch + c[2];
    const long long t = (long long)a * n[b_ax] + b;
    const float pb = p[ib], pi = p[ii];
    const float out = plane_op_value(op, op.prev[t], pb, pi);
    p[ib] = out;
    mirror_injection(L, 0, ib, out);          // a y / z face cell on a cut plane: keep the neighbour's ghost current
    op.prev[t] = pi;
}

// ------------------------------------------------------------------------------------------
// K3: source injection and probe / microphone recording (core/solver.py:2386-2439,
// microphones.cpp:82-116).  `step_ctr` is a device counter so the kernels can live in a
// replayed CUDA graph; the recording kernel advances it.
// ------------------------------------------------------------------------------------------
struct SourceTable {
    int n_sources, n_cells;
    const long long *cell_off;          // padded-layout offset of each cell
    const int *start, *src_id, *field;
    const double *weight;
};

__global__ void k3_inject(SourceTable T, float *p, float *vx, float *vy, float *vz,
                          const double *src_vals, const int *step_ctr, PeerLink L)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= T.n_cells) return;
    const double *w = src_vals + (long long)(*step_ctr) * T.n_sources;
    const long long off = T.cell_off[u];
    for (int e = T.start[u]; e < T.start[u + 1]; e++) {
        float *f = T.field[e] == 0 ? p : (T.field[e] == 1 ? vx : (T.field[e] == 2 ? vy : vz));
        // float64 add, float32 store (solver.py:2421); w*weight is one fp64 multiply (solver.py:2404)
        const float v = (float)((double)f[off] + __dmul_rn(w[T.src_id[e]], T.weight[e]));
        f[off] = v;
        mirror_injection(L, T.field[e], off, v);
    }
}

__global__ void k3_record(FieldPtrs F, int n_probes, const long long *probe_off,
                          int n_mics, const int *mic_field, const long long *mic_off8, const float *mic_w8,
                          float *record_out, int *step_ctr)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_rec = n_probes + n_mics;
    const int step = *step_ctr;
    if (t < n_probes) {
        record_out[(long long)step * n_rec + t] = F.f[0][probe_off[t]];
    } else if (t < n_rec) {
        record_out[(long long)step * n_rec + t] = gather8(F, mic_field, mic_off8, mic_w8, t - n_probes);
    }
}

// records one step into an explicit row (tail of a chunk in single-kernel-step mode)
__global__ void k3_record_row(FieldPtrs F, int n_probes, const long long *probe_off, int n_mics, const int *mic_field,
                              const long long *mic_off8, const float *mic_w8, float *rec_row)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_probes) rec_row[t] = F.f[0][probe_off[t]];
    else if (t < n_probes + n_mics) rec_row[t] = gather8(F, mic_field, mic_off8, mic_w8, t - n_probes);
}

__global__ void k3_advance(int *step_ctr, PeerLink L) { *step_ctr += 1; signal_step_done(L); }

// Small-problem variant: one block does inject -> record -> advance (saves two launches).
__global__ void __launch_bounds__(1024) k3_small(SourceTable T, float *p, float *vx, float *vy, float *vz,
                                                 const double *src_vals, int n_probes, const long long *probe_off,
                                                 int n_mics, const int *mic_field, const long long *mic_off8,
                                                 const float *mic_w8, float *record_out, int *step_ctr, PeerLink L)
{
    const int step = *step_ctr;
    for (int u = threadIdx.x; u < T.n_cells; u += blockDim.x) {
        const double *w = src_vals + (long long)step * T.n_sources;
        const long long off = T.cell_off[u];
        for (int e = T.start[u]; e < T.start[u + 1]; e++) {
            float *f = T.field[e] == 0 ? p : (T.field[e] == 1 ? vx : (T.field[e] == 2 ? vy : vz));
            const float v = (float)((double)f[off] + __dmul_rn(w[T.src_id[e]], T.weight[e]));
            f[off] = v;
            mirror_injection(L, T.field[e], off, v);
        }
    }
    __syncthreads();
    const int n_rec = n_probes + n_mics;
    for (int t = threadIdx.x; t < n_rec; t += blockDim.x) {
        if (t < n_probes) {
            record_out[(long long)step * n_rec + t] = p[probe_off[t]];
        } else {
            const FieldPtrs F{{p, vx, vy, vz}};
            record_out[(long long)step * n_rec + t] = gather8(F, mic_field, mic_off8, mic_w8, t - n_probes);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { *step_ctr = step + 1; signal_step_done(L); }
}

// ------------------------------------------------------------------------------------------
// K2: ADE recursions over the compact list of material cells (ade.cpp:25-692), in the order of
// core/solver.py:2135-2193.  K1 has already written the material-free update of every cell to
// the "out" set; because the "in" set is untouched, K2b can recompute the few cells and faces
// that the auxiliary fields change and overwrite K1's values there:
//   faces whose two cells carry the same material with density poles  (ade.cpp:228-401)
//   pressure of every material cell                                   (ade.cpp:403-473)
// ------------------------------------------------------------------------------------------
constexpr int MAX_POLES = 16;
struct PoleDev {
    int mat_id, is_lorentz, target;      // target 0 = density, 1 = modulus
    float c0, c1, c2;                    // Debye: alpha, beta.  Lorentz: a, b, d
    float vcoef;                         // -dt / rho_inf * inv_dx          (ade.cpp:242)
    float pcoef;                         // -K_inf * dt                     (ade.cpp:417)
};
struct AdeTable {
    int n_cells, n_poles;
    const long long *cell_off;           // padded offset of each material cell
    const int *cell_ijk;                 // 3 ints per cell
    const uint8_t *cell_mat;
    const int *nbr;                      // [6][n_cells]: +x,+y,+z,-x,-y,-z slot of a SAME-material neighbour or -1
    float *J, *Jp;                       // [n_poles][n_cells]
    float inv_dx;                        // uniform-grid divergence scale (solver.py:3158)
    // dense layout (materials that fill most of their bounding box): n_cells = cells of the box, a cell's slot is its
    // box-local index, neighbours are found geometrically -- no index lists, no dependent slot -> J loads
    int dense, bi0, bj0, bk0, bx, by, bz;
    const uint8_t *mat_box;              // material id per box cell, 0 where the material has no poles
    PoleDev poles[MAX_POLES];
};

__device__ __forceinline__ void ade_pole_update(const PoleDev &q, float *J, float *Jp, long long s, float src)
{
    if (!q.is_lorentz) {
        J[s] = q.c0 * J[s] + q.c1 * src;                                   // ade.cpp:57-59
    } else {
        const float jo = J[s], jpo = Jp[s];
        const float jn = q.c0 * jo + q.c1 * jpo + q.c2 * src;               // ade.cpp:158-160
        Jp[s] = jo; J[s] = jn;
    }
}

// K2a: density poles, source = p of the previous step (solver.py:2138-2139)
__global__ void k2a_density(AdeTable A, const float *p_in)
{
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < A.n_cells; s += gridDim.x * blockDim.x) {
        const int mat = A.cell_mat[s];
        const float src = p_in[A.cell_off[s]];
        for (int q = 0; q < A.n_poles; q++) {
            const PoleDev &Q = A.poles[q];
            if (Q.target == 0 && Q.mat_id == mat)
                ade_pole_update(Q, A.J + (long long)q * A.n_cells, A.Jp + (long long)q * A.n_cells, s, src);
        }
    }
}

// everything the auxiliary fields change at material cell (i,j,k) with slot s: its three + faces, its pressure, its
// modulus poles; sxp .. szm = slots of the +x,+y,+z,-x,-y,-z neighbours if they carry the same material, else -1
__device__ __forceinline__ void ade_fixup_cell(const StepParams &P, const AdeTable &A, int s, long long c, int i, int j, int k,
                                               int mat, int sxp, int syp, int szp, int sxm, int sym, int szm)
{
    const int n = A.n_cells;
    const bool upd_x = (i < P.nx - 1) || P.has_upper, upd_y = j < P.ny - 1, upd_z = k < P.nz - 1;
    const bool sub_x = (i > 0) || P.has_lower, sub_y = j > 0, sub_z = k > 0;
    // the six undamped face velocities around the cell, material-free part (fdtd_step.cpp:34-80 / 235-307)
    const float pc = P.p_in[c];
    float vxn = P.vx_in[c], vyn = P.vy_in[c], vzn = P.vz_in[c], vxm = 0.0f, vym = 0.0f, vzm = 0.0f;
    if (upd_x) vxn = vxn + P.cvx[i] * (P.p_in[c + P.plane] - pc);
    if (upd_y) vyn = vyn + P.cvy[j] * (P.p_in[c + P.pitch] - pc);
    if (upd_z) vzn = vzn + P.cvz[k] * (P.p_in[c + 1] - pc);
    if (sub_x) vxm = P.vx_in[c - P.plane] + P.cvx[i - 1] * (pc - P.p_in[c - P.plane]);
    if (sub_y) vym = P.vy_in[c - P.pitch] + P.cvy[j - 1] * (pc - P.p_in[c - P.pitch]);
    if (sub_z) vzm = P.vz_in[c - 1] + P.cvz[k - 1] * (pc - P.p_in[c - 1]);
    // density-pole corrections of the faces whose two cells carry this material, pole after pole (ade.cpp:228-401:
    // v += vcoef * (J[upper cell] - J[lower cell])); one pass over the poles serves all six faces
    const bool cxp = upd_x && sxp >= 0, cyp = upd_y && syp >= 0, czp = upd_z && szp >= 0;
    const bool cxm = sub_x && sxm >= 0, cym = sub_y && sym >= 0, czm = sub_z && szm >= 0;
    if (cxp || cyp || czp || cxm || cym || czm)
        _Pragma("unroll 1")
        for (int q = 0; q < A.n_poles; q++) {
            const PoleDev &Q = A.poles[q];
            if (Q.target != 0 || Q.mat_id != mat) continue;
            const float *J = A.J + (long long)q * n;
            const float js = J[s], vc = Q.vcoef;
            if (cxp) vxn = vxn + vc * (J[sxp] - js);
            if (cyp) vyn = vyn + vc * (J[syp] - js);
            if (czp) vzn = vzn + vc * (J[szp] - js);
            if (cxm) vxm = vxm + vc * (js - J[sxm]);
            if (cym) vym = vym + vc * (js - J[sym]);
            if (czm) vzm = vzm + vc * (js - J[szm]);
        }
    if (P.mask) {                                                           // rigid faces (boundaries.cpp:66-89): updated faces only
        const uint8_t m = P.mask[c];
        if (upd_x && !(m & M_XOPEN)) vxn = 0.0f;
        if (upd_y && !(m & M_YOPEN)) vyn = 0.0f;
        if (upd_z && !(m & M_ZOPEN)) vzn = 0.0f;
        if (sub_x && !(P.mask[c - P.plane] & M_XOPEN)) vxm = 0.0f;
        if (sub_y && !(P.mask[c - P.pitch] & M_YOPEN)) vym = 0.0f;
        if (sub_z && !(P.mask[c - 1] & M_ZOPEN)) vzm = 0.0f;
    }
    float ddx = vxn, ddy = vyn, ddz = vzn;
    if (sub_x) {
        ddx = vxn - vxm;
        if (i == 0 && sxm >= 0) {            // the redundantly kept ghost face vx[-1] carries the correction too
            float og = vxm;
            _Pragma("unroll 1")
            for (int sp = 0; sp < P.n_sponge; sp++) og = og * P.decx[sp][-1];
            P.vx_out[c - P.plane] = og;
        }
    }
    if (sub_y) ddy = vyn - vym;
    if (sub_z) ddz = vzn - vzm;

    // divergence for the modulus poles: d = dx*ix; d += dy*iy; d += dz*iz   (ade.cpp:479-692)
    const float ix = P.icx ? P.icx[i] : A.inv_dx, iy = P.icy ? P.icy[j] : A.inv_dx, iz = P.icz ? P.icz[k] : A.inv_dx;
    float div = ddx * ix;
    div = div + ddy * iy;
    div = div + ddz * iz;
    _Pragma("unroll 1")
    for (int q = 0; q < A.n_poles; q++) {
        const PoleDev &Q = A.poles[q];
        if (Q.target == 1 && Q.mat_id == mat)
            ade_pole_update(Q, A.J + (long long)q * n, A.Jp + (long long)q * n, s, div);
    }
    // pressure (fdtd_step.cpp:109-211 / 345-355), then modulus correction (ade.cpp:403-473)
    float ex = ddx, ey = ddy, ez = ddz;
    if (P.icx) { ex = P.icx[i] * ddx; ey = P.icy[j] * ddy; ez = P.icz[k] * ddz; }
    float pn = P.p_in[c] + P.cp * ((ex + ey) + ez);
    const bool air = !P.mask || (P.mask[c] & M_AIR);
    if (!air) pn = 0.0f;
    _Pragma("unroll 1")
    for (int q = 0; q < A.n_poles; q++) {
        const PoleDev &Q = A.poles[q];
        if (Q.target == 1 && Q.mat_id == mat) pn = pn + Q.pcoef * A.J[(long long)q * n + s];
    }
    if (!air) pn = 0.0f;                                                    // solver.py:2193
    float ox = vxn, oy = vyn, oz = vzn;
    _Pragma("unroll 1")
    for (int sp = 0; sp < P.n_sponge; sp++) {
        const float dx = P.decx[sp][i], dy = P.decy[sp][j], dz = P.decz[sp][k];
        ox = ox * dx; oy = oy * dy; oz = oz * dz;
        pn = ((pn * dx) * dy) * dz;
    }
    P.p_out[c] = pn;
    if (P.peer_lo_p && i == 0) P.peer_lo_p[c] = pn;                           // keep the neighbours' ghost planes current
    if (P.peer_hi_p && i == P.nx - 1) P.peer_hi_p[c - (long long)i * P.plane] = pn;
    if (sxp >= 0 && upd_x) P.vx_out[c] = ox;       // only faces the ADE correction touched
    if (syp >= 0 && upd_y) P.vy_out[c] = oy;
    if (szp >= 0 && upd_z) P.vz_out[c] = oz;
}

__global__ void k2b_fixup(StepParams P, AdeTable A)
{
    const int n = A.n_cells;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const long long c = A.cell_off[s];
        const int i = A.cell_ijk[3 * s], j = A.cell_ijk[3 * s + 1], k = A.cell_ijk[3 * s + 2];
        if (i < 0 || i >= P.nx) continue;    // ghost-plane cell of a slab: only its J is kept here (K2a), its owner does the rest
        ade_fixup_cell(P, A, s, c, i, j, k, A.cell_mat[s], A.nbr[s], A.nbr[n + s], A.nbr[2 * n + s],
                       A.nbr[3 * n + s], A.nbr[4 * n + s], A.nbr[5 * n + s]);
    }
}

// dense layout: one thread per cell of the materials' bounding box, k fastest
__global__ void k2a_density_dense(AdeTable A, const float *p_in, int pitch, long long plane)
{
    const int kk = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y, ii = blockIdx.z;    // grid = (k blocks, by, bx)
    if (kk >= A.bz) return;
    const int s = (ii * A.by + jj) * A.bz + kk;
    const int mat = A.mat_box[s];
    if (!mat) return;
    const float src = p_in[(long long)(A.bi0 + ii) * plane + (long long)(A.bj0 + jj) * pitch + (A.bk0 + kk)];
    for (int q = 0; q < A.n_poles; q++) {
        const PoleDev &Q = A.poles[q];
        if (Q.target == 0 && Q.mat_id == mat)
            ade_pole_update(Q, A.J + (long long)q * A.n_cells, A.Jp + (long long)q * A.n_cells, s, src);
    }
}

__global__ void k2b_fixup_dense(StepParams P, AdeTable A)
{
    const int kk = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y, ii = blockIdx.z;    // grid = (k blocks, by, bx)
    if (kk >= A.bz) return;
    const int s = (ii * A.by + jj) * A.bz + kk;
    const int mat = A.mat_box[s];
    if (!mat) return;
    const int i = A.bi0 + ii, j = A.bj0 + jj, k = A.bk0 + kk;
    const int sj = A.bz, si = A.by * A.bz;
    const uint8_t *m = A.mat_box;
    ade_fixup_cell(P, A, s, (long long)i * P.plane + (long long)j * P.pitch + k, i, j, k, mat,
                   (ii + 1 < A.bx && m[s + si] == mat) ? s + si : -1, (jj + 1 < A.by && m[s + sj] == mat) ? s + sj : -1,
                   (kk + 1 < A.bz && m[s + 1] == mat) ? s + 1 : -1,
                   (ii > 0 && m[s - si] == mat) ? s - si : -1, (jj > 0 && m[s - sj] == mat) ? s - sj : -1,
                   (kk > 0 && m[s - 1] == mat) ? s - 1 : -1);
}

// ------------------------------------------------------------------------------------------
// K1-ADE: the marching kernel with the auxiliary-field recursions folded in, for the tiles that meet the bounding box of
// the dispersive materials (the plain K1 launch of the same step skips that box, StepParams::box_mode).  A material
// cell is computed ONCE, in the order of core/solver.py:2135-2193:
//   density poles  J <- alpha J + beta p   (ade.cpp:25-68, 118-170)      from the input set's p
//   velocity       v += cv dp              (fdtd_step.cpp:34-80)          then, pole after pole,
//                  v += vcoef (J[c+] - J[c])  on faces whose two cells carry the pole's material (ade.cpp:228-401)
//   rigid faces, divergence, modulus poles J <- f(J, div) (ade.cpp:70-112, 172-222, 479-692)
//   pressure       p += cp div;  p += pcoef J  per modulus pole (ade.cpp:403-473);  p = 0 in solids;  sponge
// The density fields of the +x / +-y / +-z neighbours are recomputed from their old values and the neighbours' input
// pressure (same operations as their owners perform), so the density J arrays are double-buffered (Lorentz poles
// rotate three buffers: J, J_prev, J_new) -- a neighbour must never see a half-updated field.  Modulus fields belong
// to one cell only and are updated in place.  Which cells and faces take part is read from the upper nibble of the
// mask byte (M_ADE, M_?SAME); the material id array is read only when more than one material carries poles.
// J arrays are dense over the bounding box of the pole's material in i and j and over whole padded rows in k:
//   index(i, j, k) = ((i - bi0) * bnj + (j - bj0)) * pitch + k.
// ------------------------------------------------------------------------------------------
struct AdeFused {
    int n_poles, multi;
    const uint8_t *mat;                  // material id per cell (padded field layout, plane-0 pointer); multi only
    float inv_dx;
    PoleDev poles[MAX_POLES];
    int bi0[MAX_POLES], bj0[MAX_POLES], bni[MAX_POLES], bnj[MAX_POLES];
    const float *Jin[MAX_POLES], *Jpin[MAX_POLES];
    float *Jout[MAX_POLES], *Jpout[MAX_POLES];       // modulus poles: the same buffers as Jin / Jpin (in place)
};

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// byte e of w has bit 7 set -> element e takes part
SB_HD float4 corr4(float4 v, unsigned w, float vc, float4 hi, float4 lo)
{
    return make_float4((w & 0x80u) ? v.x + vc * (hi.x - lo.x) : v.x, (w & 0x8000u) ? v.y + vc * (hi.y - lo.y) : v.y,
                       (w & 0x800000u) ? v.z + vc * (hi.z - lo.z) : v.z, (w & 0x80000000u) ? v.w + vc * (hi.w - lo.w) : v.w);
}
SB_HD float4 pick4(unsigned w, float4 a, float4 b)
{
    return make_float4((w & 0x80u) ? a.x : b.x, (w & 0x8000u) ? a.y : b.y, (w & 0x800000u) ? a.z : b.z, (w & 0x80000000u) ? a.w : b.w);
}
SB_HD float pole_upd(const PoleDev &Q, float J, float Jp, float src)
{
    return Q.is_lorentz ? Q.c0 * J + Q.c1 * Jp + Q.c2 * src : Q.c0 * J + Q.c1 * src;      // ade.cpp:158-160 / 57-59
}
SB_HD float4 pole_upd4(const PoleDev &Q, float4 J, float4 Jp, float4 s)
{
    if (Q.is_lorentz)                                         // one (warp-uniform) branch instead of a select per element
        return add4(add4(mul4s(J, Q.c0), mul4s(Jp, Q.c1)), mul4s(s, Q.c2));
    return add4(mul4s(J, Q.c0), mul4s(s, Q.c1));
}
// corr4 / pick4 when the warp has agreed that every element of every lane takes part (the interior of a material)
constexpr unsigned ALL_SEL = 0x80808080u;
SB_HD float4 corr4_all(float4 v, float vc, float4 hi, float4 lo) { return add4(v, mul4s(sub4(hi, lo), vc)); }
// which of the four cells of a mask word carry pole Q's material (bit 7 of each byte)
__device__ __forceinline__ unsigned ade_sel(const AdeFused &A, const PoleDev &Q, unsigned mk, unsigned matw)
{
    unsigned s = mk & 0x80808080u;
    if (A.multi) s &= __vcmpeq4(matw, (unsigned)Q.mat_id * 0x01010101u);
    return s;
}

template <bool UNI, bool FLAT>
__device__ __forceinline__ void k1_tile_ade(const StepParams &P, const AdeFused &A, const FieldSet &F, int bx, int by, int bz)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int k0, j0;
    bool warp_out;
    if (FLAT) {
        const int P4 = P.pitch >> 2, warps_x = blockDim.x >> 5;
        const long long g0 = ((long long)bx * (warps_x * blockDim.y) + threadIdx.y * warps_x + (threadIdx.x >> 5)) * 32;
        const long long groups = (long long)P.ny * P4;
        warp_out = g0 >= groups;
        const long long g = g0 + lane;
        j0 = (int)(g / P4);
        k0 = 4 * (int)(g - (long long)j0 * P4);
    } else {
        // strips start at the box, not at multiples of 128: a material 100 cells wide that straddles a multiple of 128
        // (config 3's sphere, centred at k = 256) is then one strip per row instead of two
        const int strip_k0 = P.bk0 + (bx * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 128;
        k0 = strip_k0 + lane * 4;
        j0 = by * blockDim.y + threadIdx.y;
        warp_out = strip_k0 >= P.bk1 || strip_k0 >= P.nz || j0 >= P.ny;
    }
    const int ib = P.i_begin + bz * P.chunk_i;
    const int ie = min(ib + P.chunk_i, P.i_end);
    if (warp_out || ib >= ie) return;
    const int nz = P.nz, ny = P.ny;
    const bool lane_ok = k0 < nz;
    const bool row0 = j0 < ny, rowm = j0 - 1 >= 0 && j0 - 1 < ny, rowp = j0 + 1 < ny;      // rows j0, j0-1, j0+1 exist
    const bool own = lane_ok && row0 && k0 >= P.bk0 && k0 < P.bk1 && j0 >= P.bj0 && j0 < P.bj1;   // this launch stores the box only
    if (!__any_sync(FULL, own)) return;
    const bool e0 = k0 < nz, e1 = k0 + 1 < nz, e2 = k0 + 2 < nz, e3 = k0 + 3 < nz;
    const bool u0 = k0 < nz - 1, u1 = k0 + 1 < nz - 1, u2 = k0 + 2 < nz - 1, u3 = k0 + 3 < nz - 1;
    const unsigned uw = (u0 ? 0x80u : 0u) | (u1 ? 0x8000u : 0u) | (u2 ? 0x800000u : 0u) | (u3 ? 0x80000000u : 0u);
    const bool edge_hi = (lane == 31) && (k0 + 4 < nz);
    const bool edge_lo = (lane == 0) && (k0 > 0);
    const bool first_in_row = lane == 0 || (FLAT && k0 == 0);
    const float4 z4 = f4(0.0f);
    const bool ok = row0 && lane_ok;

    const float4 cvz4 = UNI ? f4(P.cv_uni) : (lane_ok ? ld4(P.cvz + k0) : z4);
    const float4 icz4 = (!UNI && lane_ok) ? ld4(P.icz + k0) : f4(1.0f);
    const float4 dz0 = (P.n_sponge > 0 && lane_ok) ? ld4(P.decz[0] + k0) : f4(1.0f);
    const float cvz_lo = UNI ? P.cv_uni : (edge_lo ? P.cvz[k0 - 1] : 0.0f);
    const bool updm = rowm && row0, updp = row0 && j0 < ny - 1;                             // the y faces j0-1 -> j0 and j0 -> j0+1 are updated
    const float cvym = UNI ? P.cv_uni : (updm ? P.cvy[j0 - 1] : 0.0f), cvyc = UNI ? P.cv_uni : (updp ? P.cvy[j0] : 0.0f);
    const float icy = (!UNI && row0) ? P.icy[j0] : 1.0f;
    const float dy0 = (P.n_sponge > 0 && row0) ? P.decy[0][j0] : 1.0f;
    const long long col = (long long)j0 * P.pitch + k0;
    auto jidx = [&](int q, int i) -> long long {
        return ((long long)(i - A.bi0[q]) * A.bnj[q] + (j0 - A.bj0[q])) * P.pitch + k0;
    };

    // ---- prologue: p of plane ib, undamped (corrected, rigid-masked) vx of plane ib-1 ---------------------------
    if (own) {                                               // the J rows of the first two planes, on their way while the fields load
        _Pragma("unroll 1")
        for (int q = 0; q < A.n_poles; q++) {
            const int lj = j0 - A.bj0[q];
            if (A.poles[q].target > 1 || lj < 0 || lj >= A.bnj[q]) continue;
            for (int a = 0; a < 2; a++) {
                const int li = ib + a - A.bi0[q];
                if (li < 0 || li >= A.bni[q]) continue;
                const long long o = ((long long)li * A.bnj[q] + lj) * P.pitch + k0;
                prefetch_l1(A.Jin[q] + o);
                if (A.Jpin[q]) prefetch_l1(A.Jpin[q] + o);
            }
        }
    }
    float4 pc = ok ? ld4(F.p_in + (long long)ib * P.plane + col) : z4;
    float4 vxp = z4;
    if ((ib > 0 || P.has_lower) && ok) {                     // (ok is not warp-uniform: no collectives in here)
        const long long cm = (long long)(ib - 1) * P.plane + col;
        const float4 pm = ld4(F.p_in + cm);
        const float cx = UNI ? P.cv_uni : P.cvx[ib - 1];
        float4 v = add4(ld4(F.vx_in + cm), mul4s(sub4(pc, pm), cx));
        const unsigned mkm = *reinterpret_cast<const unsigned *>(P.mask + cm);
        if (mkm & 0x80808080u & ((mkm & 0x10101010u) << 3)) {
            const unsigned matw = A.multi ? *reinterpret_cast<const unsigned *>(A.mat + cm) : 0u;
            _Pragma("unroll 1")
            for (int q = 0; q < A.n_poles; q++) {
                const PoleDev &Q = A.poles[q];
                if (Q.target != 0) continue;
                const unsigned xw = ade_sel(A, Q, mkm, matw) & ((mkm & 0x10101010u) << 3);
                if (!xw) continue;
                const long long jl = jidx(q, ib - 1), jh = jl + (long long)A.bnj[q] * P.pitch;
                const float4 Jl = pole_upd4(Q, ld4(A.Jin[q] + jl), Q.is_lorentz ? ld4(A.Jpin[q] + jl) : z4, pm);
                const float4 Jh = pole_upd4(Q, ld4(A.Jin[q] + jh), Q.is_lorentz ? ld4(A.Jpin[q] + jh) : z4, pc);
                v = corr4(v, xw, Q.vcoef, Jh, Jl);
            }
        }
        vxp = keep4(v, mkm, M_XOPEN);
    }

    // ---- march -----------------------------------------------------------------------------------------------
    for (int i = ib; i < ie; i++) {
        const long long c = (long long)i * P.plane + col;
        const bool upd_x = (i < P.nx - 1) || P.has_upper;
        const float cx = UNI ? P.cv_uni : (upd_x ? P.cvx[i] : 0.0f);
        const float icx = UNI ? 1.0f : P.icx[i];
        const float dx0 = (P.n_sponge > 0) ? P.decx[0][i] : 1.0f;

        const float4 pn = (ok && upd_x) ? ld4(F.p_in + c + P.plane) : z4;
        const float4 vx = ok ? ld4(F.vx_in + c) : z4;
        const float4 vz = ok ? ld4(F.vz_in + c) : z4;
        const float4 vyc = ok ? ld4(F.vy_in + c) : z4;
        const float4 vym = (rowm && lane_ok) ? ld4(F.vy_in + c - P.pitch) : z4;
        const float4 pm = (rowm && lane_ok) ? ld4(F.p_in + c - P.pitch) : z4;
        const float4 pp = (rowp && lane_ok) ? ld4(F.p_in + c + P.pitch) : z4;
        const unsigned mkc = ok ? *reinterpret_cast<const unsigned *>(P.mask + c) : ALL_OPEN;
        const unsigned mkm = (rowm && lane_ok) ? *reinterpret_cast<const unsigned *>(P.mask + c - P.pitch) : ALL_OPEN;
        const float p_hi = (edge_hi && row0) ? F.p_in[c + 4] : 0.0f;
        const float p_lo = (edge_lo && row0) ? F.p_in[c - 1] : 0.0f;
        const float vz_lo = (edge_lo && row0) ? F.vz_in[c - 1] : 0.0f;
        const uint8_t m_lo = (edge_lo && row0) ? P.mask[c - 1] : (uint8_t)0x0F;
        // The J rows a plane needs are only known once its mask bytes have arrived, and every pole adds a dependent round
        // of loads: ask for the next plane's rows now (density poles: the row that becomes the +x neighbour, i.e. plane
        // i+2; modulus poles: plane i+1), one iteration -- far more than a DRAM round trip -- ahead of their use.
        if (own) {
            _Pragma("unroll 1")
            for (int q = 0; q < A.n_poles; q++) {
                const int ahead = A.poles[q].target == 0 ? 2 : 1;
                const int li = i + ahead - A.bi0[q], lj = j0 - A.bj0[q];
                if (A.poles[q].target > 1 || li < 0 || li >= A.bni[q] || lj < 0 || lj >= A.bnj[q]) continue;
                const long long o = ((long long)li * A.bnj[q] + lj) * P.pitch + k0;
                prefetch_l1(A.Jin[q] + o);
                if (A.Jpin[q]) prefetch_l1(A.Jpin[q] + o);
            }
        }
        const bool masked = !__all_sync(FULL, (mkc & ALL_OPEN) == ALL_OPEN && (mkm & ALL_OPEN) == ALL_OPEN && (m_lo & 0x0F) == 0x0F);
        const bool ade = __any_sync(FULL, (mkc & 0x80808080u) != 0u);

        // material-free velocity updates (fdtd_step.cpp:34-80 / 235-307); rigid faces come after the ADE corrections
        const float4 p = pc;
        float4 vyn_m = vym, vyn_c = vyc, vxn = vx, vzn = vz;
        if (updm) vyn_m = add4(vym, mul4s(sub4(p, pm), cvym));
        if (updp) vyn_c = add4(vyc, mul4s(sub4(pp, p), cvyc));
        if (upd_x) vxn = add4(vx, mul4s(sub4(pn, p), cx));
        float p_next = __shfl_down_sync(FULL, p.x, 1);
        if (lane == 31) p_next = p_hi;
        vzn = sel4(u0, u1, u2, u3, add4(vz, mul4(sub4(make_float4(p.y, p.z, p.w, p_next), p), cvz4)), vz);
        float vz_edge = edge_lo ? vz_lo + cvz_lo * (p.x - p_lo) : 0.0f;        // z face k0-1 of the previous strip's cell

        unsigned matw = 0u;
        if (ade) {
            if (A.multi && ok) matw = *reinterpret_cast<const unsigned *>(A.mat + c);
            _Pragma("unroll 1")
            for (int q = 0; q < A.n_poles; q++) {
                const PoleDev &Q = A.poles[q];
                if (Q.target != 0) continue;
                const unsigned selw = ade_sel(A, Q, mkc, matw);
                if (!__any_sync(FULL, selw != 0u)) continue;
                const bool lor = Q.is_lorentz != 0;
                const float *Jin = A.Jin[q], *Jpin = A.Jpin[q];
                const long long jb = jidx(q, i);
                const unsigned xw = upd_x ? (selw & ((mkc & 0x10101010u) << 3)) : 0u;
                const unsigned yw = updp ? (selw & ((mkc & 0x20202020u) << 2)) : 0u;
                const unsigned mw = updm ? (selw & ((mkm & 0x20202020u) << 2)) : 0u;      // the face j0-1 -> j0 belongs to the row below
                // inside a material every element of every lane takes part: plain float4 arithmetic, no selects
                const bool all_sel = __all_sync(FULL, selw == ALL_SEL), all_x = __all_sync(FULL, xw == ALL_SEL);
                const bool all_y = __all_sync(FULL, yw == ALL_SEL), all_m = __all_sync(FULL, mw == ALL_SEL);
                float4 Jn = z4;
                if (selw) {
                    Jn = pole_upd4(Q, ld4(Jin + jb), lor ? ld4(Jpin + jb) : z4, p);
                    if (!all_sel) Jn = pick4(selw, Jn, z4);
                }
                if (xw) {
                    const long long jx = jb + (long long)A.bnj[q] * P.pitch;
                    const float4 Jx = pole_upd4(Q, ld4(Jin + jx), lor ? ld4(Jpin + jx) : z4, pn);
                    vxn = all_x ? corr4_all(vxn, Q.vcoef, Jx, Jn) : corr4(vxn, xw, Q.vcoef, Jx, Jn);
                }
                if (yw) {
                    const long long jy = jb + P.pitch;
                    const float4 Jy = pole_upd4(Q, ld4(Jin + jy), lor ? ld4(Jpin + jy) : z4, pp);
                    vyn_c = all_y ? corr4_all(vyn_c, Q.vcoef, Jy, Jn) : corr4(vyn_c, yw, Q.vcoef, Jy, Jn);
                }
                if (mw) {
                    const long long jm = jb - P.pitch;
                    const float4 Jm = pole_upd4(Q, ld4(Jin + jm), lor ? ld4(Jpin + jm) : z4, pm);
                    vyn_m = all_m ? corr4_all(vyn_m, Q.vcoef, Jn, Jm) : corr4(vyn_m, mw, Q.vcoef, Jn, Jm);
                }
                const unsigned zw = selw & ((mkc & 0x40404040u) << 1) & uw;
                float Jn_next = __shfl_down_sync(FULL, Jn.x, 1);
                if (lane == 31) {
                    Jn_next = 0.0f;
                    if (edge_hi && (zw & 0x80000000u)) Jn_next = pole_upd(Q, Jin[jb + 4], lor ? Jpin[jb + 4] : 0.0f, p_hi);
                }
                vzn = __all_sync(FULL, zw == ALL_SEL) ? corr4_all(vzn, Q.vcoef, make_float4(Jn.y, Jn.z, Jn.w, Jn_next), Jn)
                                                      : corr4(vzn, zw, Q.vcoef, make_float4(Jn.y, Jn.z, Jn.w, Jn_next), Jn);
                if (edge_lo && (selw & 0x80u) && (m_lo & M_ZSAME))
                    vz_edge = vz_edge + Q.vcoef * (Jn.x - pole_upd(Q, Jin[jb - 1], lor ? Jpin[jb - 1] : 0.0f, p_lo));
                if (selw && own) st4(A.Jout[q] + jb, Jn);
            }
        }
        // rigid faces (boundaries.cpp:66-89): updated faces only
        if (masked) {
            if (updm) vyn_m = keep4(vyn_m, mkm, M_YOPEN);
            if (updp) vyn_c = keep4(vyn_c, mkc, M_YOPEN);
            if (upd_x) vxn = keep4(vxn, mkc, M_XOPEN);
            vzn = keep4(vzn, mkc | (u0 ? 0u : 0x08u) | (u1 ? 0u : 0x0800u) | (u2 ? 0u : 0x080000u) | (u3 ? 0u : 0x08000000u), M_ZOPEN);
            if (edge_lo && !(m_lo & M_ZOPEN)) vz_edge = 0.0f;
        }
        float vz_prev = __shfl_up_sync(FULL, vzn.w, 1);
        if (first_in_row) vz_prev = vz_edge;
        const float4 vzm = make_float4(vz_prev, vzn.x, vzn.y, vzn.z);
        // divergence and pressure
        const float4 ddx = sub4(vxn, vxp), ddy = sub4(vyn_c, vyn_m), ddz = sub4(vzn, vzm);
        float4 ex = ddx, ey = ddy, ez = ddz;
        if (!UNI) { ex = mul4s(ddx, icx); ey = mul4s(ddy, icy); ez = mul4(ddz, icz4); }
        float4 pnew = add4(p, mul4s(add4(add4(ex, ey), ez), P.cp));
        if (masked) pnew = keep4(pnew, mkc, M_AIR);
        if (ade) {
            // modulus poles: source = divergence accumulated as d = dx*ix; d += dy*iy; d += dz*iz (ade.cpp:479-692)
            const float ix = UNI ? A.inv_dx : icx, iy = UNI ? A.inv_dx : icy;
            const float4 iz = UNI ? f4(A.inv_dx) : icz4;
            const float4 div = add4(add4(mul4s(ddx, ix), mul4s(ddy, iy)), mul4(ddz, iz));
            _Pragma("unroll 1")
            for (int q = 0; q < A.n_poles; q++) {
                const PoleDev &Q = A.poles[q];
                if (Q.target != 1) continue;
                const unsigned selw = own ? ade_sel(A, Q, mkc, matw) : 0u;
                if (!selw) continue;
                const long long jb = jidx(q, i);
                const float4 J = ld4(A.Jin[q] + jb);
                const float4 Jp = Q.is_lorentz ? ld4(A.Jpin[q] + jb) : z4;
                const float4 Jn = pole_upd4(Q, J, Jp, div);
                if (selw == ALL_SEL) {                                       // (per lane: no collective in this loop)
                    st4(A.Jout[q] + jb, Jn);
                    if (Q.is_lorentz) st4(A.Jpout[q] + jb, J);
                    pnew = add4(pnew, mul4s(Jn, Q.pcoef));                   // p += (-K_inf dt) J   (ade.cpp:403-473)
                } else {
                    st4(A.Jout[q] + jb, pick4(selw, Jn, J));
                    if (Q.is_lorentz) st4(A.Jpout[q] + jb, pick4(selw, J, Jp));
                    pnew = corr4(pnew, selw, Q.pcoef, Jn, z4);
                }
            }
            if (masked) pnew = keep4(pnew, mkc, M_AIR);                      // solver.py:2193
        }
        // sponge
        float4 ox = vxn, oy = vyn_c, oz = vzn;
        if (P.n_sponge > 0) {
            ox = mul4s(ox, dx0); oy = mul4s(oy, dy0); oz = mul4(oz, dz0);
            pnew = mul4(mul4s(mul4s(pnew, dx0), dy0), dz0);
            _Pragma("unroll 1")
            for (int s = 1; s < P.n_sponge; s++) {
                const float dxs = P.decx[s][i];
                const float dys = row0 ? P.decy[s][j0] : 1.0f;
                const float4 dzs = lane_ok ? ld4(P.decz[s] + k0) : f4(1.0f);
                ox = mul4s(ox, dxs); oy = mul4s(oy, dys); oz = mul4(oz, dzs);
                pnew = mul4(mul4s(mul4s(pnew, dxs), dys), dzs);
            }
        }
        if (own) {
            st4(F.p_out + c, sel4(e0, e1, e2, e3, pnew, z4));
            st4(F.vx_out + c, sel4(e0, e1, e2, e3, ox, z4));
            st4(F.vy_out + c, sel4(e0, e1, e2, e3, oy, z4));
            st4(F.vz_out + c, sel4(e0, e1, e2, e3, oz, z4));
        }
        vxp = vxn;
        pc = pn;
    }
}

// MINB = blocks per SM the register allocation aims at: 2 = whatever the code needs (104-122 registers), 3 = at most 80
// (a few spilled words, half again as many warps in flight -- the kernel waits on memory most of the time)
template <bool UNI, bool FLAT, int MINB>
__global__ void __launch_bounds__(256, MINB) k1_step_march_ade(const __grid_constant__ StepParams P, const __grid_constant__ AdeFused A)
{
    const FieldSet F{P.p_in, P.vx_in, P.vy_in, P.vz_in, P.p_out, P.vx_out, P.vy_out, P.vz_out};
    k1_tile_ade<UNI, FLAT>(P, A, F, (int)blockIdx.x + P.bx_off, (int)blockIdx.y + P.by_off, (int)blockIdx.z + P.bz_off);
}

// Upper nibble of the mask bytes from the dense material-id array (with the slab's live ghost planes, as the geometry):
// M_ADE where the cell's material carries poles, M_?SAME where the + neighbour carries the same material; optionally
// the ids themselves in the padded field layout.  used[m] != 0 <=> material m has poles.
struct UsedIds { uint8_t used[256]; };
__global__ void k_ade_bits(const uint8_t *mat, uint8_t *mask, uint8_t *mat_pad, UsedIds U, int nx, int ny, int nz, int pitch,
                           long long plane, int has_lower, int has_upper)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int i = (int)blockIdx.z - 1;                            // -1 .. nx
    if (k >= nz) return;
    const long long c = (long long)i * plane + (long long)j * pitch + k;
    const bool live = !((i < 0 && !has_lower) || (i >= nx && !has_upper));
    auto M = [&](int ii, int jj, int kk) -> uint8_t { return mat[((long long)(ii + has_lower) * ny + jj) * nz + kk]; };
    uint8_t bits = 0, id = 0;
    if (live && mat) {                                            // mat == nullptr: clear the nibble
        id = M(i, j, k);
        if (U.used[id]) {
            bits = M_ADE;
            const bool have_xn = (i + 1 < nx) || (i + 1 == nx && has_upper);
            if (have_xn && M(i + 1, j, k) == id) bits |= M_XSAME;
            if (j + 1 < ny && M(i, j + 1, k) == id) bits |= M_YSAME;
            if (k + 1 < nz && M(i, j, k + 1) == id) bits |= M_ZSAME;
        } else id = 0;
    }
    mask[c] = (uint8_t)((mask[c] & 0x0F) | bits);
    if (mat_pad) mat_pad[c] = id;
}

// ------------------------------------------------------------------------------------------
// Energy (core/solver.py:2689-2706): sums of p^2 and v^2 over air cells, fp64 accumulation.
// ------------------------------------------------------------------------------------------
__global__ void k_energy(const float *p, const float *vx, const float *vy, const float *vz,
                         const uint8_t *mask, int nx, int ny, int nz, int pitch, long long plane,
                         double *out2)
{
    double sp = 0.0, sv = 0.0;
    const long long rows = (long long)nx * ny;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const long long c0 = (row / ny) * plane + (row % ny) * pitch;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) {
            const long long c = c0 + k;
            if (mask && !(mask[c] & M_AIR)) continue;
            const float pp = p[c] * p[c];
            const float vv = (vx[c] * vx[c] + vy[c] * vy[c]) + vz[c] * vz[c];
            sp += (double)pp; sv += (double)vv;
        }
    }
    __shared__ double s_p[32], s_v[32];
    for (int o = 16; o > 0; o >>= 1) { sp += __shfl_down_sync(0xffffffffu, sp, o); sv += __shfl_down_sync(0xffffffffu, sv, o); }
    if ((threadIdx.x & 31) == 0) { s_p[threadIdx.x >> 5] = sp; s_v[threadIdx.x >> 5] = sv; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        sp = threadIdx.x < nw ? s_p[threadIdx.x] : 0.0; sv = threadIdx.x < nw ? s_v[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) { sp += __shfl_down_sync(0xffffffffu, sp, o); sv += __shfl_down_sync(0xffffffffu, sv, o); }
        if (threadIdx.x == 0) { atomicAdd(out2, sp); atomicAdd(out2 + 1, sv); }
    }
}

}  // namespace sb
