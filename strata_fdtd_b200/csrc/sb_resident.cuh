// sb_resident.cuh -- K5: shared-memory-resident persistent step kernel (sm_100a).
//
// For grids small enough that the four fields fit in the aggregate shared memory of the GPU
// (148 SMs x 227 KB = 33 MB, about 1.3 M cells: BASELINE config 1 and every 100^3-class example
// script of the reference) a step is latency-bound, not bandwidth-bound: K1 streams the grid
// through L2 once per launch and pays a launch + ramp + tail per step.  K5 instead keeps the grid
// ON CHIP for a whole chunk of steps:
//
//   * the grid is cut into nbi x nbj boxes (full rows along k), one CTA per box, at most one CTA per
//     SM, all co-resident (cooperative launch);
//   * every CTA loads its box of {p,vx,vy,vz} into shared memory once, advances it n_steps times in
//     place (halo receive, barrier, velocity phase, barrier, pressure phase) and stores it back once -- HBM/L2 see
//     2 x 16 B per cell per CHUNK instead of 32 B per cell per STEP;
//   * per step only the four p faces of a box cross CTAs.  They go through an exchange area in L2 as
//     (value, step tag) pairs in 8-byte units, two per 16-byte volatile store -- the receiver polls the
//     data itself until the tag is the step it needs, so there is no flag, no fence and no grid-wide
//     barrier on the path (the "LL" protocol of NCCL, here between SMs of one GPU).  Slots are double
//     buffered by step parity; a box can only be one step ahead of a neighbour, which makes that enough.
//     The face velocities on the low sides (vx at i0-1, vy at j0-1) are kept redundantly in the box,
//     updated with the owner's exact operations (the same trick as the multi-GPU slabs, DESIGN.md
//     section 5), so nothing but p travels;
//   * optionally (R.split, off by default: measured slower, the second pass over the items costs more than the
//     exchange latency it hides) the part of the velocity phase that needs no halo value runs while the
//     neighbours' faces are in flight.
//
// Arithmetic is the contract of sb_kernels.cuh: separately rounded fp32 operations in the reference's
// order (fdtd_step.cpp:34-80, 109-211 / 235-439; boundaries.cpp:66-89; pml.cpp:47-149;
// core/solver.py:2386-2439).  The sponge multiply of the velocities is deferred to their next use (the
// pressure update reads them undamped, exactly as the reference applies the sponge after both updates),
// which performs the same operations on the same values.
//
// The phase functions are __host__ __device__ so that tests/emu/ can run the box decomposition in
// lockstep on the CPU against the oracle (index logic, halos, deferred damping); the product only ever
// calls them from k5_resident.
#pragma once
#include "sb_kernels.cuh"
#include <string.h>

#ifdef __CUDA_ARCH__
#define SB_LDG(ptr) __ldg(ptr)
#define SB_DMUL(a, b) __dmul_rn((a), (b))
#else
#define SB_LDG(ptr) (*(ptr))
#define SB_DMUL(a, b) ((a) * (b))
#endif

namespace sb {

#ifndef SB_K5_NT
#define SB_K5_NT 512
#endif
constexpr int K5_NT = SB_K5_NT;            // threads per CTA
constexpr int K5_MAX_PROBES = 1024;

struct ResParams {
    float *set[2][4];                      // plane-0 pointers of both ping-pong sets; set[cur] holds the state on entry
    int cur, n_steps;
    const uint8_t *mask;                   // plane-0 pointer of the face-mask bytes, or nullptr (all air)
    const float *cvx, *cvy, *cvz;          // per-face velocity coefficients
    const float *icx, *icy, *icz;          // per-cell inverse spacing, or nullptr (uniform grid)
    const float *decx[MAX_SPONGES], *decy[MAX_SPONGES], *decz[MAX_SPONGES];
    int n_sponge;
    float cp;
    int nx, ny, nz, pitch;
    long long plane;
    int nbi, nbj, LI, LJ, kp;              // box grid, largest box extents, shared-memory row pitch (floats, multiple of 4)
    int n_inline;                          // point sources (p field), list order
    int inl_i[SB_MAX_INLINE], inl_j[SB_MAX_INLINE], inl_k[SB_MAX_INLINE], inl_src[SB_MAX_INLINE];
    double inl_weight[SB_MAX_INLINE];
    const double *src_vals;                // [n_steps][n_sources]
    int n_sources;
    int n_probes, n_rec;
    const int *probe_ijk;                  // 3 ints per probe
    float *rec;                            // [n_steps][n_rec]
    uint4 *xch;                            // exchange area [2 parities][boxes][4 faces][xch_face] of (value, tag) pairs
    int xch_face;                          // uint4 per face slot = max(LI, LJ) * kp / 2
    unsigned tag_base;                     // p after step s of this launch travels with tag tag_base + s + 1 (never 0)
    int *err_flag;
    int split;                             // 1 = overlap the halo-free part of the velocity phase with the exchange
    int n_ops;                             // Mur / radiation planes, list order; with any, the faces are published and the
    PlaneOp ops[SB_MAX_PLANE_OPS];         // sources added in a pass of their own after the planes (res_after_planes)
};

// shared-memory layout, float offsets; p carries a one-cell halo on all four sides, vx / vy a low-side ghost;
// the x tables (they vary along the planes a thread loops over) and the face-mask words live there too
struct ResMap {
    int kp, LI, LJ, o_vx, o_vy, o_vz, o_xt, o_mask, o_end;
    SB_HD ResMap(int LI_, int LJ_, int kp_, bool geom) : kp(kp_), LI(LI_), LJ(LJ_)
    {
        o_vx = (LI + 2) * (LJ + 2) * kp;
        o_vy = o_vx + (LI + 1) * LJ * kp;
        o_vz = o_vy + LI * (LJ + 1) * kp;
        o_xt = o_vz + LI * LJ * kp;
        o_mask = o_xt + (3 * (LI + 1) + 3) / 4 * 4;
        o_end = o_mask + (geom ? (LI + 2) * (LJ + 2) * (kp / 4) : 0);
    }
    SB_HD explicit ResMap(const ResParams &R) : ResMap(R.LI, R.LJ, R.kp, R.mask != nullptr) {}
    SB_HD int p(int li, int lj) const { return ((li + 1) * (LJ + 2) + (lj + 1)) * kp; }     // li -1..LI, lj -1..LJ
    SB_HD int vx(int li, int lj) const { return o_vx + ((li + 1) * LJ + lj) * kp; }          // li -1..LI-1
    SB_HD int vy(int li, int lj) const { return o_vy + (li * (LJ + 1) + (lj + 1)) * kp; }    // lj -1..LJ-1
    SB_HD int vz(int li, int lj) const { return o_vz + (li * LJ + lj) * kp; }
    SB_HD int cvx(int li) const { return o_xt + li + 1; }                                    // li -1..LI-1
    SB_HD int icx(int li) const { return o_xt + (LI + 1) + li + 1; }
    SB_HD int dx0(int li) const { return o_xt + 2 * (LI + 1) + li + 1; }
    SB_HD int mw(int p_off) const { return o_mask + (p_off >> 2); }                          // mask word of the 4 cells at p_off
};

static inline long long res_smem_bytes(int LI, int LJ, int kp, int n_probes, bool geom)
{
    const ResMap M(LI, LJ, kp, geom);
    return (long long)M.o_end * 4 + (long long)(2 * n_probes + 4) * 4;
}

// Box grid for a given SM count and shared-memory limit: fewest items per thread, then least shared memory.
// Returns false when the grid does not fit on chip.
// min_li / min_lj: smallest box extent allowed along i / j (2 where a Mur / radiation plane lies on that axis: a face cell
// and its interior neighbour must share a box).
static inline bool res_choose_partition(int nx, int ny, int nz, int n_sm, long long smem_limit, int n_probes, bool geom,
                                        int *nbi_out, int *nbj_out, int min_li = 1, int min_lj = 1)
{
    const int kp = (nz + 3) / 4 * 4, K4 = kp / 4;
    long long best_cost = -1;
    for (int nbi = 1; nbi <= nx && nbi <= n_sm; nbi++) {
        int nbj = ny < n_sm / nbi ? ny : n_sm / nbi;
        if (nbj < 1) break;
        if (nx / nbi < min_li) break;
        if (ny / nbj < min_lj) nbj = ny / min_lj;
        if (nbj < 1) continue;
        const int LI = (nx + nbi - 1) / nbi, LJ = (ny + nbj - 1) / nbj;
        if ((long long)LJ * K4 > K5_NT) continue;                     // one thread per (row, float4) column at least
        if ((long long)(LI + 2) * (LJ + 2) * kp * 2 >= (1LL << 30)) continue;
        const long long bytes = res_smem_bytes(LI, LJ, kp, n_probes, geom);
        if (bytes > smem_limit) continue;
        // work of the busiest thread: its share of the planes plus the ghost-face items handed to the idle ones
        const int ncol = LJ * K4, G = K5_NT / ncol;
        const long long items = (long long)LI * ncol + ncol + (long long)LI * K4;
        const long long iters = (items + K5_NT - 1) / K5_NT > (LI + G - 1) / G ? (items + K5_NT - 1) / K5_NT : (LI + G - 1) / G;
        const long long cost = iters * (1LL << 32) + bytes;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; *nbi_out = nbi; *nbj_out = nbj; }
    }
    return best_cost >= 0;
}

struct ResBlock { int bi, bj, i0, j0, li_n, lj_n; };

SB_HD ResBlock res_block(const ResParams &R, int b)
{
    ResBlock B;
    B.bi = b / R.nbj; B.bj = b - B.bi * R.nbj;
    B.i0 = (int)((long long)B.bi * R.nx / R.nbi);
    B.li_n = (int)((long long)(B.bi + 1) * R.nx / R.nbi) - B.i0;
    B.j0 = (int)((long long)B.bj * R.ny / R.nbj);
    B.lj_n = (int)((long long)(B.bj + 1) * R.ny / R.nbj) - B.j0;
    return B;
}

// Per-thread constants: a thread owns one (row, float4) column of the box and every G-th plane of it.
struct ResThread {
    int tid, g, G, lj, gj, k0;
    int op, ox, oy, oz;                    // shared-memory offsets of the thread's first item (plane li = g)
    bool active;
    bool e0, e1, e2, e3, u0, u1, u2, u3;   // element inside the grid / its z face is updated
    bool upd_y, sub_y, up_i, low_i;        // y face updated; j > 0; a box above / below exists
    bool pub_jlo, pub_jhi, pub_j;          // the column lies on a j face that a neighbour needs
    float cvy, icy, dy0;
    float4 cvz4, icz4, dz0;
    unsigned inl_mask;
};

SB_HD ResThread res_thread(const ResParams &R, const ResBlock &B, int tid)
{
    ResThread T;
    const ResMap M(R);
    const int K4 = R.kp >> 2, ncol = B.lj_n * K4;
    T.tid = tid;
    T.G = K5_NT / ncol;
    T.g = tid / ncol;
    const int c = tid - T.g * ncol;
    T.active = T.g < T.G;
    T.lj = c / K4; T.k0 = 4 * (c - T.lj * K4); T.gj = B.j0 + T.lj;
    T.op = M.p(T.g, T.lj) + T.k0; T.ox = M.vx(T.g, T.lj) + T.k0; T.oy = M.vy(T.g, T.lj) + T.k0; T.oz = M.vz(T.g, T.lj) + T.k0;
    T.e0 = T.k0 < R.nz; T.e1 = T.k0 + 1 < R.nz; T.e2 = T.k0 + 2 < R.nz; T.e3 = T.k0 + 3 < R.nz;
    T.u0 = T.k0 < R.nz - 1; T.u1 = T.k0 + 1 < R.nz - 1; T.u2 = T.k0 + 2 < R.nz - 1; T.u3 = T.k0 + 3 < R.nz - 1;
    T.upd_y = T.gj < R.ny - 1; T.sub_y = T.gj > 0;
    T.up_i = B.i0 + B.li_n < R.nx; T.low_i = B.i0 > 0;
    T.pub_jlo = T.lj == 0 && T.gj > 0; T.pub_jhi = T.lj == B.lj_n - 1 && T.gj < R.ny - 1;
    T.pub_j = T.pub_jlo || T.pub_jhi;
    T.cvy = 0.0f; T.icy = 1.0f; T.dy0 = 1.0f;
    T.cvz4 = f4(0.0f); T.icz4 = f4(1.0f); T.dz0 = f4(1.0f);
    T.inl_mask = 0;
    if (!T.active) return T;
    if (T.upd_y) T.cvy = SB_LDG(R.cvy + T.gj);
    T.cvz4 = ld4(R.cvz + T.k0);
    if (R.icx) { T.icy = SB_LDG(R.icy + T.gj); T.icz4 = ld4(R.icz + T.k0); }
    if (R.n_sponge > 0) { T.dy0 = SB_LDG(R.decy[0] + T.gj); T.dz0 = ld4(R.decz[0] + T.k0); }
    for (int q = 0; q < R.n_inline; q++)
        if (R.inl_j[q] == T.gj && R.inl_k[q] >= T.k0 && R.inl_k[q] < T.k0 + 4 &&
            R.inl_i[q] >= B.i0 && R.inl_i[q] < B.i0 + B.li_n) T.inl_mask |= 1u << q;
    return T;
}

// ---- box <-> global --------------------------------------------------------------------------------
SB_HD void res_load(const ResParams &R, const ResBlock &B, float *sm, int tid)
{
    const ResMap M(R);
    const int K4 = R.kp >> 2;
    const float4 z4 = f4(0.0f);
    float *const *F = R.set[R.cur];
    {   // p with its halo (and the face-mask words of the same cells)
        const int nj = B.lj_n + 2, n = (B.li_n + 2) * nj * K4;
        unsigned *smw = reinterpret_cast<unsigned *>(sm);
        for (int idx = tid; idx < n; idx += K5_NT) {
            const int k4 = idx % K4, r = idx / K4, lj = r % nj - 1, li = r / nj - 1;
            const int gi = B.i0 + li, gj = B.j0 + lj;
            const bool in = gi >= 0 && gi < R.nx && gj >= 0 && gj < R.ny;
            const long long c = (long long)gi * R.plane + (long long)gj * R.pitch + 4 * k4;
            st4(sm + M.p(li, lj) + 4 * k4, in ? ld4(F[0] + c) : z4);
            if (R.mask) smw[M.mw(M.p(li, lj) + 4 * k4)] = in ? *reinterpret_cast<const unsigned *>(R.mask + c) : 0u;
        }
    }
    {   // vx with the plane below
        const int n = (B.li_n + 1) * B.lj_n * K4;
        for (int idx = tid; idx < n; idx += K5_NT) {
            const int k4 = idx % K4, r = idx / K4, lj = r % B.lj_n, li = r / B.lj_n - 1;
            const int gi = B.i0 + li, gj = B.j0 + lj;
            st4(sm + M.vx(li, lj) + 4 * k4, gi >= 0 ? ld4(F[1] + (long long)gi * R.plane + (long long)gj * R.pitch + 4 * k4) : z4);
        }
    }
    {   // vy with the row below
        const int nj = B.lj_n + 1, n = B.li_n * nj * K4;
        for (int idx = tid; idx < n; idx += K5_NT) {
            const int k4 = idx % K4, r = idx / K4, lj = r % nj - 1, li = r / nj;
            const int gi = B.i0 + li, gj = B.j0 + lj;
            st4(sm + M.vy(li, lj) + 4 * k4, gj >= 0 ? ld4(F[2] + (long long)gi * R.plane + (long long)gj * R.pitch + 4 * k4) : z4);
        }
    }
    {
        const int n = B.li_n * B.lj_n * K4;
        for (int idx = tid; idx < n; idx += K5_NT) {
            const int k4 = idx % K4, r = idx / K4, lj = r % B.lj_n, li = r / B.lj_n;
            st4(sm + M.vz(li, lj) + 4 * k4, ld4(F[3] + (long long)(B.i0 + li) * R.plane + (long long)(B.j0 + lj) * R.pitch + 4 * k4));
        }
    }
    for (int li = tid - 1; li < B.li_n; li += K5_NT) {                  // x tables of planes i0-1 .. i0+li_n-1
        const int gi = B.i0 + li;
        sm[M.cvx(li)] = gi >= 0 ? R.cvx[gi] : 0.0f;
        sm[M.icx(li)] = (R.icx && gi >= 0) ? R.icx[gi] : 1.0f;
        sm[M.dx0(li)] = (R.n_sponge > 0 && gi >= 0) ? R.decx[0][gi] : 1.0f;
    }
}

// final state -> set[(cur + n_steps) & 1]; the sponge multiply of the last step's velocities happens here
SB_HD void res_store(const ResParams &R, const ResBlock &B, const float *sm, int tid)
{
    const ResMap M(R);
    const int K4 = R.kp >> 2, n = B.li_n * B.lj_n * K4;
    float *const *F = R.set[(R.cur + R.n_steps) & 1];
    const float4 z4 = f4(0.0f);
    for (int idx = tid; idx < n; idx += K5_NT) {
        const int k4 = idx % K4, r = idx / K4, lj = r % B.lj_n, li = r / B.lj_n, k0 = 4 * k4;
        const int gi = B.i0 + li, gj = B.j0 + lj;
        const bool e0 = k0 < R.nz, e1 = k0 + 1 < R.nz, e2 = k0 + 2 < R.nz, e3 = k0 + 3 < R.nz;
        float4 vx = ld4(sm + M.vx(li, lj) + k0), vy = ld4(sm + M.vy(li, lj) + k0), vz = ld4(sm + M.vz(li, lj) + k0);
        if (R.n_steps > 0)
            for (int q = 0; q < R.n_sponge; q++) {
                vx = mul4s(vx, SB_LDG(R.decx[q] + gi)); vy = mul4s(vy, SB_LDG(R.decy[q] + gj)); vz = mul4(vz, ld4(R.decz[q] + k0));
            }
        const long long c = (long long)gi * R.plane + (long long)gj * R.pitch + k0;
        st4(F[0] + c, sel4(e0, e1, e2, e3, ld4(sm + M.p(li, lj) + k0), z4));
        st4(F[1] + c, sel4(e0, e1, e2, e3, vx, z4));
        st4(F[2] + c, sel4(e0, e1, e2, e3, vy, z4));
        st4(F[3] + c, sel4(e0, e1, e2, e3, vz, z4));
    }
}

// ---- exchange of p faces between boxes ------------------------------------------------------------
// face 0: plane li = 0, 1: plane li = li_n-1, 2: row lj = 0, 3: row lj = lj_n-1 of the publishing box
SB_HD long long res_xch_slot(const ResParams &R, int parity, int box, int face)
{
    return (((long long)parity * (R.nbi * R.nbj) + box) * 4 + face) * R.xch_face;
}

SB_HD unsigned res_bits(float f)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    unsigned u; memcpy(&u, &f, 4); return u;
#endif
}

SB_HD void res_publish(uint4 *dst, float4 v, unsigned tag)
{
    const uint4 a = make_uint4(res_bits(v.x), tag, res_bits(v.y), tag);
    const uint4 b = make_uint4(res_bits(v.z), tag, res_bits(v.w), tag);
#ifdef __CUDA_ARCH__
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 1), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
#else
    dst[0] = a; dst[1] = b;
#endif
}

// Halo items of a box, flattened over its (up to four) faces so that they spread evenly over the threads:
// item idx -> the 32-byte source in the publisher's slot (parity `par`) and the shared-memory offset it lands at.
struct ResHalo {
    int n0, n1, n2, n3, total;
    SB_HD ResHalo(const ResParams &R, const ResBlock &B)
    {
        const int K4 = R.kp >> 2;
        n0 = B.bi > 0 ? B.lj_n * K4 : 0;                               // plane i0-1      = face 1 of the box below
        n1 = B.bi < R.nbi - 1 ? B.lj_n * K4 : 0;                       // plane i0+li_n   = face 0 of the box above
        n2 = B.bj > 0 ? B.li_n * K4 : 0;                               // row j0-1        = face 3 of the box to the left
        n3 = B.bj < R.nbj - 1 ? B.li_n * K4 : 0;                       // row j0+lj_n     = face 2 of the box to the right
        total = n0 + n1 + n2 + n3;
    }
    SB_HD void item(const ResParams &R, const ResBlock &B, const ResMap &M, int par, int idx, const uint4 *&src, int &dst) const
    {
        const int K4 = R.kp >> 2, K2 = R.kp >> 1;
        int f = 0;
        if (idx >= n0) { idx -= n0; f = 1;
            if (idx >= n1) { idx -= n1; f = 2;
                if (idx >= n2) { idx -= n2; f = 3; } } }
        const int row = idx / K4, k0 = 4 * (idx - row * K4);
        const int box = f == 0 ? (B.bi - 1) * R.nbj + B.bj : f == 1 ? (B.bi + 1) * R.nbj + B.bj
                      : f == 2 ? B.bi * R.nbj + B.bj - 1 : B.bi * R.nbj + B.bj + 1;
        const int their_face = f == 0 ? 1 : f == 1 ? 0 : f == 2 ? 3 : 2;
        src = R.xch + res_xch_slot(R, par, box, their_face) + row * K2 + (k0 >> 1);
        dst = (f == 0 ? M.p(-1, row) : f == 1 ? M.p(B.li_n, row) : f == 2 ? M.p(row, -1) : M.p(row, B.lj_n)) + k0;
    }
};

// Compile-time variants: GEOM face masks; UNI uniform grid (no inverse-cell multiplies); NS sponge layers: 0, 1, or
// 2 = any number (layers beyond the first read their tables from global memory)
// ---- velocity phase: v += cv * grad p on the box and on its two low-side ghost faces ---------------
// pass 0: the items that read no halo value; pass 1: the others and the ghost faces; pass 2: everything
template <bool GEOM, bool UNI, int NS>
SB_HD void res_phase_v(const ResParams &R, const ResBlock &B, const ResThread &T, float *sm, int s, int pass)
{
    const ResMap M(R);
    const int sti = (R.LJ + 2) * R.kp;                                 // plane stride of the p array
    const bool damp = s > 0;                                           // sponge of the previous step, applied on first use
    const unsigned *smw = reinterpret_cast<const unsigned *>(sm);
    if (T.active) {
        const int dp = T.G * sti, dvx = T.G * R.LJ * R.kp, dvy = T.G * (R.LJ + 1) * R.kp;
        int op = T.op, ox = T.ox, oy = T.oy, oz = T.oz;
        for (int li = T.g; li < B.li_n; li += T.G, op += dp, ox += dvx, oy += dvy, oz += dvx) {
            const bool upd_x = li < B.li_n - 1 || T.up_i;
            const bool halo = (upd_x && li == B.li_n - 1) || T.pub_jhi;
            if (pass != 2 && halo != (pass == 1)) continue;
            const float *pp = sm + op;
            const float4 p = ld4(pp);
            float4 vx = ld4(sm + ox), vy = ld4(sm + oy), vz = ld4(sm + oz);
            if (damp && NS > 0) {                                      // pml.cpp:47-98, sponge of the previous step
                vx = mul4s(vx, sm[M.dx0(li)]); vy = mul4s(vy, T.dy0); vz = mul4(vz, T.dz0);
                if (NS == 2)
                    for (int q = 1; q < R.n_sponge; q++) {
                        vx = mul4s(vx, SB_LDG(R.decx[q] + B.i0 + li)); vy = mul4s(vy, SB_LDG(R.decy[q] + T.gj));
                        vz = mul4(vz, ld4(R.decz[q] + T.k0));
                    }
            }
            unsigned mk = ALL_OPEN;
            if (GEOM) mk = smw[M.mw(op)];
            if (upd_x) {                                               // fdtd_step.cpp:34-47 / 255-269
                vx = add4(vx, mul4s(sub4(ld4(pp + sti), p), sm[M.cvx(li)]));
                if (GEOM) vx = keep4(vx, mk, M_XOPEN);                 // boundaries.cpp:66-89
            }
            if (T.upd_y) {                                             // fdtd_step.cpp:53-64 / 272-288
                vy = add4(vy, mul4s(sub4(ld4(pp + R.kp), p), T.cvy));
                if (GEOM) vy = keep4(vy, mk, M_YOPEN);
            }
            {                                                          // fdtd_step.cpp:71-81 / 291-306
                // branch-free: the float4 that holds the last face of a row (or its padding) sits in every warp, so a
                // branch on it made every warp run both versions; faces that are not updated keep their value
                const float p_next = T.u3 ? pp[4] : 0.0f;
                const float4 upd = add4(vz, mul4(sub4(make_float4(p.y, p.z, p.w, p_next), p), T.cvz4));
                vz = sel4(T.u0, T.u1, T.u2, T.u3, upd, vz);
                if (GEOM) {                                            // only updated faces are zeroed
                    const unsigned m = mk | (T.u0 ? 0u : 0x08u) | (T.u1 ? 0u : 0x0800u) | (T.u2 ? 0u : 0x080000u) | (T.u3 ? 0u : 0x08000000u);
                    vz = keep4(vz, m, M_ZOPEN);
                }
            }
            st4(sm + ox, vx); st4(sm + oy, vy); st4(sm + oz, vz);
        }
    }
    if (pass == 0) return;
    // The two ghost faces go to the threads with the least work of their own: vx of plane i0-1 to the last group
    // of columns, vy of row j0-1 to the threads counted from the top (the idle ones first).
    if (T.low_i && T.active && T.g == T.G - 1) {                       // redundant with the owner of plane i0-1
        const int gi = B.i0 - 1;
        float *q = sm + M.vx(-1, T.lj) + T.k0;
        float4 v = ld4(q);
        if (damp && NS > 0) {
            v = mul4s(v, sm[M.dx0(-1)]);
            if (NS == 2) for (int w = 1; w < R.n_sponge; w++) v = mul4s(v, SB_LDG(R.decx[w] + gi));
        }
        const int o0 = M.p(0, T.lj) + T.k0, om = M.p(-1, T.lj) + T.k0;
        v = add4(v, mul4s(sub4(ld4(sm + o0), ld4(sm + om)), sm[M.cvx(-1)]));
        if (GEOM) v = keep4(v, smw[M.mw(om)], M_XOPEN);
        st4(q, v);
    }
    if (B.j0 > 0) {                                                    // redundant with the owner of row j0-1
        const int K4 = R.kp >> 2, gj = B.j0 - 1, n = B.li_n * K4;
        const float cy = SB_LDG(R.cvy + gj);
        for (int idx = K5_NT - 1 - T.tid; idx < n; idx += K5_NT) {
            const int li = idx / K4, k0 = 4 * (idx - li * K4);
            float *q = sm + M.vy(li, -1) + k0;
            float4 v = ld4(q);
            if (damp && NS > 0) for (int w = 0; w < R.n_sponge; w++) v = mul4s(v, SB_LDG(R.decy[w] + gj));
            const int o0 = M.p(li, 0) + k0, om = M.p(li, -1) + k0;
            v = add4(v, mul4s(sub4(ld4(sm + o0), ld4(sm + om)), cy));
            if (GEOM) v = keep4(v, smw[M.mw(om)], M_YOPEN);
            st4(q, v);
        }
    }
}

// ---- pressure phase: p += cp * div v, solids, sponge, point sources; publishes the box faces ----------
// OPS: Mur / radiation planes are registered -- sources and face publishing are left to res_after_planes
template <bool GEOM, bool UNI, int NS, bool OPS = false>
SB_HD void res_phase_p(const ResParams &R, const ResBlock &B, const ResThread &T, float *sm, int s)
{
    if (!T.active) return;
    const ResMap M(R);
    const int K2 = R.kp >> 1;
    const unsigned tag = R.tag_base + (unsigned)s + 1u;                // p after step s
    uint4 *xo = R.xch + res_xch_slot(R, (s + 1) & 1, B.bi * R.nbj + B.bj, 0) + (T.k0 >> 1);
    const unsigned *smw = reinterpret_cast<const unsigned *>(sm);
    const int sti = (R.LJ + 2) * R.kp, svx = R.LJ * R.kp;
    const int dp = T.G * sti, dvx = T.G * svx, dvy = T.G * (R.LJ + 1) * R.kp;
    // Faces first: a thread visits its lowest plane, then its highest, then the ones in between, so that the two x faces
    // of the box are on their way to the neighbours while the interior is still being computed -- by the time the
    // neighbours ask for them (right after their own pressure phase) they have long arrived.  Cells are independent
    // within this phase, so the order changes nothing in the results.
    const int n_it = T.g < B.li_n ? (B.li_n - T.g + T.G - 1) / T.G : 0;
    for (int it = 0; it < n_it; it++) {
        const int m = it == 0 ? 0 : (it == 1 ? n_it - 1 : it - 1);
        const int li = T.g + m * T.G;
        const int op = T.op + m * dp, ox = T.ox + m * dvx, oy = T.oy + m * dvy, oz = T.oz + m * dvx;
        const float4 vx = ld4(sm + ox), vy = ld4(sm + oy), vz = ld4(sm + oz);
        float4 ddx = vx, ddy = vy;                                     // fdtd_step.cpp:109-211: zero ghost at index 0
        if (li > 0 || T.low_i) ddx = sub4(vx, ld4(sm + ox - svx));
        if (T.sub_y) ddy = sub4(vy, ld4(sm + oy - R.kp));
        const float vz_prev = T.k0 > 0 ? sm[oz - 1] : 0.0f;
        float4 ddz = sub4(vz, make_float4(vz_prev, vz.x, vz.y, vz.z));
        if (!UNI) { ddx = mul4s(ddx, sm[M.icx(li)]); ddy = mul4s(ddy, T.icy); ddz = mul4(ddz, T.icz4); }
        float4 pn = add4(ld4(sm + op), mul4s(add4(add4(ddx, ddy), ddz), R.cp));
        if (GEOM) pn = keep4(pn, smw[M.mw(op)], M_AIR);
        if (NS > 0) {                                                  // pml.cpp:100-149
            pn = mul4(mul4s(mul4s(pn, sm[M.dx0(li)]), T.dy0), T.dz0);
            if (NS == 2)
                for (int q = 1; q < R.n_sponge; q++)
                    pn = mul4(mul4s(mul4s(pn, SB_LDG(R.decx[q] + B.i0 + li)), SB_LDG(R.decy[q] + T.gj)), ld4(R.decz[q] + T.k0));
        }
        // (row padding is not masked here: no valid cell ever reads a padded element, and res_store writes zeros there)
        if (!OPS && T.inl_mask) {                                      // float64 add, fp32 store (solver.py:2421), list order
            for (int q = 0; q < R.n_inline; q++)
                if (((T.inl_mask >> q) & 1u) && R.inl_i[q] == B.i0 + li) {
                    const double w = SB_DMUL(R.src_vals[(long long)s * R.n_sources + R.inl_src[q]], R.inl_weight[q]);
                    const int e = R.inl_k[q] - T.k0;
                    if (e == 0) pn.x = (float)((double)pn.x + w);
                    else if (e == 1) pn.y = (float)((double)pn.y + w);
                    else if (e == 2) pn.z = (float)((double)pn.z + w);
                    else pn.w = (float)((double)pn.w + w);
                }
        }
        st4(sm + op, pn);
        const bool f_lo = li == 0 && T.low_i, f_hi = li == B.li_n - 1 && T.up_i;
        if (!OPS && (f_lo || f_hi || T.pub_j)) {                       // a face some neighbour needs
            if (f_lo) res_publish(xo + T.lj * K2, pn, tag);
            if (f_hi) res_publish(xo + R.xch_face + T.lj * K2, pn, tag);
            if (T.pub_jlo) res_publish(xo + 2 * R.xch_face + li * K2, pn, tag);
            if (T.pub_jhi) res_publish(xo + 3 * R.xch_face + li * K2, pn, tag);
        }
    }
}

// calls f.template run<GEOM, UNI, NS>() for the variant a configuration needs
template <typename F> static inline int res_dispatch(bool geom, bool uni, int n_sponge, F &f)
{
    const int ns = n_sponge > 2 ? 2 : n_sponge;
    if (geom) {
        if (uni) return ns == 0 ? f.template run<true, true, 0>() : ns == 1 ? f.template run<true, true, 1>() : f.template run<true, true, 2>();
        return ns == 0 ? f.template run<true, false, 0>() : ns == 1 ? f.template run<true, false, 1>() : f.template run<true, false, 2>();
    }
    if (uni) return ns == 0 ? f.template run<false, true, 0>() : ns == 1 ? f.template run<false, true, 1>() : f.template run<false, true, 2>();
    return ns == 0 ? f.template run<false, false, 0>() : ns == 1 ? f.template run<false, false, 1>() : f.template run<false, false, 2>();
}

#if defined(__CUDACC__) && !defined(SB_RESIDENT_NO_KERNEL)
// ---- Mur / radiation planes, then point sources, then the faces (core/solver.py:2572-2584 order) ----------------------
// Runs after the pressure phase when plane updates are registered.  A plane update reads and writes a face cell and its
// interior neighbour, both inside one box (the host only picks box grids that keep them together) and in shared memory;
// planes on different axes meet on the edges of the grid, so the list is walked in order with a block barrier wherever the
// axis changes -- the sequential order of boundaries/_boundaries.py:476-513, 700-760.  `prev` stays in global
// memory: every entry is read and written by one thread of one box, step after step.
__device__ __noinline__ void res_after_planes(const ResParams &R, const ResBlock &B, const ResMap &M, float *sm, int tid, int s)
{
    const int n[3] = {R.nx, R.ny, R.nz}, lo[3] = {B.i0, B.j0, 0}, ext[3] = {B.li_n, B.lj_n, R.nz};
    const int stride[3] = {(R.LJ + 2) * R.kp, R.kp, 1};                  // of p in shared memory along i, j, k
    const int base = M.p(0, 0);
    __syncthreads();                                                   // p of this step complete in the box
    int last_ax = -1;
    for (int o = 0; o < R.n_ops; o++) {
        const PlaneOp &op = R.ops[o];
        const int ax = op.axis, a_ax = ax == 0 ? 1 : 0, b_ax = ax == 2 ? 1 : 2;
        const int face = op.side ? n[ax] - 1 : 0, inner = op.side ? n[ax] - 2 : 1;
        if (face < lo[ax] || face >= lo[ax] + ext[ax]) continue;       // (uniform over the block)
        // consecutive planes on one axis need no barrier: a thread meets the same (a, b) column in each of them, so the two
        // sides (disjoint cells) and repeated updates of one face are ordered by the thread's own program order
        if (last_ax >= 0 && last_ax != ax) __syncthreads();
        last_ax = ax;
        const int eb = ext[b_ax], cells = ext[a_ax] * eb;
        for (int idx = tid; idx < cells; idx += K5_NT) {
            const int la = idx / eb, lb = idx - la * eb;
            const int off = base + la * stride[a_ax] + lb * stride[b_ax];
            const int ob = off + (face - lo[ax]) * stride[ax], oi = off + (inner - lo[ax]) * stride[ax];
            float *prev = op.prev + (long long)(lo[a_ax] + la) * n[b_ax] + (lo[b_ax] + lb);
            const float pi = sm[oi];
            sm[ob] = plane_op_value(op, *prev, sm[ob], pi);
            *prev = pi;
        }
    }
    if (last_ax >= 0 && R.n_inline) __syncthreads();
    if (tid == 0)                                                      // float64 add, fp32 store (solver.py:2421), list order
        for (int q = 0; q < R.n_inline; q++) {
            const int li = R.inl_i[q] - B.i0, lj = R.inl_j[q] - B.j0;
            if (li < 0 || li >= B.li_n || lj < 0 || lj >= B.lj_n) continue;
            float *cell = sm + M.p(li, lj) + R.inl_k[q];
            *cell = (float)((double)*cell + __dmul_rn(R.src_vals[(long long)s * R.n_sources + R.inl_src[q]], R.inl_weight[q]));
        }
    __syncthreads();
    // the faces the neighbouring boxes need, as res_phase_p publishes them when there are no planes
    const int K4 = R.kp >> 2, K2 = R.kp >> 1;
    const unsigned tag = R.tag_base + (unsigned)s + 1u;
    uint4 *xo = R.xch + res_xch_slot(R, (s + 1) & 1, B.bi * R.nbj + B.bj, 0);
    const int ni = B.lj_n * K4, nj = B.li_n * K4;
    const int n0 = B.i0 > 0 ? ni : 0, n1 = B.i0 + B.li_n < R.nx ? ni : 0, n2 = B.j0 > 0 ? nj : 0, n3 = B.j0 + B.lj_n < R.ny ? nj : 0;
    for (int idx = tid; idx < n0 + n1 + n2 + n3; idx += K5_NT) {
        int f = 0, q = idx;
        if (q >= n0) { q -= n0; f = 1;
            if (q >= n1) { q -= n1; f = 2;
                if (q >= n2) { q -= n2; f = 3; } } }
        const int row = q / K4, k0 = 4 * (q - row * K4);
        const int li = f == 0 ? 0 : f == 1 ? B.li_n - 1 : row, lj = f < 2 ? row : (f == 2 ? 0 : B.lj_n - 1);
        res_publish(xo + (long long)f * R.xch_face + row * K2 + (k0 >> 1), ld4(sm + M.p(li, lj) + k0), tag);
    }
}

// Receiver: the (value, tag) units of p after step s-1 that the neighbours published -> halo of the box.  A thread
// issues the loads of all its items at once and only then looks at the tags, so a face that has already arrived
// costs one L2 round trip, not one per item; units whose tag is not yet the wanted step are simply read again.
struct RecvPoll {
    static constexpr int Q = 4;              // halo items a thread keeps the addresses of (more are looked up per step)
    int *err_flag;
    bool dead;                               // a wait timed out: stop waiting, finish the chunk, report
    const uint4 *src0[Q];                    // the item's 32-byte source in the publisher's parity-0 slot (nullptr = no item)
    int dst[Q];                              // where it lands in this box's shared memory
    long long par_stride;                    // uint4 distance between the two parities of the exchange area
    static __device__ __forceinline__ uint4 ldv(const uint4 *src)
    {
        uint4 v;
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
        return v;
    }
    // which items a thread receives never changes during a launch: resolve (face, row, column) -> addresses once,
    // outside the step loop (the lookup is full of integer divisions)
    __device__ __forceinline__ void prepare(const ResParams &R, const ResBlock &B, const ResMap &M, const ResHalo &H, int tid)
    {
        par_stride = res_xch_slot(R, 1, 0, 0) - res_xch_slot(R, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int idx = tid + q * K5_NT;
            src0[q] = nullptr; dst[q] = 0;
            if (idx < H.total) H.item(R, B, M, 0, idx, src0[q], dst[q]);
        }
    }
    __device__ __forceinline__ void take(float *sm, const uint4 *src, int d, unsigned tag, uint4 a, uint4 b)
    {
        long long t0 = 0;
        while (!(a.y == tag && a.w == tag && b.y == tag && b.w == tag) && !dead) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > (2LL << 30)) { atomicExch(err_flag, 2); dead = true; }   // ~1 s: never hang the GPU
            a = ldv(src); b = ldv(src + 1);
        }
        st4(sm + d, make_float4(__uint_as_float(a.x), __uint_as_float(a.z), __uint_as_float(b.x), __uint_as_float(b.z)));
    }
    __device__ __forceinline__ void run(const ResParams &R, const ResBlock &B, const ResMap &M, const ResHalo &H,
                                        float *sm, int tid, int s)
    {
        const unsigned tag = R.tag_base + (unsigned)s;
        const long long po = (s & 1) ? par_stride : 0;
        uint4 a[Q], b[Q];
#pragma unroll
        for (int q = 0; q < Q; q++)                                    // all loads first, then the tags: one L2 round trip
            if (src0[q]) { a[q] = ldv(src0[q] + po); b[q] = ldv(src0[q] + po + 1); }
#pragma unroll
        for (int q = 0; q < Q; q++)
            if (src0[q]) take(sm, src0[q] + po, dst[q], tag, a[q], b[q]);
        for (int idx = tid + Q * K5_NT; idx < H.total; idx += K5_NT) { // boxes with very long faces
            const uint4 *src; int d;
            H.item(R, B, M, s & 1, idx, src, d);
            take(sm, src, d, tag, ldv(src), ldv(src + 1));
        }
    }
};

template <bool GEOM, bool UNI, int NS, bool OPS>
__global__ void __launch_bounds__(K5_NT, 1) k5_resident(const __grid_constant__ ResParams R)
{
    extern __shared__ float4 k5_smem4[];
    float *sm = reinterpret_cast<float *>(k5_smem4);
    const int tid = threadIdx.x, b = blockIdx.x;
    const ResBlock B = res_block(R, b);
    const ResThread T = res_thread(R, B, tid);
    const ResMap M(R);
    int *spr = reinterpret_cast<int *>(sm + M.o_end);                  // [0] = probes owned, then (slot, offset) pairs
    if (tid == 0) spr[0] = 0;
    res_load(R, B, sm, tid);
    __syncthreads();
    for (int t = tid; t < R.n_probes; t += K5_NT) {
        const int i = R.probe_ijk[3 * t] - B.i0, j = R.probe_ijk[3 * t + 1] - B.j0, k = R.probe_ijk[3 * t + 2];
        if (i >= 0 && i < B.li_n && j >= 0 && j < B.lj_n) {
            const int q = atomicAdd(&spr[0], 1);
            spr[1 + 2 * q] = t; spr[2 + 2 * q] = M.p(i, j) + k;
        }
    }
    __syncthreads();
    const int n_own = spr[0];
    RecvPoll recv;
    recv.err_flag = R.err_flag; recv.dead = false;
    const ResHalo H(R, B);
    recv.prepare(R, B, M, H, tid);

    for (int s = 0; s < R.n_steps; s++) {
        if (s > 0) {
            if (R.split) {                                             // halo-free velocities while the faces are in flight
                __syncthreads();
                res_phase_v<GEOM, UNI, NS>(R, B, T, sm, s, 0);
            }
            recv.run(R, B, M, H, sm, tid, s);
        }
        __syncthreads();                                               // p of step s-1 complete in the box, halo in place
        if (s > 0)
            for (int q = tid; q < n_own; q += K5_NT)                   // core/solver.py:2435-2439
                R.rec[(long long)(s - 1) * R.n_rec + spr[1 + 2 * q]] = sm[spr[2 + 2 * q]];
        res_phase_v<GEOM, UNI, NS>(R, B, T, sm, s, (s > 0 && R.split) ? 1 : 2);
        __syncthreads();
        res_phase_p<GEOM, UNI, NS, OPS>(R, B, T, sm, s);               // publishes the box faces as it goes
        if (OPS) res_after_planes(R, B, M, sm, tid, s);                // ... unless plane updates have to come first
    }
    __syncthreads();
    if (R.n_steps > 0)
        for (int q = tid; q < n_own; q += K5_NT)
            R.rec[(long long)(R.n_steps - 1) * R.n_rec + spr[1 + 2 * q]] = sm[spr[2 + 2 * q]];
    res_store(R, B, sm, tid);
}
#endif  // __CUDACC__

}  // namespace sb
