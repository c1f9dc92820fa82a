// sb_pipeline.cuh -- K6: the fused step kernel (K1), pipelined across time steps in one persistent launch.
//
// Launched once per step, K1 pays a ramp and a tail every step: on mid-sized grids (a few million to a few
// tens of millions of cells -- BASELINE config 2, 200^3) the tiles of one step are only ~2 waves of resident
// CTAs, so for a sizeable part of every step half the SMs are idle waiting for the last tiles.  K6 runs the
// very same tile code (k1_tile) from persistent CTAs that draw (step, chunk, tile) work items from a global
// ticket counter in step-major order:
//
//   * a tile of step t, chunk c reads planes of chunks c-1 .. c+1 written in step t-1 and overwrites planes
//     that step t-1 read from the same chunks, so it may start as soon as those three chunks have completed
//     step t-1 (one counter per chunk, bumped with a release by every finished tile);
//   * tickets are handed out in increasing order and a tile only ever waits for smaller tickets, which are
//     held by co-resident CTAs (cooperative launch): no deadlock, no grid-wide barrier, and the tail of step t
//     overlaps the head of step t+1;
//   * Mur / radiation planes, point sources and probes are served by the tile that owns their cells, right after its
//     stores (the values are final there: sponge applied, solver.py:2386-2439 order), so no separate kernel runs per
//     step.  A plane update reads and writes a face cell, its interior neighbour and one `prev` entry: the host only
//     picks tile shapes that keep each such pair inside one tile, so running the list of planes in order inside every
//     tile (a block barrier between two planes: they meet on the edges of the grid) is the sequential order of
//     boundaries/_boundaries.py.
//
// Results are bit-identical to the step-by-step path: same tile code, same operations per cell.
#pragma once
#include "sb_kernels.cuh"

namespace sb {

constexpr int K6_MAX_OPS = SB_MAX_PLANE_OPS;

struct PipeParams {
    StepParams S;                          // tables, extents, tile shape (chunk_i); S.*_in = set holding the state on entry
    int n_steps;
    int gx, gy, nchunks;                   // tiles of one step: gx * gy per chunk of planes
    int *ticket;                           // zero on entry
    int *done;                             // [nchunks] tiles completed per chunk, all steps together; zero on entry
    int *err_flag;
    const double *src_vals;                // [n_steps][n_sources]
    int n_sources;
    float *rec;                            // [n_steps][n_rec]
    int n_rec, n_probes;
    const int *probe_ijk;                  // 3 ints per probe
    int n_ops;                             // Mur / radiation planes, list order (strip mapping only)
    PlaneOp ops[K6_MAX_OPS];
};

__device__ __forceinline__ int ld_acquire_gpu_s32(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// the plane updates of the cells a strip-mode tile owns: i0 <= i < i1, j0 <= j < j1, k0 <= k < k1 (already clipped)
__device__ __noinline__ void k6_plane_ops(const PipeParams &Q, float *p, int i0, int i1, int j0, int j1, int k0, int k1)
{
    const StepParams &P = Q.S;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    const int n[3] = {P.nx, P.ny, P.nz}, lo[3] = {i0, j0, k0}, hi[3] = {i1, j1, k1};
    int last_ax = -1;
    for (int o = 0; o < Q.n_ops; o++) {
        const PlaneOp &op = Q.ops[o];
        const int ax = op.axis, a_ax = ax == 0 ? 1 : 0, b_ax = ax == 2 ? 1 : 2;
        const int face = op.side ? n[ax] - 1 : 0, inner = op.side ? n[ax] - 2 : 1;
        if (face >= lo[ax] && face < hi[ax]) {                                   // (uniform over the block)
            // planes on another axis may meet the previous ones on an edge; on the same axis a thread meets the same
            // (a, b) column again and its own program order is enough
            if (last_ax >= 0 && last_ax != ax) __syncthreads();
            last_ax = ax;
            const int eb = hi[b_ax] - lo[b_ax], cells = (hi[a_ax] - lo[a_ax]) * eb;
            const long long stride[3] = {P.plane, (long long)P.pitch, 1};
            for (int idx = tid; idx < cells; idx += nt) {
                const int a = lo[a_ax] + idx / eb, b = lo[b_ax] + idx % eb;
                const long long base = (long long)a * stride[a_ax] + (long long)b * stride[b_ax];
                const long long ib = base + face * stride[ax], ii = base + inner * stride[ax];
                float *prev = op.prev + (long long)a * n[b_ax] + b;
                const float pi = __ldcg(p + ii);
                __stcg(p + ib, plane_op_value(op, __ldcg(prev), __ldcg(p + ib), pi));
                __stcg(prev, pi);
            }
        }
    }
    if (last_ax >= 0) __syncthreads();                                           // sources and probes come next
}

#ifdef SB_K6_MINB_FIXED                    // experiments
#define SB_K6_MINB(RJ, UNI) SB_K6_MINB_FIXED
#else
#define SB_K6_MINB(RJ, UNI) (((RJ) == 1 && (UNI)) ? 4 : 2)
#endif
template <int RJ, bool GEOM, bool UNI, bool FLAT>
__global__ void __launch_bounds__(256, SB_K6_MINB(RJ, UNI)) k6_pipeline(const __grid_constant__ PipeParams Q)
{
    __shared__ int s_ticket;
    const StepParams &P = Q.S;
    const bool lead = threadIdx.x == 0 && threadIdx.y == 0;
    const int tiles_chunk = Q.gx * Q.gy;
    const long long tiles_step = (long long)tiles_chunk * Q.nchunks, total = tiles_step * Q.n_steps;
    const int rows_tile = RJ * blockDim.y, cols_tile = 4 * blockDim.x;          // strip mode: cells of a tile along j and k
    const int P4 = P.pitch >> 2, groups_tile = (int)(blockDim.x * blockDim.y);   // flat mode: float4 groups of a tile
    bool dead = false;
    for (;;) {
        __syncthreads();                                                         // the previous ticket has been read
        if (lead) s_ticket = atomicAdd(Q.ticket, 1);
        __syncthreads();
        const long long tk = s_ticket;
        if (tk >= total) return;
        const int t = (int)(tk / tiles_step);
        const int r = (int)(tk - (long long)t * tiles_step);
        const int c = r / tiles_chunk, tile = r - c * tiles_chunk;
        const int by = tile / Q.gx, bx = tile - by * Q.gx;
        if (t > 0) {
            // acquire: the chunks this tile reads (and overwrites the inputs of) have finished step t-1.  The acquire
            // load also drops this SM's L1 lines, which may still hold these addresses from two steps ago.
            if (lead && !dead) {
                const int want = t * tiles_chunk;
                const long long t0 = clock64();
                for (int cc = (c > 0 ? c - 1 : c); cc <= (c + 1 < Q.nchunks ? c + 1 : c); cc++)
                    while (ld_acquire_gpu_s32(Q.done + cc) < want)
                        if (clock64() - t0 > (4LL << 30)) { atomicExch(Q.err_flag, 3); dead = true; break; }   // ~2 s
            }
            __syncthreads();
        }
        FieldSet F;
        if (t & 1) F = FieldSet{P.p_out, P.vx_out, P.vy_out, P.vz_out, const_cast<float *>(P.p_in), const_cast<float *>(P.vx_in),
                                const_cast<float *>(P.vy_in), const_cast<float *>(P.vz_in)};
        else       F = FieldSet{P.p_in, P.vx_in, P.vy_in, P.vz_in, P.p_out, P.vx_out, P.vy_out, P.vz_out};
        k1_tile<RJ, GEOM, UNI, false, false, FLAT>(P, F, bx, by, c);
        __syncthreads();                                                         // the tile's stores are visible to the block
        if (!FLAT && Q.n_ops) {
            const int i0 = c * P.chunk_i, j0 = by * rows_tile, k0 = bx * cols_tile;
            k6_plane_ops(Q, F.p_out, i0, min(i0 + P.chunk_i, P.nx), j0, min(j0 + rows_tile, P.ny), k0, min(k0 + cols_tile, P.nz));
        }
        if (threadIdx.y == 0 && threadIdx.x < 32) {                              // sources, then probes, of the cells this tile owns
            const int i0 = c * P.chunk_i, i1 = min(i0 + P.chunk_i, P.nx);
            const int j0 = by * rows_tile, j1 = j0 + rows_tile, k0 = bx * cols_tile, k1 = k0 + cols_tile;
            auto owns = [&](int i, int j, int k) -> bool {
                if (i < i0 || i >= i1) return false;
                if (FLAT) return ((long long)(j / RJ) * P4 + (k >> 2)) / groups_tile == bx;
                return j >= j0 && j < j1 && k >= k0 && k < k1;
            };
            if (threadIdx.x == 0)
                for (int q = 0; q < P.n_inline; q++) {                           // float64 add, fp32 store, list order
                    const int i = P.inl_i[q], j = P.inl_j[q], k = P.inl_k[q];
                    if (owns(i, j, k)) {
                        float *cell = F.p_out + (long long)i * P.plane + (long long)j * P.pitch + k;
                        *cell = (float)((double)*cell + __dmul_rn(Q.src_vals[(long long)t * Q.n_sources + P.inl_src[q]], P.inl_weight[q]));
                    }
                }
            __syncwarp();
            for (int q = threadIdx.x; q < Q.n_probes; q += 32) {
                const int i = Q.probe_ijk[3 * q], j = Q.probe_ijk[3 * q + 1], k = Q.probe_ijk[3 * q + 2];
                if (owns(i, j, k))
                    Q.rec[(long long)t * Q.n_rec + q] = F.p_out[(long long)i * P.plane + (long long)j * P.pitch + k];
            }
            __syncwarp();
            if (threadIdx.x == 0) { __threadfence(); atomicAdd(Q.done + c, 1); }  // release: this tile of step t is complete
        }
    }
}

}  // namespace sb
