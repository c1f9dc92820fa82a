// sb_api.cu -- C ABI (include/strata_b200.h) over the sm_100a kernels in sb_kernels.cuh.
// Host side only: table upload, list building, launch sequencing, CUDA-graph caching.
#include "../../include/strata_b200.h"
#include "sb_kernels.cuh"
#include "sb_resident.cuh"
#include "sb_pipeline.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

using namespace sb;

static thread_local std::string g_err;
static int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return 1;
}
#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CHECK_H(h) do { if (!(h)) return fail("null handle"); CU(cudaSetDevice((h)->device)); } while (0)

template <typename T> struct DBuf {
    T *p = nullptr; size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return 0;
        release();
        if (count == 0) count = 1;
        CU(cudaMalloc(&p, count * sizeof(T)));
        n = count; return 0;
    }
    int upload(const T *host, size_t count, cudaStream_t s) {
        if (alloc(count)) return 1;
        if (count) { CU(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s)); CU(cudaStreamSynchronize(s)); }
        return 0;
    }
    int upload(const std::vector<T> &v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

struct Sponge { DBuf<float> x, y, z; };
struct PlaneOpHost { PlaneOp op; DBuf<float> prev; };

struct sb_solver {
    sb_grid_desc d{};
    int device = 0;
    cudaStream_t stream = nullptr;
    long long plane = 0, elems = 0;
    float *set[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    int cur = 0;
    // tables
    DBuf<float> cvx, cvy, cvz, icx, icy, icz;
    bool nonuniform = false, have_coeffs = false;
    float cp = 0.f, cv_uni = 0.f;
    DBuf<uint8_t> mask; bool have_mask = false;
    std::vector<Sponge *> sponges;
    std::vector<PlaneOpHost *> plane_ops;
    // sources / records
    int n_sources = 0, n_src_cells = 0;
    DBuf<long long> src_off; DBuf<int> src_start, src_id, src_field; DBuf<double> src_weight;
    // host copy of small source tables: enables the single-kernel step (injection inside K1)
    int n_src_entries = 0; bool inline_ok = false;
    int inl_i[SB_MAX_INLINE], inl_j[SB_MAX_INLINE], inl_k[SB_MAX_INLINE], inl_src[SB_MAX_INLINE]; double inl_weight[SB_MAX_INLINE];
    int opt_fuse_k3 = 1;
    int tuned[4] = {1, 0, 1, 0}, tuned_key = -1; float tuned_ms = 0.f;
    int n_probes = 0, n_mics = 0;
    DBuf<long long> probe_off, mic_off; DBuf<float> mic_w; DBuf<int> mic_field; bool have_mic_field = false;
    DBuf<double> d_src_vals; DBuf<float> d_record; DBuf<int> d_step_ctr;
    // ADE
    bool have_ade = false;
    long long ade_material_cells = 0;
    AdeTable ade{};
    DBuf<long long> ade_off; DBuf<int> ade_ijk, ade_nbr; DBuf<uint8_t> ade_mat, ade_mat_box; DBuf<float> ade_J, ade_Jp;
    int opt_ade_layout = 0;                // 0 = auto (fused into K1 on a single slab, else compact / dense list), 1 = compact list, 2 = dense box, 3 = fused
    // fused layout (K1-ADE, sb_kernels.cuh): per pole, J buffers dense over its material's bounding box
    struct FusedPole { DBuf<float> b[3]; int nbuf = 0, bi0 = 0, bj0 = 0, bni = 0, bnj = 0; };
    bool ade_fused = false;
    bool ade_concurrent = false;           // list layout whose K2a / K2b run beside K1 (K1 leaves their cells out: ADEX variants)
    FusedPole fpole[MAX_POLES];
    DBuf<uint8_t> ade_matpad; bool ade_multi = false;
    int fbox[6] = {0, 0, 0, 0, 0, 0};      // bounding box of all pole-carrying cells: i0, i1, j0, j1, k0, k1 (inclusive)
    long long ade_phase = 0;               // steps since the ADE state was zeroed: selects the density-pole buffers
    bool mask_alloc_fresh = true;
    cudaStream_t side = nullptr, side_hi = nullptr;                                  // K1 beside K1-ADE; ADE list kernels beside K1 (high priority)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // options
    int opt_kernel = SB_KERNEL_AUTO, opt_rj = 0, opt_wj = 0, opt_wk = 1, opt_chunk_i = 0, opt_graph = -1;   // 0 / -1 = auto
    // graph cache: key = (n_steps, starting set, source-table pointer, record pointer)
    struct GraphKey {
        int n, cur; const void *src, *rec; int phase;
        bool operator<(const GraphKey &o) const {
            return std::tie(n, cur, src, rec, phase) < std::tie(o.n, o.cur, o.src, o.rec, o.phase);
        }
    };
    struct GraphVal { cudaGraphExec_t exec; long long launches; };
    std::map<GraphKey, GraphVal> graphs;
    // stats
    long long steps_done = 0, kernels_launched = 0;
    int last_variant = 0;
    DBuf<double> d_energy;
    // peer-to-peer halo (sb_set_peers)
    bool have_peers = false;
    float *peer_lo_set[2] = {nullptr, nullptr}, *peer_hi_set[2] = {nullptr, nullptr};   // neighbours' p buffers (base)
    int *my_flags = nullptr, *sig_lo = nullptr, *sig_hi = nullptr;
    int peer_lo_nx = 0;
    DBuf<int> d_step_global, d_err;
    // optional per-launch timing of the fused step kernel (SB_OPT_PROFILE); a resident launch covers `steps` steps
    int opt_profile = 0;
    struct ProfEntry { cudaEvent_t e0, e1; int steps; };
    std::vector<ProfEntry> prof;
    // shared-memory-resident kernel (K5, sb_resident.cuh)
    int n_sm = 0; long long smem_optin = 0;
    bool coop_ok = false;                  // cooperative launches available (K5 / K6 need all their CTAs co-resident)
    int opt_res_split = 0, opt_res_min_steps = 4;
    std::vector<int> probe_ijk_host; DBuf<int> d_probe_ijk; DBuf<uint4> d_res_xch;
    unsigned res_epoch = 0;                // step tags of the face exchange keep growing across launches
    bool resident_used = false;
    int res_nbi = 0, res_nbj = 0;
    int res_tuned_key = -1, res_tuned_nbi = 0, res_tuned_nbj = 0;   // box grid measured best for the current configuration
    DBuf<float> d_res_scratch;
    // step-pipelined kernel (K6, sb_pipeline.cuh)
    DBuf<int> d_pipe_ctr;                  // [0] ticket, [1 ..] chunk counters
    long long opt_pipe_min_cells = 6LL << 20, opt_pipe_max_cells = 40LL << 20;   // where pipelining the steps was measured to pay
    int opt_plane_map = 0;                 // how K1 deals the (j, k) plane to warps: 0 = auto, 1 = strips, 2 = flat
    // two sets of staging buffers for sb_step_n_submit / sb_step_n_wait (a chunk in flight while the host prepares the next)
    DBuf<double> stage_src[2]; DBuf<float> stage_rec[2]; cudaEvent_t stage_done[2] = {nullptr, nullptr};
    int cut_done = 0;                      // planes next to each cut already computed for the step about to be enqueued
    int opt_ade_chunk = 0, opt_ade_warps = 0;   // K1-ADE launch shape: planes per tile, warps per block (0 = default)
    int opt_ade_occ = 3;                        // K1-ADE register allocation: blocks of 256 threads per SM aimed at (2 or 3)
};

// Launch shapes measured once per (device, grid, kernel variant) are remembered for the life of the process: a second
// solver of the same configuration (parameter sweeps, reset-and-rebuild loops) does not pay the trial launches again.
struct TuneKey {
    int device, nx, ny, nz, pitch, key, kind;            // kind 0 = K1 launch shape, 1 = K5 box grid
    bool operator<(const TuneKey &o) const {
        return std::tie(device, nx, ny, nz, pitch, key, kind) < std::tie(o.device, o.nx, o.ny, o.nz, o.pitch, o.key, o.kind);
    }
};
struct TuneVal { int v[4]; float ms; };
static std::map<TuneKey, TuneVal> g_tuned;
static std::mutex g_tuned_mu;
static bool tune_lookup(const sb_solver *h, int key, int kind, TuneVal &out);
static void tune_store(const sb_solver *h, int key, int kind, const TuneVal &v);

static void drop_graphs(sb_solver *h)
{
    h->tuned_key = -1; h->res_tuned_key = -1;                       // configuration changed: re-measure the launch shape as well
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
    h->graphs.clear();
}

static bool tune_lookup(const sb_solver *h, int key, int kind, TuneVal &out)
{
    std::lock_guard<std::mutex> lk(g_tuned_mu);
    auto it = g_tuned.find(TuneKey{h->device, h->d.nx, h->d.ny, h->d.nz, h->d.pitch, key, kind});
    if (it == g_tuned.end()) return false;
    out = it->second;
    return true;
}
static void tune_store(const sb_solver *h, int key, int kind, const TuneVal &v)
{
    std::lock_guard<std::mutex> lk(g_tuned_mu);
    g_tuned[TuneKey{h->device, h->d.nx, h->d.ny, h->d.nz, h->d.pitch, key, kind}] = v;
}

extern "C" const char *sb_last_error(void) { return g_err.c_str(); }
extern "C" int sb_abi_version(void) { return SB_ABI_VERSION; }

extern "C" int sb_choose_pitch(int32_t nz, int32_t *pitch_out)
{
    if (nz <= 0 || !pitch_out) return fail("bad nz");
    *pitch_out = (nz + 7) / 8 * 8;             // rows start on 32-byte sectors; float4 access is always aligned
    return 0;
}

extern "C" int64_t sb_field_elems(const sb_grid_desc *d)
{
    int32_t pitch = d->pitch;
    if (pitch == 0) sb_choose_pitch(d->nz, &pitch);
    return (int64_t)(d->nx + 2) * d->ny * pitch;
}

extern "C" int sb_create(const sb_grid_desc *desc, int device, void *stream, sb_solver **out)
{
    if (!desc || !out) return fail("null argument");
    if (desc->nx < 1 || desc->ny < 1 || desc->nz < 1) return fail("grid extents must be >= 1");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail("no CUDA device available (%s); the b200 backend has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail("device %d out of range (%d devices)", device, count);
    CU(cudaSetDevice(device));
    sb_solver *h = new sb_solver();
    h->d = *desc;
    if (h->d.pitch == 0) sb_choose_pitch(h->d.nz, &h->d.pitch);
    if (h->d.pitch < h->d.nz || h->d.pitch % 4) { delete h; return fail("pitch must be >= nz and a multiple of 4"); }
    if ((long long)h->d.ny * h->d.pitch * (long long)(h->d.nx + 2) >= (1LL << 40)) { delete h; return fail("slab too large"); }
    h->device = device;
    h->stream = (cudaStream_t)stream;
    h->plane = (long long)h->d.ny * h->d.pitch;
    h->elems = sb_field_elems(&h->d);
    if (h->d_step_ctr.alloc(1)) { delete h; return 1; }
    if (h->d_energy.alloc(2)) { delete h; return 1; }
    if (h->d_step_global.alloc(1) || h->d_err.alloc(1)) { delete h; return 1; }
    cudaMemset(h->d_step_global.p, 0, sizeof(int));
    cudaMemset(h->d_err.p, 0, sizeof(int));
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) h->n_sm = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess) h->smem_optin = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, device) == cudaSuccess) h->coop_ok = v != 0;
    *out = h;
    return 0;
}

extern "C" int sb_destroy(sb_solver *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    drop_graphs(h);
    for (auto *s : h->sponges) { s->x.release(); s->y.release(); s->z.release(); delete s; }
    for (auto *po : h->plane_ops) { po->prev.release(); delete po; }
    for (DBuf<float> *b : {&h->cvx, &h->cvy, &h->cvz, &h->icx, &h->icy, &h->icz, &h->mic_w, &h->d_record, &h->ade_J, &h->ade_Jp}) b->release();
    h->mask.release(); h->src_off.release(); h->src_start.release(); h->src_id.release(); h->src_field.release();
    h->src_weight.release(); h->probe_off.release(); h->mic_off.release(); h->d_src_vals.release();
    h->d_step_ctr.release(); h->ade_off.release(); h->ade_ijk.release(); h->ade_nbr.release(); h->ade_mat.release(); h->ade_mat_box.release();
    h->d_energy.release(); h->d_step_global.release(); h->d_err.release();
    h->d_probe_ijk.release(); h->d_res_xch.release(); h->d_pipe_ctr.release(); h->d_res_scratch.release();
    for (auto &fp : h->fpole) for (auto &b : fp.b) b.release();
    for (int q = 0; q < 2; q++) { h->stage_src[q].release(); h->stage_rec[q].release(); if (h->stage_done[q]) cudaEventDestroy(h->stage_done[q]); }
    h->ade_matpad.release();
    if (h->side) cudaStreamDestroy(h->side);
    if (h->side_hi) cudaStreamDestroy(h->side_hi);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
    return 0;
}

extern "C" int sb_bind_fields(sb_solver *h, float *const set0[4], float *const set1[4])
{
    CHECK_H(h);
    for (int f = 0; f < 4; f++) {
        if (!set0[f] || !set1[f]) return fail("null field buffer");
        if (((uintptr_t)set0[f] | (uintptr_t)set1[f]) & 15) return fail("field buffers must be 16-byte aligned");
        h->set[0][f] = set0[f]; h->set[1][f] = set1[f];
    }
    h->cur = 0;
    drop_graphs(h);
    return 0;
}

extern "C" int sb_current_set(sb_solver *h, int *set_out) { if (!h || !set_out) return fail("null argument"); *set_out = h->cur; return 0; }

static inline float *plane0(sb_solver *h, int set, int f) { return h->set[set][f] + h->plane; }

extern "C" int sb_upload_field(sb_solver *h, int field, const float *host)
{
    CHECK_H(h);
    if (field < 0 || field > 3 || !host) return fail("bad field / null host pointer");
    if (!h->set[0][0]) return fail("fields not bound");
    const sb_grid_desc &d = h->d;
    CU(cudaMemcpy2DAsync(plane0(h, h->cur, field), (size_t)d.pitch * 4, host, (size_t)d.nz * 4, (size_t)d.nz * 4,
                         (size_t)d.nx * d.ny, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int sb_download_field(sb_solver *h, int field, float *host)
{
    CHECK_H(h);
    if (field < 0 || field > 3 || !host) return fail("bad field / null host pointer");
    if (!h->set[0][0]) return fail("fields not bound");
    const sb_grid_desc &d = h->d;
    CU(cudaMemcpy2DAsync(host, (size_t)d.nz * 4, plane0(h, h->cur, field), (size_t)d.pitch * 4, (size_t)d.nz * 4,
                         (size_t)d.nx * d.ny, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

// Auxiliary (ADE) field of pole `pole` as a dense [nx][ny][nz] host array (zero outside the pole's material), or the
// reverse: the checkpoint / resume counterpart of sb_download_field / sb_upload_field.  which: 0 = J, 1 = J_prev (Lorentz).
// Works for every device layout (fused: per-material boxes with rotating buffers; compact list; dense box).
extern "C" int sb_ade_state(sb_solver *h, int pole, int which, float *host, int upload)
{
    CHECK_H(h);
    if (!host) return fail("null host pointer");
    if (!h->have_ade) return fail("no ADE materials are set");
    if (pole < 0 || pole >= h->ade.n_poles || which < 0 || which > 1) return fail("bad pole / field selector");
    const sb_grid_desc &d = h->d;
    const PoleDev &Q = h->ade.poles[pole];
    if (which == 1 && !Q.is_lorentz) return fail("pole %d is a Debye pole: it has no J_prev", pole);
    CU(cudaStreamSynchronize(h->stream));
    const size_t cells = (size_t)d.nx * d.ny * d.nz;
    if (!upload) memset(host, 0, cells * sizeof(float));
    const cudaMemcpyKind kind = upload ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (h->ade_fused) {
        sb_solver::FusedPole &fp = h->fpole[pole];
        if (fp.nbuf == 0) return 0;                                    // no cell carries this pole's material
        const long long ph = h->ade_phase;
        float *buf = Q.target == 0 ? (fp.nbuf == 3 ? fp.b[which == 0 ? ph % 3 : (ph + 2) % 3].p : fp.b[ph % 2].p)
                                   : fp.b[which].p;
        for (int ii = 0; ii < fp.bni; ii++) {                          // one strided copy per plane of the material's box
            const int i = fp.bi0 + ii;
            if (i < 0 || i >= d.nx) continue;
            float *dev = buf + (size_t)ii * fp.bnj * d.pitch;
            float *hst = host + ((size_t)i * d.ny + fp.bj0) * d.nz;
            if (upload) CU(cudaMemcpy2DAsync(dev, (size_t)d.pitch * 4, hst, (size_t)d.nz * 4, (size_t)d.nz * 4, fp.bnj, kind, h->stream));
            else        CU(cudaMemcpy2DAsync(hst, (size_t)d.nz * 4, dev, (size_t)d.pitch * 4, (size_t)d.nz * 4, fp.bnj, kind, h->stream));
        }
        CU(cudaStreamSynchronize(h->stream));
        return 0;
    }
    const AdeTable &A = h->ade;
    float *base = (which == 0 ? A.J : A.Jp) + (size_t)pole * A.n_cells;
    std::vector<float> tmp((size_t)A.n_cells);
    if (!upload) CU(cudaMemcpy(tmp.data(), base, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    if (A.dense) {
        std::vector<uint8_t> mb((size_t)A.n_cells);
        CU(cudaMemcpy(mb.data(), A.mat_box, mb.size(), cudaMemcpyDeviceToHost));
        for (int ii = 0; ii < A.bx; ii++)
            for (int jj = 0; jj < A.by; jj++)
                for (int kk = 0; kk < A.bz; kk++) {
                    const size_t s = ((size_t)ii * A.by + jj) * A.bz + kk;
                    if (mb[s] != Q.mat_id) continue;
                    float &cell = host[((size_t)(A.bi0 + ii) * d.ny + (A.bj0 + jj)) * d.nz + (A.bk0 + kk)];
                    if (upload) tmp[s] = cell; else cell = tmp[s];
                }
    } else {
        std::vector<int> ijk((size_t)A.n_cells * 3);
        std::vector<uint8_t> cm((size_t)A.n_cells);
        CU(cudaMemcpy(ijk.data(), A.cell_ijk, ijk.size() * sizeof(int), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(cm.data(), A.cell_mat, cm.size(), cudaMemcpyDeviceToHost));
        if (upload) CU(cudaMemcpy(tmp.data(), base, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));   // keep the ghost-plane cells
        for (int s = 0; s < A.n_cells; s++) {
            const int i = ijk[3 * (size_t)s], j = ijk[3 * (size_t)s + 1], k = ijk[3 * (size_t)s + 2];
            if (i < 0 || i >= d.nx || cm[s] != Q.mat_id) continue;    // (ghost-plane cells of a slab belong to the neighbour)
            float &cell = host[((size_t)i * d.ny + j) * d.nz + k];
            if (upload) tmp[s] = cell; else cell = tmp[s];
        }
    }
    if (upload) {
        if (A.dense) {                                                 // cells of other materials keep their device values
            std::vector<float> cur((size_t)A.n_cells);
            std::vector<uint8_t> mb((size_t)A.n_cells);
            CU(cudaMemcpy(cur.data(), base, cur.size() * sizeof(float), cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(mb.data(), A.mat_box, mb.size(), cudaMemcpyDeviceToHost));
            for (size_t s = 0; s < cur.size(); s++) if (mb[s] != Q.mat_id) tmp[s] = cur[s];
        }
        CU(cudaMemcpy(base, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

// table of n entries padded to `padded` with `fill`, optionally with one leading ghost entry
static int upload_table(DBuf<float> &buf, const float *src, int n, int padded, float fill, cudaStream_t s)
{
    std::vector<float> t((size_t)padded, fill);
    if (src) std::copy(src, src + n, t.begin());
    return buf.upload(t, s);
}

extern "C" int sb_set_coefficients(sb_solver *h, const float *cv_x, const float *cv_y, const float *cv_z,
                                   const float *ic_x, const float *ic_y, const float *ic_z, float cp)
{
    CHECK_H(h);
    if (!cv_x || !cv_y || !cv_z) return fail("null coefficient table");
    const bool nu = ic_x || ic_y || ic_z;
    if (nu && !(ic_x && ic_y && ic_z)) return fail("inv_cell tables must be all NULL or all given");
    const sb_grid_desc &d = h->d;
    const int gx = d.nx + d.has_lower;            // x tables arrive with the live lower ghost first
    {   // x tables are stored with one leading slot so that index -1 is addressable
        std::vector<float> t((size_t)d.nx + 2, 0.0f);
        std::copy(cv_x, cv_x + gx, t.begin() + (d.has_lower ? 0 : 1));
        if (h->cvx.upload(t, h->stream)) return 1;
        if (nu) {
            std::vector<float> u((size_t)d.nx + 2, 1.0f);
            std::copy(ic_x, ic_x + gx, u.begin() + (d.has_lower ? 0 : 1));
            if (h->icx.upload(u, h->stream)) return 1;
        }
    }
    if (upload_table(h->cvy, cv_y, d.ny, d.ny + 4, 0.0f, h->stream)) return 1;
    if (upload_table(h->cvz, cv_z, d.nz, d.pitch + 4, 0.0f, h->stream)) return 1;
    if (nu) {
        if (upload_table(h->icy, ic_y, d.ny, d.ny + 4, 1.0f, h->stream)) return 1;
        if (upload_table(h->icz, ic_z, d.nz, d.pitch + 4, 1.0f, h->stream)) return 1;
    }
    h->nonuniform = nu; h->cp = cp; h->have_coeffs = true;
    h->cv_uni = cv_x[0];
    if (!nu) {                                   // the UNI kernels use one scalar: the tables must really be constant
        for (int q = 0; q < gx; q++) if (cv_x[q] != cv_x[0]) return fail("uniform grid needs a constant cv table");
        for (int q = 0; q < d.ny; q++) if (cv_y[q] != cv_x[0]) return fail("uniform grid needs a constant cv table");
        for (int q = 0; q < d.nz; q++) if (cv_z[q] != cv_x[0]) return fail("uniform grid needs a constant cv table");
    }
    drop_graphs(h);
    return 0;
}

// (re)builds the lower nibble of the mask bytes; geom_dev == nullptr means all air.  The upper nibble (ADE bits) is kept.
static int build_mask(sb_solver *h, const uint8_t *geom_dev, int rigid)
{
    const sb_grid_desc &d = h->d;
    if (!h->mask.p) {
        if (h->mask.alloc((size_t)h->elems)) return 1;
        CU(cudaMemsetAsync(h->mask.p, 0, (size_t)h->elems, h->stream));
    }
    dim3 blk(128), grd((d.nz + 127) / 128, d.ny, d.nx + 2);
    k_build_mask<<<grd, blk, 0, h->stream>>>(geom_dev, h->mask.p + h->plane, d.nx, d.ny, d.nz, d.pitch, h->plane,
                                             d.has_lower, d.has_upper, rigid ? 1 : 0);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    h->kernels_launched++;
    return 0;
}

extern "C" int sb_set_geometry(sb_solver *h, const uint8_t *geom_host, int rigid)
{
    CHECK_H(h);
    drop_graphs(h);
    const sb_grid_desc &d = h->d;
    const size_t gplanes = (size_t)d.nx + d.has_lower + d.has_upper;
    const size_t gbytes = gplanes * d.ny * d.nz;
    // all-air + nothing to zero -> no mask traffic at all
    bool all_air = true;
    if (geom_host) for (size_t q = 0; q < gbytes; q++) if (!geom_host[q]) { all_air = false; break; }
    if (all_air) {
        h->have_mask = false;
        return h->mask.p ? build_mask(h, nullptr, 0) : 0;        // a mask kept for its ADE bits: every cell air, every face open
    }
    DBuf<uint8_t> g;
    if (g.upload(geom_host, gbytes, h->stream)) return 1;
    const int rc = build_mask(h, g.p, rigid);
    g.release();
    if (rc) return 1;
    h->have_mask = true;
    return 0;
}

extern "C" int sb_sponge_decay(const float *sigma, int n, float dt, float *out)
{
    if (!sigma || !out || n < 0) return fail("bad argument");
    for (int m = 0; m < n; m++) out[m] = std::exp(-sigma[m] * dt);     // float overload = glibc expf
    return 0;
}

extern "C" int sb_clear_sponges(sb_solver *h)
{
    CHECK_H(h);
    for (auto *s : h->sponges) { s->x.release(); s->y.release(); s->z.release(); delete s; }
    h->sponges.clear();
    drop_graphs(h);
    return 0;
}

extern "C" int sb_add_sponge(sb_solver *h, const float *dx, const float *dy, const float *dz)
{
    CHECK_H(h);
    if ((int)h->sponges.size() >= MAX_SPONGES) return fail("at most %d sponge layers", MAX_SPONGES);
    const sb_grid_desc &d = h->d;
    Sponge *s = new Sponge();
    std::vector<float> t((size_t)d.nx + 2, 1.0f);         // absent axis = multiply by 1.0f (exact identity)
    if (dx) std::copy(dx, dx + d.nx + d.has_lower, t.begin() + (d.has_lower ? 0 : 1));
    if (s->x.upload(t, h->stream) || upload_table(s->y, dy, d.ny, d.ny + 4, 1.0f, h->stream) ||
        upload_table(s->z, dz, d.nz, d.pitch + 4, 1.0f, h->stream)) { delete s; return 1; }
    h->sponges.push_back(s);
    drop_graphs(h);
    return 0;
}

extern "C" int sb_clear_plane_ops(sb_solver *h)
{
    CHECK_H(h);
    for (auto *po : h->plane_ops) { po->prev.release(); delete po; }
    h->plane_ops.clear();
    drop_graphs(h);
    return 0;
}

extern "C" int sb_plane_op_state(sb_solver *h, int op, float *host, int64_t *elems_out, int upload)
{
    CHECK_H(h);
    if (op < 0 || op >= (int)h->plane_ops.size()) return fail("no plane update %d", op);
    PlaneOpHost *po = h->plane_ops[op];
    if (elems_out) *elems_out = (int64_t)po->prev.n;
    if (!host) return elems_out ? 0 : fail("null host pointer");
    CU(cudaStreamSynchronize(h->stream));
    if (upload) CU(cudaMemcpy(po->prev.p, host, po->prev.n * sizeof(float), cudaMemcpyHostToDevice));
    else        CU(cudaMemcpy(host, po->prev.p, po->prev.n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sb_add_plane_op(sb_solver *h, int axis, int side, int kind, double mur, double R, int weak_r)
{
    CHECK_H(h);
    if (axis < 0 || axis > 2 || side < 0 || side > 1 || kind < 0 || kind > 1) return fail("bad plane-op arguments");
    // on a slab, an axis-0 face exists only where the slab touches the outer face of the grid
    if (axis == 0 && ((side == 0 && h->d.has_lower) || (side == 1 && h->d.has_upper)))
        return fail("this slab does not own the requested axis-0 face");
    const int n[3] = {h->d.nx, h->d.ny, h->d.nz};
    if (n[axis] < (axis == 0 && (h->d.has_lower || h->d.has_upper) ? 2 : 3)) return fail("plane op needs at least 3 cells along its axis");
    PlaneOpHost *po = new PlaneOpHost();
    const size_t cells = (size_t)n[axis == 0 ? 1 : 0] * n[axis == 2 ? 1 : 2];
    if (po->prev.alloc(cells)) { delete po; return 1; }
    CU(cudaMemsetAsync(po->prev.p, 0, cells * 4, h->stream));
    po->op.axis = axis; po->op.side = side; po->op.kind = kind; po->op.weak_r = weak_r ? 1 : 0;
    po->op.mur = mur; po->op.R = R; po->op.one_minus_R = 1 - R; po->op.r32 = (float)R; po->op.prev = po->prev.p;
    h->plane_ops.push_back(po);
    drop_graphs(h);
    return 0;
}

static inline long long dense_to_off(const sb_solver *h, long long dense)
{
    const sb_grid_desc &d = h->d;
    const long long pl = (long long)d.nz * d.ny;
    const long long i = dense >= 0 ? dense / pl : -((-dense + pl - 1) / pl);       // floor: the lower ghost plane is i = -1
    const long long r = dense - i * pl, k = r % d.nz, j = r / d.nz;
    return i * h->plane + j * d.pitch + k;
}

extern "C" int sb_set_sources(sb_solver *h, int n_sources, int n_cells, const int64_t *cell_idx,
                              const int32_t *start, const int32_t *src_id, const int32_t *field,
                              const double *weight)
{
    CHECK_H(h);
    drop_graphs(h);
    h->n_sources = n_sources; h->n_src_cells = n_cells;
    h->n_src_entries = 0; h->inline_ok = true;
    if (n_cells == 0) return 0;
    if (!cell_idx || !start || !src_id || !field || !weight) return fail("null source table");
    const long long ncell = (long long)h->d.nx * h->d.ny * h->d.nz;
    std::vector<long long> off((size_t)n_cells);
    const int n_ent = start[n_cells];
    const long long ghost_lo = h->d.has_lower ? -(long long)h->d.ny * h->d.nz : 0;   // vx entries may sit on the lower ghost plane
    for (int u = 0; u < n_cells; u++) {
        if (cell_idx[u] < ghost_lo || cell_idx[u] >= ncell) return fail("source cell %d out of range", u);
        if (cell_idx[u] < 0)
            for (int e = start[u]; e < start[u + 1]; e++)
                if (field[e] != 1) return fail("source cell %d: only vx entries may lie on the ghost plane", u);
        off[u] = dense_to_off(h, cell_idx[u]);
    }
    for (int e = 0; e < n_ent; e++)
        if (src_id[e] < 0 || src_id[e] >= n_sources || field[e] < 0 || field[e] > 3) return fail("bad source entry %d", e);
    h->n_src_entries = n_ent;
    h->inline_ok = n_ent <= SB_MAX_INLINE;
    for (int u = 0; u < n_cells && h->inline_ok; u++)
        for (int e = start[u]; e < start[u + 1]; e++) {
            if (field[e] != 0 || cell_idx[u] < 0) { h->inline_ok = false; break; }
            const long long dn = cell_idx[u];
            h->inl_k[e] = (int)(dn % h->d.nz); h->inl_j[e] = (int)((dn / h->d.nz) % h->d.ny);
            h->inl_i[e] = (int)(dn / ((long long)h->d.nz * h->d.ny));
            h->inl_src[e] = src_id[e]; h->inl_weight[e] = weight[e];
        }
    if (h->src_off.upload(off, h->stream) || h->src_start.upload(start, (size_t)n_cells + 1, h->stream) ||
        h->src_id.upload(src_id, (size_t)n_ent, h->stream) || h->src_field.upload(field, (size_t)n_ent, h->stream) ||
        h->src_weight.upload(weight, (size_t)n_ent, h->stream)) return 1;
    return 0;
}

extern "C" int sb_set_probes(sb_solver *h, int n_probes, const int64_t *flat_idx)
{
    CHECK_H(h);
    drop_graphs(h);
    const long long ncell = (long long)h->d.nx * h->d.ny * h->d.nz;
    std::vector<long long> off((size_t)n_probes);
    for (int t = 0; t < n_probes; t++) {
        if (flat_idx[t] < 0 || flat_idx[t] >= ncell) return fail("probe %d out of range", t);
        off[t] = dense_to_off(h, flat_idx[t]);
    }
    h->n_probes = n_probes;
    h->probe_ijk_host.assign((size_t)n_probes * 3, 0);
    for (int t = 0; t < n_probes; t++) {
        const long long dn = flat_idx[t];
        h->probe_ijk_host[3 * t] = (int)(dn / ((long long)h->d.nz * h->d.ny));
        h->probe_ijk_host[3 * t + 1] = (int)((dn / h->d.nz) % h->d.ny);
        h->probe_ijk_host[3 * t + 2] = (int)(dn % h->d.nz);
    }
    if (!n_probes) return 0;
    return h->probe_off.upload(off, h->stream) || h->d_probe_ijk.upload(h->probe_ijk_host, h->stream);
}

extern "C" int sb_set_gathers(sb_solver *h, int n, const int32_t *field, const int64_t *idx8, const float *w8)
{
    if (sb_set_mics(h, n, idx8, w8)) return 1;
    h->have_mic_field = false;
    if (!n || !field) return 0;
    for (int t = 0; t < n; t++) if (field[t] < 0 || field[t] > 3) return fail("gather %d: bad field", t);
    if (h->mic_field.upload(field, (size_t)n, h->stream)) return 1;
    h->have_mic_field = true;
    return 0;
}

extern "C" int sb_set_mics(sb_solver *h, int n_mics, const int64_t *idx8, const float *w8)
{
    CHECK_H(h);
    drop_graphs(h);
    h->have_mic_field = false;
    const long long ncell = (long long)h->d.nx * h->d.ny * h->d.nz;
    std::vector<long long> off((size_t)n_mics * 8);
    for (int t = 0; t < n_mics * 8; t++) {
        if (idx8[t] < 0 || idx8[t] >= ncell) return fail("microphone corner %d out of range", t);
        off[t] = dense_to_off(h, idx8[t]);
    }
    h->n_mics = n_mics;
    if (!n_mics) return 0;
    return h->mic_off.upload(off, h->stream) || h->mic_w.upload(w8, (size_t)n_mics * 8, h->stream);
}

// microphones.cpp:16-80 restated (fp32 arithmetic, corner order of microphones.hpp:30-32)
extern "C" int sb_mic_tables(const float *gp, int n_mics, int ny, int nz, int64_t *idx8, float *w8)
{
    if (!gp || !idx8 || !w8) return fail("null argument");
    for (int m = 0; m < n_mics; m++) {
        const float g[3] = {gp[3 * m], gp[3 * m + 1], gp[3 * m + 2]};
        int base[3]; float lo[3], hi[3];
        for (int a = 0; a < 3; a++) { base[a] = (int)g[a]; hi[a] = g[a] - (float)base[a]; lo[a] = 1.0f - hi[a]; }
        for (int c = 0; c < 8; c++) {
            const int bx = c & 1, by = (c >> 1) & 1, bz = (c >> 2) & 1;
            idx8[8 * m + c] = ((int64_t)(base[0] + bx) * ny + (base[1] + by)) * nz + (base[2] + bz);
            w8[8 * m + c] = ((bx ? hi[0] : lo[0]) * (by ? hi[1] : lo[1])) * (bz ? hi[2] : lo[2]);
        }
    }
    return 0;
}

extern "C" int sb_set_ade(sb_solver *h, const sb_pole *poles, int n_poles, const uint8_t *mat,
                          const float *rho_inf, const float *K_inf, int n_ids, float dt, float inv_dx)
{
    CHECK_H(h);
    drop_graphs(h);
    h->have_ade = false; h->ade_fused = false; h->ade_concurrent = false;
    if (n_poles == 0 || !mat) return 0;
    if (n_poles > MAX_POLES) return fail("at most %d ADE poles", MAX_POLES);
    const sb_grid_desc &d = h->d;
    bool used[256] = {false};
    AdeTable &A = h->ade;
    A = AdeTable{};
    for (int q = 0; q < n_poles; q++) {
        const sb_pole &s = poles[q];
        if (s.material_id < 1 || s.material_id > 255 || s.material_id >= n_ids) return fail("pole %d: bad material id", q);
        used[s.material_id] = true;
        PoleDev &Q = A.poles[q];
        Q.mat_id = s.material_id; Q.is_lorentz = s.is_lorentz; Q.target = s.target;
        Q.c0 = s.c0; Q.c1 = s.c1; Q.c2 = s.c2;
        Q.vcoef = -dt / rho_inf[s.material_id] * inv_dx;          // ade.cpp:242 (fp32, left to right)
        Q.pcoef = -K_inf[s.material_id] * dt;                     // ade.cpp:417
    }
    // Compact list of the cells whose material carries poles, in dense order, over the owned planes AND the live
    // ghost planes of a slab: a ghost cell's density-pole J is advanced redundantly from the ghost p plane (same
    // operations as its owner's), so the velocity correction across a cut needs no extra exchange.  Neighbour
    // slots come from rolling plane maps.  `mat` arrives with the ghost planes, lower one first.
    const size_t pl = (size_t)d.ny * d.nz;
    const int i_lo = -d.has_lower, i_hi = d.nx + d.has_upper;     // planes [i_lo, i_hi)
    auto mplane = [&](int i) { return mat + (size_t)(i - i_lo) * pl; };
    std::vector<long long> off; std::vector<int> ijk; std::vector<uint8_t> cm;
    std::vector<long long> plane_start((size_t)(i_hi - i_lo) + 1, 0);
    int b_lo[3] = {1 << 30, 1 << 30, 1 << 30}, b_hi[3] = {-1, -1, -1};        // bounding box of the pole-carrying cells
    std::vector<int> m_lo(256 * 3, 1 << 30), m_hi(256 * 3, -1);               // ... and of every material by itself
    for (int i = i_lo; i < i_hi; i++) {
        const uint8_t *m = mplane(i);
        long long cnt = 0;
        for (int j = 0; j < d.ny; j++) {
            const uint8_t *row = m + (size_t)j * d.nz;
            int k_first = -1, k_last = -1;
            for (int k = 0; k < d.nz; k++) {
                const uint8_t id = row[k];
                if (!used[id]) continue;
                cnt++; if (k_first < 0) k_first = k; k_last = k;
                int *lo = &m_lo[3 * id], *hi = &m_hi[3 * id];
                if (i < lo[0]) lo[0] = i; if (i > hi[0]) hi[0] = i;
                if (j < lo[1]) lo[1] = j; if (j > hi[1]) hi[1] = j;
                if (k < lo[2]) lo[2] = k; if (k > hi[2]) hi[2] = k;
            }
            if (k_first >= 0) {
                b_lo[0] = std::min(b_lo[0], i); b_hi[0] = std::max(b_hi[0], i);
                b_lo[1] = std::min(b_lo[1], j); b_hi[1] = std::max(b_hi[1], j);
                b_lo[2] = std::min(b_lo[2], k_first); b_hi[2] = std::max(b_hi[2], k_last);
            }
        }
        plane_start[i - i_lo + 1] = plane_start[i - i_lo] + cnt;
    }
    const long long n = plane_start[i_hi - i_lo];
    if (n == 0) return 0;
    if (n >= (1LL << 31)) return fail("too many material cells");
    const bool slab = d.has_lower || d.has_upper;
    // Fused layout (K1-ADE): the material cells are updated by a variant of the step kernel itself instead of being
    // recomputed afterwards.  Per pole, J lives in buffers that are dense over the bounding box of the pole's material
    // (in i and j; whole padded rows in k): 2 buffers for a density Debye pole, 3 for a density Lorentz pole (its
    // neighbours re-derive the new value from the old ones), 1 / 2 (J, J_prev) for modulus poles, updated in place.
    if (h->opt_ade_layout == 3 && slab) return fail("the fused ADE layout is not available on decomposed slabs");
    const bool march_ok = h->opt_kernel == SB_KERNEL_AUTO || h->opt_kernel == SB_KERNEL_MARCH;
    if (h->opt_ade_layout == 3 && !march_ok) return fail("the fused ADE layout needs the marching kernel");
    // ADE bits in the upper nibble of the mask bytes (lower nibble: geometry, all open if there is none) and, with
    // several materials, the ids in the field layout: what K1-ADE and the ADEX variants of K1 read
    int n_mat = 0;
    for (int id = 1; id < 256; id++) if (m_hi[3 * id] >= 0) n_mat++;
    auto write_ade_bits = [&](bool want_ids) -> int {
        if (!h->mask.p && build_mask(h, nullptr, 0)) return 1;
        DBuf<uint8_t> md;
        if (md.upload(mat, (size_t)(i_hi - i_lo) * pl, h->stream)) return 1;
        h->ade_multi = want_ids && n_mat > 1;
        if (h->ade_multi && h->ade_matpad.alloc((size_t)h->elems)) { md.release(); return 1; }
        if (h->ade_multi) CU(cudaMemsetAsync(h->ade_matpad.p, 0, (size_t)h->elems, h->stream));
        UsedIds U;
        for (int id = 0; id < 256; id++) U.used[id] = used[id] ? 1 : 0;
        dim3 blk(128), grd((d.nz + 127) / 128, d.ny, d.nx + 2);
        k_ade_bits<<<grd, blk, 0, h->stream>>>(md.p, h->mask.p + h->plane, h->ade_multi ? h->ade_matpad.p + h->plane : nullptr, U,
                                               d.nx, d.ny, d.nz, d.pitch, h->plane, d.has_lower, d.has_upper);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
        md.release();
        h->kernels_launched++;
        for (int a = 0; a < 3; a++) { h->fbox[2 * a] = b_lo[a]; h->fbox[2 * a + 1] = b_hi[a]; }
        return 0;
    };
    // Automatic choice on one GPU: materials that fill a sizeable part of the grid are updated inside the step kernel
    // (fused); a sparse inclusion -- config 3's sphere is 0.4 % of the cells -- costs less as two small list kernels that
    // run BESIDE K1 (K1-ADE's long-running, register-heavy blocks would displace more of K1 than they save).
    const long long grid_cells = (long long)d.nx * d.ny * d.nz;
    const bool want_fused = h->opt_ade_layout == 3 || (h->opt_ade_layout == 0 && (double)n >= 0.02 * (double)grid_cells);
    if (!slab && march_ok && want_fused) {
        size_t total = 0, free_b = 0, total_b = 0;
        for (int q = 0; q < n_poles; q++) {
            const int id = poles[q].material_id;
            sb_solver::FusedPole &fp = h->fpole[q];
            fp.nbuf = 0;
            if (m_hi[3 * id] < 0) continue;                       // registered, but no cell carries it
            fp.bi0 = m_lo[3 * id]; fp.bj0 = m_lo[3 * id + 1];
            fp.bni = m_hi[3 * id] - fp.bi0 + 1; fp.bnj = m_hi[3 * id + 1] - fp.bj0 + 1;
            fp.nbuf = poles[q].target == 0 ? (poles[q].is_lorentz ? 3 : 2) : (poles[q].is_lorentz ? 2 : 1);
            total += (size_t)fp.nbuf * fp.bni * fp.bnj * d.pitch * sizeof(float);
        }
        cudaMemGetInfo(&free_b, &total_b);
        if (total <= free_b / 2 || h->opt_ade_layout == 3) {
            for (int q = 0; q < n_poles; q++) {
                sb_solver::FusedPole &fp = h->fpole[q];
                const size_t cnt = (size_t)fp.bni * fp.bnj * d.pitch;
                for (int b = 0; b < fp.nbuf; b++) {
                    if (fp.b[b].alloc(cnt)) return 1;
                    CU(cudaMemsetAsync(fp.b[b].p, 0, fp.b[b].n * sizeof(float), h->stream));
                }
            }
            if (write_ade_bits(true)) return 1;
            A.n_cells = (int)std::min<long long>(n, (1LL << 31) - 1); A.n_poles = n_poles; A.inv_dx = inv_dx;
            h->ade_material_cells = n;
            h->ade_phase = 0;
            h->ade_fused = true;
            h->have_ade = true;
            return 0;
        }
    }
    if (!slab && march_ok) {                                      // list layouts on one GPU: K2a / K2b run beside K1
        if (write_ade_bits(false)) return 1;
        h->ade_concurrent = true;
    } else if (h->mask.p) {                                       // decomposed slabs: no ADE bits in the mask bytes
        dim3 blk(128), grd((d.nz + 127) / 128, d.ny, d.nx + 2);
        k_ade_bits<<<grd, blk, 0, h->stream>>>(nullptr, h->mask.p + h->plane, nullptr, UsedIds{}, d.nx, d.ny, d.nz, d.pitch, h->plane,
                                               d.has_lower, d.has_upper);
        CU(cudaGetLastError());
    }
    // Dense layout: when the materials fill most of their bounding box, index lists and slot indirections cost more
    // than the few empty cells of the box (a thread per box cell, neighbours found geometrically).  Single slab only.
    const long long box_cells = (long long)(b_hi[0] - b_lo[0] + 1) * (b_hi[1] - b_lo[1] + 1) * (b_hi[2] - b_lo[2] + 1);
    if (h->opt_ade_layout == 2 && slab) return fail("the dense ADE layout is not available on decomposed slabs");
    if (!slab && box_cells < (1LL << 31) && b_hi[0] - b_lo[0] < 65535 && b_hi[1] - b_lo[1] < 65535 &&
        (h->opt_ade_layout == 2 || (h->opt_ade_layout == 0 && (double)n >= 0.6 * (double)box_cells))) {
        const int bx = b_hi[0] - b_lo[0] + 1, by = b_hi[1] - b_lo[1] + 1, bz = b_hi[2] - b_lo[2] + 1;
        std::vector<uint8_t> mb((size_t)box_cells);
        for (int ii = 0; ii < bx; ii++)
            for (int jj = 0; jj < by; jj++) {
                const uint8_t *row = mplane(b_lo[0] + ii) + (size_t)(b_lo[1] + jj) * d.nz + b_lo[2];
                uint8_t *dst = mb.data() + ((size_t)ii * by + jj) * bz;
                for (int kk = 0; kk < bz; kk++) dst[kk] = used[row[kk]] ? row[kk] : 0;
            }
        if (h->ade_mat_box.upload(mb, h->stream)) return 1;
        if (h->ade_J.alloc((size_t)box_cells * n_poles) || h->ade_Jp.alloc((size_t)box_cells * n_poles)) return 1;
        CU(cudaMemsetAsync(h->ade_J.p, 0, (size_t)box_cells * n_poles * 4, h->stream));
        CU(cudaMemsetAsync(h->ade_Jp.p, 0, (size_t)box_cells * n_poles * 4, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        A.n_cells = (int)box_cells; A.n_poles = n_poles;
        A.dense = 1; A.bi0 = b_lo[0]; A.bj0 = b_lo[1]; A.bk0 = b_lo[2]; A.bx = bx; A.by = by; A.bz = bz;
        A.mat_box = h->ade_mat_box.p;
        A.J = h->ade_J.p; A.Jp = h->ade_Jp.p; A.inv_dx = inv_dx;
        h->ade_material_cells = n;
        h->have_ade = true;
        return 0;
    }
    off.resize((size_t)n); ijk.resize((size_t)n * 3); cm.resize((size_t)n);
    std::vector<int> nbr((size_t)n * 6, -1);
    std::vector<int> slot_prev(pl, -1), slot_cur(pl, -1), slot_next(pl, -1);
    auto fill_slots = [&](int i, std::vector<int> &slots) {
        if (i < i_lo || i >= i_hi) { std::fill(slots.begin(), slots.end(), -1); return; }
        const uint8_t *m = mplane(i);
        int s = (int)plane_start[i - i_lo];
        for (size_t q = 0; q < pl; q++) slots[q] = used[m[q]] ? s++ : -1;
    };
    fill_slots(i_lo, slot_cur); fill_slots(i_lo + 1, slot_next);
    for (int i = i_lo; i < i_hi; i++) {
        const uint8_t *m = mplane(i);
        const uint8_t *mp = i > i_lo ? m - pl : nullptr, *mn = i + 1 < i_hi ? m + pl : nullptr;
        for (int j = 0; j < d.ny; j++)
            for (int k = 0; k < d.nz; k++) {
                const size_t q = (size_t)j * d.nz + k;
                const int s = slot_cur[q];
                if (s < 0) continue;
                const uint8_t id = m[q];
                off[s] = (long long)i * h->plane + (long long)j * d.pitch + k;
                ijk[3 * (size_t)s] = i; ijk[3 * (size_t)s + 1] = j; ijk[3 * (size_t)s + 2] = k; cm[s] = id;
                if (mn && mn[q] == id) nbr[0 * n + s] = slot_next[q];
                if (j + 1 < d.ny && m[q + d.nz] == id) nbr[1 * n + s] = slot_cur[q + d.nz];
                if (k + 1 < d.nz && m[q + 1] == id) nbr[2 * n + s] = slot_cur[q + 1];
                if (mp && mp[q] == id) nbr[3 * n + s] = slot_prev[q];
                if (j > 0 && m[q - d.nz] == id) nbr[4 * n + s] = slot_cur[q - d.nz];
                if (k > 0 && m[q - 1] == id) nbr[5 * n + s] = slot_cur[q - 1];
            }
        slot_prev.swap(slot_cur); slot_cur.swap(slot_next); fill_slots(i + 2, slot_next);
    }
    if (h->ade_off.upload(off, h->stream) || h->ade_ijk.upload(ijk, h->stream) || h->ade_mat.upload(cm, h->stream) ||
        h->ade_nbr.upload(nbr, h->stream)) return 1;
    if (h->ade_J.alloc((size_t)n * n_poles) || h->ade_Jp.alloc((size_t)n * n_poles)) return 1;
    CU(cudaMemsetAsync(h->ade_J.p, 0, (size_t)n * n_poles * 4, h->stream));
    CU(cudaMemsetAsync(h->ade_Jp.p, 0, (size_t)n * n_poles * 4, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    A.n_cells = (int)n; A.n_poles = n_poles;
    A.cell_off = h->ade_off.p; A.cell_ijk = h->ade_ijk.p; A.cell_mat = h->ade_mat.p; A.nbr = h->ade_nbr.p;
    A.J = h->ade_J.p; A.Jp = h->ade_Jp.p; A.inv_dx = inv_dx;
    h->ade_material_cells = n;
    h->have_ade = true;
    return 0;
}

// ------------------------------------------------------------------------------------------ stepping
static void fill_params(sb_solver *h, StepParams &P)
{
    const int in = h->cur, out = 1 - h->cur;
    const sb_grid_desc &d = h->d;
    P.p_in = plane0(h, in, 0); P.vx_in = plane0(h, in, 1); P.vy_in = plane0(h, in, 2); P.vz_in = plane0(h, in, 3);
    P.p_out = plane0(h, out, 0); P.vx_out = plane0(h, out, 1); P.vy_out = plane0(h, out, 2); P.vz_out = plane0(h, out, 3);
    P.mask = h->have_mask ? h->mask.p + h->plane : nullptr;
    P.cvx = h->cvx.p + 1; P.cvy = h->cvy.p; P.cvz = h->cvz.p;
    P.icx = h->nonuniform ? h->icx.p + 1 : nullptr; P.icy = h->nonuniform ? h->icy.p : nullptr; P.icz = h->nonuniform ? h->icz.p : nullptr;
    P.n_sponge = (int)h->sponges.size();
    for (int s = 0; s < MAX_SPONGES; s++) {
        const bool on = s < P.n_sponge;
        P.decx[s] = on ? h->sponges[s]->x.p + 1 : nullptr;
        P.decy[s] = on ? h->sponges[s]->y.p : nullptr;
        P.decz[s] = on ? h->sponges[s]->z.p : nullptr;
    }
    P.cp = h->cp;
    P.nx = d.nx; P.ny = d.ny; P.nz = d.nz; P.pitch = d.pitch; P.plane = h->plane;
    P.has_lower = d.has_lower; P.has_upper = d.has_upper;
    P.i_begin = 0; P.i_end = d.nx; P.chunk_i = d.nx;
    P.peer_lo_p = P.peer_hi_p = nullptr; P.flag_lo = P.flag_hi = nullptr;
    P.step_global = h->d_step_global.p; P.err_flag = h->d_err.p; P.two_range = 0;
    P.cv_uni = h->cv_uni;
    P.n_inline = 0; P.src_row = nullptr; P.rec_prev = 0; P.n_probes = P.n_mics = 0;
    P.probe_off = P.mic_off8 = nullptr; P.mic_field = nullptr; P.mic_w8 = nullptr; P.rec_row = nullptr;
    P.box_mode = 0; P.bi0 = P.bi1 = P.bj0 = P.bj1 = P.bk0 = P.bk1 = 0; P.bx_off = P.by_off = P.bz_off = 0;
    P.ade_mask = nullptr;
    if (h->have_peers) {
        if (d.has_lower) { P.peer_lo_p = h->peer_lo_set[out] + (long long)(h->peer_lo_nx + 1) * h->plane; P.flag_lo = h->my_flags; }
        if (d.has_upper) { P.peer_hi_p = h->peer_hi_set[out]; P.flag_hi = h->my_flags + 1; }
    }
}

static PeerLink peer_link(sb_solver *h, const StepParams &P)
{
    PeerLink L{};
    L.peer_lo_p = P.peer_lo_p; L.peer_hi_p = P.peer_hi_p;
    L.sig_lo = h->have_peers && h->d.has_lower ? h->sig_lo : nullptr;
    L.sig_hi = h->have_peers && h->d.has_upper ? h->sig_hi : nullptr;
    L.step_global = h->d_step_global.p; L.nx = h->d.nx; L.plane = h->plane;
    return L;
}

// ---- K1 dispatch over the compile-time variants <RJ, GEOM, UNI, PEER, FUSE, FLAT> ------------------------
// (PEER launches never inject inline: a slab's step always ends with K3, which also publishes the step flag)
template <int RJ, bool GEOM, bool UNI>
static void launch_march3(bool peer, bool fuse, bool flat, int boxm, const StepParams &P, dim3 grd, dim3 blk, cudaStream_t st)
{
    // (the box variants exist apart so that the plain kernel carries none of their code: a run-time test of the box mode
    //  alone cost the 64-register uniform kernel 3.5 % -- tools/k1_ab.cu, 201.1 -> 194.1 Gcell-updates/s on one box)
    if (boxm == 3) { if (flat) k1_step_march<RJ, GEOM, UNI, false, false, true, 3><<<grd, blk, 0, st>>>(P);
                     else      k1_step_march<RJ, GEOM, UNI, false, false, false, 3><<<grd, blk, 0, st>>>(P); }
    else if (boxm == 1) { if (flat) k1_step_march<RJ, GEOM, UNI, false, false, true, 1><<<grd, blk, 0, st>>>(P);
                          else      k1_step_march<RJ, GEOM, UNI, false, false, false, 1><<<grd, blk, 0, st>>>(P); }
    else if (peer) { if (flat) k1_step_march<RJ, GEOM, UNI, true, false, true><<<grd, blk, 0, st>>>(P);
                     else      k1_step_march<RJ, GEOM, UNI, true, false, false><<<grd, blk, 0, st>>>(P); }
    else if (fuse) { if (flat) k1_step_march<RJ, GEOM, UNI, false, true, true><<<grd, blk, 0, st>>>(P);
                     else      k1_step_march<RJ, GEOM, UNI, false, true, false><<<grd, blk, 0, st>>>(P); }
    else           { if (flat) k1_step_march<RJ, GEOM, UNI, false, false, true><<<grd, blk, 0, st>>>(P);
                     else      k1_step_march<RJ, GEOM, UNI, false, false, false><<<grd, blk, 0, st>>>(P); }
}
static void launch_march(int rj, bool peer, bool fuse, bool flat, const StepParams &P, dim3 grd, dim3 blk, cudaStream_t st,
                         int boxm = 0)
{
    const bool geom = P.mask != nullptr, uni = P.icx == nullptr;
    if (rj == 1) { if (geom) { if (uni) launch_march3<1, true, true>(peer, fuse, flat, boxm, P, grd, blk, st); else launch_march3<1, true, false>(peer, fuse, flat, boxm, P, grd, blk, st); }
                   else      { if (uni) launch_march3<1, false, true>(peer, fuse, flat, boxm, P, grd, blk, st); else launch_march3<1, false, false>(peer, fuse, flat, boxm, P, grd, blk, st); } }
    else         { if (geom) { if (uni) launch_march3<2, true, true>(peer, fuse, flat, boxm, P, grd, blk, st); else launch_march3<2, true, false>(peer, fuse, flat, boxm, P, grd, blk, st); }
                   else      { if (uni) launch_march3<2, false, true>(peer, fuse, flat, boxm, P, grd, blk, st); else launch_march3<2, false, false>(peer, fuse, flat, boxm, P, grd, blk, st); } }
}

// launch shape of the marching kernel: rows per thread, warps along j / k, planes per chunk, tiles along k / j
static int march_shape(sb_solver *h, bool flat_ok, int &rj, int &wj, int &wk, int &chunk, int &gx, int &gy, bool &flat)
{
    const sb_grid_desc &d = h->d;
    // plane mapping: strips of 128 cells per warp unless too many of their lanes would hang over the end of a row
    const double strip_fill = (double)d.nz / ((double)((d.nz + 127) / 128) * 128.0);
    flat = flat_ok && (h->opt_plane_map == 2 || (h->opt_plane_map == 0 && strip_fill < 0.95));
    rj = h->opt_rj; wk = h->opt_wk;
    wj = h->opt_wj; int chunk_opt = h->opt_chunk_i;
    if (rj == 0) {                          // auto: the configuration measured by autotune() for this variant
        rj = h->tuned[0]; wj = h->tuned[1]; wk = h->tuned[2]; chunk_opt = h->tuned[3];
    }
    if (rj != 1 && rj != 2) return fail("rows_per_thread must be 1 or 2");
    const long long groups = (long long)((d.ny + rj - 1) / rj) * (d.pitch / 4);      // flat mode: float4 groups of a plane
    if (wj <= 0) {                          // auto: 8 warps per block unless the grid would be too small
        wj = std::max(1, 8 / wk);
        auto tiles = [&](int wj_) {
            return flat ? (groups + 32LL * wj_ * wk - 1) / (32LL * wj_ * wk)
                        : (long long)((d.nz + 128 * wk - 1) / (128 * wk)) * ((d.ny + rj * wj_ - 1) / (rj * wj_));
        };
        while (wj > 1 && tiles(wj) * ((d.nx + 7) / 8) < 148LL * 6) wj >>= 1;
    }
    if (wj * wk * 32 > 256 || wj < 1 || wk < 1) return fail("warps_j*warps_k must be <= 8");
    if (flat) { gx = (int)((groups + 32LL * wj * wk - 1) / (32LL * wj * wk)); gy = 1; }
    else      { gx = (d.nz + 128 * wk - 1) / (128 * wk); gy = (d.ny + rj * wj - 1) / (rj * wj); }
    chunk = chunk_opt;
    if (chunk <= 0) {                       // enough blocks for ~8 waves of 148 SMs, chunks of 8..64 planes
        const long long want = 148LL * 8;
        long long nchunks = (want + (long long)gx * gy - 1) / ((long long)gx * gy);
        nchunks = std::max(1LL, std::min<long long>(nchunks, (d.nx + 7) / 8));
        chunk = (int)((d.nx + nchunks - 1) / nchunks);
        chunk = std::max(chunk, std::min(d.nx, 8));
        chunk = std::min(chunk, 64);
    }
    return 0;
}

// the box of the dispersive materials in K1's terms: mode 1 = K1 skips it (aligned to its chunks of planes; K1-ADE owns
// it), mode 3 = K1 leaves out, cell by cell, what the ADE list kernels write
static void k1_set_box(const sb_solver *h, StepParams &P, int chunk, int mode)
{
    const sb_grid_desc &d = h->d;
    P.box_mode = mode;
    if (mode == 1) { P.bi0 = h->fbox[0] / chunk * chunk; P.bi1 = std::min(d.nx, (h->fbox[1] / chunk + 1) * chunk); }
    else           { P.bi0 = h->fbox[0]; P.bi1 = h->fbox[1] + 1; }
    P.bj0 = h->fbox[2]; P.bj1 = h->fbox[3] + 1;
    P.bk0 = h->fbox[4] / 4 * 4; P.bk1 = (h->fbox[5] / 4 + 1) * 4;
    P.ade_mask = h->mask.p ? h->mask.p + h->plane : nullptr;
}

static int launch_step_kernel(sb_solver *h, StepParams &P, bool fuse, int boxm = 0)
{
    const sb_grid_desc &d = h->d;
    int variant = h->opt_kernel == SB_KERNEL_AUTO ? SB_KERNEL_MARCH : h->opt_kernel;
    if (h->have_peers) variant = SB_KERNEL_MARCH;   // only K1 pushes halos to peers
    h->last_variant = variant;
    if (variant == SB_KERNEL_NAIVE) {
        P.i_begin = d.has_lower ? -1 : 0; P.i_end = d.nx;
        dim3 blk(128), grd((d.nz + 127) / 128, d.ny, P.i_end - P.i_begin);
        if (P.mask) k0_step_naive<true><<<grd, blk, 0, h->stream>>>(P);
        else        k0_step_naive<false><<<grd, blk, 0, h->stream>>>(P);
        h->kernels_launched++;
        return 0;
    }
    int rj, wj, wk, chunk, gx, gy; bool flat;
    if (march_shape(h, true, rj, wj, wk, chunk, gx, gy, flat)) return 1;
    const dim3 blk(32 * wk, wj);
    if (!h->have_peers) {
        // planes next to a cut that sb_step_cuts_async has already computed for this step are left out
        const int cb = h->cut_done;
        h->cut_done = 0;
        P.i_begin = (cb && d.has_lower) ? cb : 0; P.i_end = (cb && d.has_upper) ? d.nx - cb : d.nx; P.chunk_i = chunk;
        const dim3 grd(gx, gy, (P.i_end - P.i_begin + chunk - 1) / chunk);
        if (grd.z > 65535) return fail("too many i-chunks");
        if (boxm) k1_set_box(h, P, chunk, boxm);
        launch_march(rj, false, fuse && !cb, flat, P, grd, blk, h->stream, boxm);
        h->kernels_launched++;
        return 0;
    }
    // Multi-GPU: the planes next to a cut go through the PEER variant (neighbour flags, NVLink peer stores); the
    // interior runs the lean variant first, so the neighbours' flags have arrived long before they are polled.
    const int cb = std::min(8, chunk);
    if (d.nx >= 4 * cb) {
        StepParams Q = P;
        Q.peer_lo_p = Q.peer_hi_p = nullptr; Q.flag_lo = Q.flag_hi = nullptr;
        Q.i_begin = cb; Q.i_end = d.nx - cb; Q.chunk_i = chunk;
        const dim3 grd(gx, gy, (Q.i_end - Q.i_begin + chunk - 1) / chunk);
        launch_march(rj, false, false, flat, Q, grd, blk, h->stream);
        P.two_range = 1; P.i_begin = 0; P.i_end = d.nx; P.chunk_i = cb;
        launch_march(rj, true, false, flat, P, dim3(gx, gy, 2), blk, h->stream);
        h->kernels_launched += 2;
    } else {                                // thin slab: everything through the PEER variant
        P.two_range = 0; P.i_begin = 0; P.i_end = d.nx; P.chunk_i = chunk;
        launch_march(rj, true, false, flat, P, dim3(gx, gy, (d.nx + chunk - 1) / chunk), blk, h->stream);
        h->kernels_launched++;
    }
    return 0;
}

// ---- K1 + K1-ADE: the plain kernel skips the bounding box of the dispersive materials, the ADE variant owns it ----------
static void fill_ade_fused(sb_solver *h, AdeFused &A)
{
    A = AdeFused{};
    A.n_poles = h->ade.n_poles; A.multi = h->ade_multi ? 1 : 0;
    A.mat = h->ade_multi ? h->ade_matpad.p + h->plane : nullptr;
    A.inv_dx = h->ade.inv_dx;
    const long long ph = h->ade_phase;
    for (int q = 0; q < A.n_poles; q++) {
        A.poles[q] = h->ade.poles[q];
        sb_solver::FusedPole &fp = h->fpole[q];
        A.bi0[q] = fp.bi0; A.bj0[q] = fp.bj0; A.bni[q] = fp.bni; A.bnj[q] = fp.bnj;
        if (fp.nbuf == 0) { A.poles[q].target = 2; continue; }               // no cell carries this material: pole unused
        if (A.poles[q].target == 0) {
            if (fp.nbuf == 3) { A.Jin[q] = fp.b[ph % 3].p; A.Jpin[q] = fp.b[(ph + 2) % 3].p; A.Jout[q] = fp.b[(ph + 1) % 3].p; }
            else              { A.Jin[q] = fp.b[ph % 2].p; A.Jpin[q] = nullptr; A.Jout[q] = fp.b[(ph + 1) % 2].p; }
            A.Jpout[q] = nullptr;
        } else {
            A.Jin[q] = A.Jout[q] = fp.b[0].p;
            A.Jpin[q] = A.Jpout[q] = fp.nbuf > 1 ? fp.b[1].p : nullptr;
        }
    }
}

static int launch_step_fused_ade(sb_solver *h, StepParams &P)
{
    const sb_grid_desc &d = h->d;
    h->last_variant = SB_KERNEL_MARCH;
    int rj, wj, wk, chunk, gx, gy; bool flat;
    if (march_shape(h, true, rj, wj, wk, chunk, gx, gy, flat)) return 1;
    if (!h->side) CU(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    if (!h->ev_fork) {
        CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    // the box, aligned to the plain launch's chunks of planes and to float4 groups
    P.i_begin = 0; P.i_end = d.nx; P.chunk_i = chunk;
    k1_set_box(h, P, chunk, 1);
    const dim3 blk(32 * wk, wj), grd(gx, gy, (d.nx + chunk - 1) / chunk);
    if (grd.z > 65535) return fail("too many i-chunks");
    // K1-ADE goes first, on the solver's own stream: its (comparatively few, long-running) blocks take their SM slots the
    // moment the previous step ends; the plain kernel follows on a second stream -- an event wait later -- and fills
    // the rest of the machine.  The other way round the plain kernel's thousands of small blocks hold every SM and the
    // ADE blocks, which need a quarter of a register file each, only get in when it drains (measured: no overlap at all).
    CU(cudaEventRecord(h->ev_fork, h->stream));
    CU(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    // the ADE variant: 1 row per thread; the same plane mapping, restricted to the tiles that meet the box
    StepParams Q = P;
    Q.mask = h->mask.p + h->plane;
    Q.box_mode = 2;
    Q.i_begin = P.bi0; Q.i_end = P.bi1;
    Q.chunk_i = h->opt_ade_chunk > 0 ? h->opt_ade_chunk : std::max(4, std::min(16, (P.bi1 - P.bi0 + 15) / 16));
    const int WJ = h->opt_ade_warps > 0 ? h->opt_ade_warps : 8;
    dim3 ablk(32, WJ), agrd;
    if (flat) {
        const long long P4 = d.pitch / 4, tile = 32LL * WJ;
        Q.bx_off = (int)((long long)P.bj0 * P4 / tile);
        agrd.x = (unsigned)(((long long)P.bj1 * P4 + tile - 1) / tile - Q.bx_off); agrd.y = 1;
    } else {
        Q.bx_off = 0; agrd.x = (unsigned)((P.bk1 - P.bk0 + 127) / 128);          // strips are counted from the box's first cell
        Q.by_off = P.bj0 / WJ;  agrd.y = (unsigned)((P.bj1 + WJ - 1) / WJ - Q.by_off);
    }
    agrd.z = (unsigned)((Q.i_end - Q.i_begin + Q.chunk_i - 1) / Q.chunk_i);
    AdeFused A;
    fill_ade_fused(h, A);
    const bool uni = Q.icx == nullptr;
    if (h->opt_ade_occ == 2) {
        if (uni) { if (flat) k1_step_march_ade<true, true, 2><<<agrd, ablk, 0, h->stream>>>(Q, A); else k1_step_march_ade<true, false, 2><<<agrd, ablk, 0, h->stream>>>(Q, A); }
        else     { if (flat) k1_step_march_ade<false, true, 2><<<agrd, ablk, 0, h->stream>>>(Q, A); else k1_step_march_ade<false, false, 2><<<agrd, ablk, 0, h->stream>>>(Q, A); }
    } else {
        if (uni) { if (flat) k1_step_march_ade<true, true, 3><<<agrd, ablk, 0, h->stream>>>(Q, A); else k1_step_march_ade<true, false, 3><<<agrd, ablk, 0, h->stream>>>(Q, A); }
        else     { if (flat) k1_step_march_ade<false, true, 3><<<agrd, ablk, 0, h->stream>>>(Q, A); else k1_step_march_ade<false, false, 3><<<agrd, ablk, 0, h->stream>>>(Q, A); }
    }
    launch_march(rj, false, false, flat, P, grd, blk, h->side, 1);
    CU(cudaEventRecord(h->ev_join, h->side));
    CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    h->kernels_launched += 2;
    h->ade_phase++;
    return 0;
}

static int autotune(sb_solver *h);
// Host-driven halo exchange (NCCL send/recv) overlapped with the interior: the planes next to the cuts are computed
// first, by themselves, so that the caller can start sending them while the rest of the step runs.
extern "C" int sb_step_cuts_async(sb_solver *h, int *applied)
{
    CHECK_H(h);
    if (!applied) return fail("null argument");
    *applied = 0;
    const sb_grid_desc &d = h->d;
    const int variant = h->opt_kernel == SB_KERNEL_AUTO ? SB_KERNEL_MARCH : h->opt_kernel;
    // only where K1 is the last writer of the cut planes' p: no ADE fix-ups, no plane boundaries; the caller
    // keeps sources off the cut planes.  Anything else takes the serial exchange after the whole step.
    if (h->have_peers || !(d.has_lower || d.has_upper) || variant != SB_KERNEL_MARCH || h->have_ade || !h->plane_ops.empty() ||
        !h->have_coeffs || !h->set[0][0]) return 0;
    if (h->opt_rj == 0 && autotune(h)) return 1;
    int rj, wj, wk, chunk, gx, gy; bool flat;
    if (march_shape(h, true, rj, wj, wk, chunk, gx, gy, flat)) return 1;
    const int cb = std::min(8, chunk);
    if (d.nx < 4 * cb) return 0;
    StepParams P;
    fill_params(h, P);
    const dim3 blk(32 * wk, wj);
    P.chunk_i = cb;
    if (d.has_lower) { P.i_begin = 0; P.i_end = cb; launch_march(rj, false, false, flat, P, dim3(gx, gy, 1), blk, h->stream); h->kernels_launched++; }
    if (d.has_upper) { P.i_begin = d.nx - cb; P.i_end = d.nx; launch_march(rj, false, false, flat, P, dim3(gx, gy, 1), blk, h->stream); h->kernels_launched++; }
    CU(cudaGetLastError());
    h->cut_done = cb;
    *applied = 1;
    return 0;
}

// ---- K1 beside the ADE list kernels: K1 (ADEX variant) leaves the material cells' p and corrected faces out, K2a / K2b
// compute them from the same input set on a second stream -- nothing orders the two until the join.
static int launch_ade_lists(sb_solver *h, const StepParams &P, cudaStream_t st)
{
    const int nb = (h->ade.n_cells + 255) / 256;
    StepParams Q = P; Q.i_begin = 0; Q.i_end = h->d.nx; Q.box_mode = 0;
    if (h->ade.dense) {
        const int tb = h->ade.bz >= 192 ? 256 : (h->ade.bz >= 96 ? 128 : 64);
        const dim3 grd((h->ade.bz + tb - 1) / tb, h->ade.by, h->ade.bx);
        k2a_density_dense<<<grd, tb, 0, st>>>(h->ade, P.p_in, h->d.pitch, h->plane);
        k2b_fixup_dense<<<grd, tb, 0, st>>>(Q, h->ade);
    } else {
        k2a_density<<<nb, 256, 0, st>>>(h->ade, P.p_in);
        k2b_fixup<<<nb, 256, 0, st>>>(Q, h->ade);
    }
    h->kernels_launched += 2;
    return 0;
}

static int launch_step_lists_concurrent(sb_solver *h, StepParams &P)
{
    const sb_grid_desc &d = h->d;
    h->last_variant = SB_KERNEL_MARCH;
    int rj, wj, wk, chunk, gx, gy; bool flat;
    if (march_shape(h, true, rj, wj, wk, chunk, gx, gy, flat)) return 1;
    if (!h->side_hi) {
        // highest priority: the list kernels' small blocks take the slots K1's blocks free as they retire, instead of
        // queueing behind K1's thousands of blocks and running as a tail after it
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CU(cudaStreamCreateWithPriority(&h->side_hi, cudaStreamNonBlocking, prio_hi));
    }
    if (!h->ev_fork) {
        CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    P.i_begin = 0; P.i_end = d.nx; P.chunk_i = chunk;
    k1_set_box(h, P, chunk, 3);
    const dim3 blk(32 * wk, wj), grd(gx, gy, (d.nx + chunk - 1) / chunk);
    if (grd.z > 65535) return fail("too many i-chunks");
    CU(cudaEventRecord(h->ev_fork, h->stream));
    CU(cudaStreamWaitEvent(h->side_hi, h->ev_fork, 0));
    if (launch_ade_lists(h, P, h->side_hi)) return 1;
    CU(cudaEventRecord(h->ev_join, h->side_hi));
    launch_march(rj, false, false, flat, P, grd, blk, h->stream, 3);
    h->kernels_launched++;
    CU(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    return 0;
}

// Launch-shape autotuning.  The best (rows per thread, warps per block, chunk length) depends on the variant
// (uniform / tables, geometry) and on the grid; all shapes give bit-identical results, so the library simply times
// a handful on the live buffers: K1 reads the current set and writes the other one, and without flipping `cur`
// every trial recomputes the same next state, which the real step then overwrites.
static int autotune(sb_solver *h)
{
    static const int cand[][4] = {{1, 8, 1, 64}, {1, 8, 1, 16}, {1, 4, 1, 16}, {1, 2, 2, 16},
                                  {2, 8, 1, 64}, {2, 8, 1, 16}, {2, 4, 1, 16}, {2, 4, 1, 0}, {1, 0, 1, 0}};
    // with dispersive materials the step runs a box variant of K1 (the plain kernel is not what gets launched, and
    // the variants rank the shapes differently: 2 rows per thread are best for the plain 512^3 kernel on some boxes and
    // 12 % behind for the variant that leaves the ADE cells out), so that variant is what the trials time
    const bool march_now = !h->have_peers && (h->opt_kernel == SB_KERNEL_AUTO || h->opt_kernel == SB_KERNEL_MARCH);
    const int boxm = (h->have_ade && march_now) ? (h->ade_fused ? 1 : (h->ade_concurrent ? 3 : 0)) : 0;
    const int key = (h->have_mask ? 1 : 0) | (h->nonuniform ? 2 : 0) | ((int)h->sponges.size() << 2) | (h->have_peers ? 64 : 0) |
                    (boxm << 8) | (boxm ? ((h->fbox[1] - h->fbox[0]) & 0xFFF) << 12 : 0);
    if (h->tuned_key == key) return 0;
    TuneVal cached;
    if (tune_lookup(h, key, 0, cached)) {
        for (int q = 0; q < 4; q++) h->tuned[q] = cached.v[q];
        h->tuned_ms = cached.ms; h->tuned_key = key;
        return 0;
    }
    const int save_rj = h->opt_rj, save_wj = h->opt_wj, save_wk = h->opt_wk, save_ch = h->opt_chunk_i;
    const bool save_peers = h->have_peers;
    const long long k0 = h->kernels_launched;
    h->have_peers = false;                                  // time the lean interior variant
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f; int best_c = 0;
    for (int c = 0; c < (int)(sizeof cand / sizeof cand[0]); c++) {
        h->opt_rj = cand[c][0]; h->opt_wj = cand[c][1]; h->opt_wk = cand[c][2]; h->opt_chunk_i = cand[c][3];
        float t_min = 1e30f;
        for (int rep = 0; rep < 4; rep++) {                  // one warm-up launch, then the best of three
            StepParams P; fill_params(h, P);
            cudaEventRecord(e0, h->stream);
            if (launch_step_kernel(h, P, false, boxm)) { h->have_peers = save_peers; return 1; }
            cudaEventRecord(e1, h->stream);
            cudaEventSynchronize(e1);
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 1) t_min = std::min(t_min, ms);
        }
        if (t_min < best) { best = t_min; best_c = c; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int q = 0; q < 4; q++) h->tuned[q] = cand[best_c][q];
    h->tuned_ms = best; h->tuned_key = key;
    tune_store(h, key, 0, TuneVal{{h->tuned[0], h->tuned[1], h->tuned[2], h->tuned[3]}, best});
    h->opt_rj = save_rj; h->opt_wj = save_wj; h->opt_wk = save_wk; h->opt_chunk_i = save_ch;
    h->have_peers = save_peers; h->kernels_launched = k0;
    CU(cudaGetLastError());
    return 0;
}

// single-kernel step: point sources injected inside K1, previous step's records taken from K1's input set
static bool fused_k3(const sb_solver *h)
{
    const int variant = h->opt_kernel == SB_KERNEL_AUTO ? SB_KERNEL_MARCH : h->opt_kernel;
    // worth it where a launch is a visible fraction of a step (the FUSE variant of K1 is ~9 % slower, a K3 launch
    // costs ~3 us): <= 4 M cells unless forced (opt_fuse_k3 == 2)
    const bool small = (long long)h->d.nx * h->d.ny * h->d.nz <= (4LL << 20);
    return h->opt_fuse_k3 && (small || h->opt_fuse_k3 == 2) && h->inline_ok && !h->have_peers && !h->have_ade &&
           h->plane_ops.empty() && variant == SB_KERNEL_MARCH;
}

static int enqueue_one_step(sb_solver *h, const double *src_dev, float *rec_dev, int step, bool last)
{
    StepParams P;
    fill_params(h, P);
    const int n_rec_all = h->n_probes + h->n_mics;
    const bool fused = fused_k3(h) && !h->cut_done;         // (planes computed ahead by sb_step_cuts_async: K3 runs by itself)
    if (fused) {
        P.n_inline = h->n_src_entries;
        for (int e = 0; e < h->n_src_entries; e++) {
            P.inl_i[e] = h->inl_i[e]; P.inl_j[e] = h->inl_j[e]; P.inl_k[e] = h->inl_k[e];
            P.inl_src[e] = h->inl_src[e]; P.inl_weight[e] = h->inl_weight[e];
        }
        P.src_row = src_dev ? src_dev + (long long)step * h->n_sources : nullptr;
        P.rec_prev = (step > 0 && n_rec_all > 0) ? 1 : 0;
        P.n_probes = h->n_probes; P.n_mics = h->n_mics;
        P.probe_off = h->probe_off.p; P.mic_off8 = h->mic_off.p; P.mic_w8 = h->mic_w.p;
        P.mic_field = h->have_mic_field ? h->mic_field.p : nullptr;
        P.rec_row = rec_dev ? rec_dev + (long long)(step - 1) * n_rec_all : nullptr;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->opt_profile) {
        cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        cudaEventRecord(ev0, h->stream);
    }
    const bool march_now = !h->have_peers && (h->opt_kernel == SB_KERNEL_AUTO || h->opt_kernel == SB_KERNEL_MARCH);
    const bool ade_in_k1 = h->have_ade && h->ade_fused && march_now;
    const bool ade_beside = h->have_ade && !h->ade_fused && h->ade_concurrent && march_now;
    if (ade_in_k1 ? launch_step_fused_ade(h, P) : ade_beside ? launch_step_lists_concurrent(h, P) : launch_step_kernel(h, P, fused)) return 1;
    if (h->opt_profile) { cudaEventRecord(ev1, h->stream); h->prof.push_back({ev0, ev1, 1}); }
    if (h->have_ade && h->ade_fused && !ade_in_k1) return fail("the fused ADE layout needs the marching kernel on a single slab");
    if (h->have_ade && !h->ade_fused && !ade_beside) {
        // after K1: on a slab the density poles of the ghost cells read the ghost p planes, and K1's cut blocks are
        // the ones that wait for the neighbour's step flag (K2a only reads the input set, K1 never touches J)
        if (launch_ade_lists(h, P, h->stream)) return 1;
    }
    for (auto *po : h->plane_ops) {                        // Mur / radiation planes, sequential by construction
        const int n[3] = {h->d.nx, h->d.ny, h->d.nz};
        const int na = n[po->op.axis == 0 ? 1 : 0], nb = n[po->op.axis == 2 ? 1 : 2];
        dim3 blk(128), grd((nb + 127) / 128, na);
        k4_plane_op<<<grd, blk, 0, h->stream>>>(po->op, P.p_out, h->d.nx, h->d.ny, h->d.nz, h->d.pitch, h->plane,
                                                peer_link(h, P));
        h->kernels_launched++;
    }
    const int n_rec = h->n_probes + h->n_mics;
    SourceTable T{h->n_sources, h->n_src_cells, h->src_off.p, h->src_start.p, h->src_id.p, h->src_field.p, h->src_weight.p};
    const PeerLink L = peer_link(h, P);
    if (fused) {
        if (last && n_rec) {                               // the chunk's last step has no successor to record it
            const FieldPtrs F{{P.p_out, P.vx_out, P.vy_out, P.vz_out}};
            k3_record_row<<<(n_rec + 255) / 256, 256, 0, h->stream>>>(F, h->n_probes, h->probe_off.p, h->n_mics,
                                                                      h->have_mic_field ? h->mic_field.p : nullptr,
                                                                      h->mic_off.p, h->mic_w.p,
                                                                      rec_dev + (long long)step * n_rec);
            h->kernels_launched++;
        }
    } else if (h->n_src_cells <= 4096 && n_rec <= 4096) {
        k3_small<<<1, 1024, 0, h->stream>>>(T, P.p_out, P.vx_out, P.vy_out, P.vz_out, src_dev, h->n_probes,
                                            h->probe_off.p, h->n_mics, h->have_mic_field ? h->mic_field.p : nullptr, h->mic_off.p,
                                            h->mic_w.p, rec_dev, h->d_step_ctr.p, L);
        h->kernels_launched++;
    } else {
        if (h->n_src_cells) {
            k3_inject<<<(h->n_src_cells + 255) / 256, 256, 0, h->stream>>>(T, P.p_out, P.vx_out, P.vy_out, P.vz_out,
                                                                            src_dev, h->d_step_ctr.p, L);
            h->kernels_launched++;
        }
        if (n_rec) {
            const FieldPtrs F{{P.p_out, P.vx_out, P.vy_out, P.vz_out}};
            k3_record<<<(n_rec + 255) / 256, 256, 0, h->stream>>>(F, h->n_probes, h->probe_off.p, h->n_mics,
                                                                  h->have_mic_field ? h->mic_field.p : nullptr,
                                                                  h->mic_off.p, h->mic_w.p, rec_dev, h->d_step_ctr.p);
            h->kernels_launched++;
        }
        k3_advance<<<1, 1, 0, h->stream>>>(h->d_step_ctr.p, L);
        h->kernels_launched++;
    }
    h->cur = 1 - h->cur;
    h->steps_done++;
    return 0;
}

// ---- K5: shared-memory-resident chunk (sb_resident.cuh) ---------------------------------------------------
// 0 = not applicable (why_not says why), 1 = plan filled in
static int resident_plan(sb_solver *h, ResParams &R, const char **why_not, int force_nbi = 0, int force_nbj = 0)
{
    const sb_grid_desc &d = h->d;
    *why_not = nullptr;
    if (d.has_lower || d.has_upper || h->have_peers) *why_not = "decomposed slab";
    else if (h->have_ade) *why_not = "ADE materials";
    else if ((int)h->plane_ops.size() > SB_MAX_PLANE_OPS) *why_not = "more than 8 Mur / radiation planes";
    else if (h->n_mics) *why_not = "microphones";
    else if (h->n_src_cells && !h->inline_ok) *why_not = "more than 32 source cells or velocity sources";
    else if (h->n_probes > K5_MAX_PROBES) *why_not = "too many probes";
    else if (h->n_sm <= 0 || h->smem_optin <= 0) *why_not = "device attributes unavailable";
    else if (!h->coop_ok) *why_not = "cooperative launch unavailable on this device / in this process";
    if (*why_not) return 0;
    // a Mur / radiation face and its interior neighbour must lie in one box: two planes / rows per box on such an axis
    int min_li = 1, min_lj = 1;
    for (auto *po : h->plane_ops) { if (po->op.axis == 0) min_li = 2; if (po->op.axis == 1) min_lj = 2; }
    int nbi = force_nbi, nbj = force_nbj;
    if (nbi <= 0 && h->res_tuned_key >= 0) { nbi = h->res_tuned_nbi; nbj = h->res_tuned_nbj; }
    if (nbi > 0 && (d.nx / nbi < min_li || d.ny / nbj < min_lj)) nbi = 0;          // (a grid tuned before the planes were added)
    if (nbi <= 0 && !res_choose_partition(d.nx, d.ny, d.nz, h->n_sm, h->smem_optin, h->n_probes, h->have_mask, &nbi, &nbj,
                                          min_li, min_lj)) {
        *why_not = h->plane_ops.empty() ? "grid does not fit in shared memory"
                                        : "grid does not fit in shared memory in boxes that hold the Mur / radiation face pairs";
        return 0;
    }
    StepParams P;
    fill_params(h, P);
    R = ResParams{};
    for (int q = 0; q < 2; q++) for (int f = 0; f < 4; f++) R.set[q][f] = plane0(h, q, f);
    R.cur = h->cur;
    R.mask = P.mask;
    R.cvx = P.cvx; R.cvy = P.cvy; R.cvz = P.cvz; R.icx = P.icx; R.icy = P.icy; R.icz = P.icz;
    R.n_sponge = P.n_sponge;
    for (int q = 0; q < MAX_SPONGES; q++) { R.decx[q] = P.decx[q]; R.decy[q] = P.decy[q]; R.decz[q] = P.decz[q]; }
    R.cp = h->cp;
    R.nx = d.nx; R.ny = d.ny; R.nz = d.nz; R.pitch = d.pitch; R.plane = h->plane;
    R.nbi = nbi; R.nbj = nbj; R.LI = (d.nx + nbi - 1) / nbi; R.LJ = (d.ny + nbj - 1) / nbj; R.kp = (d.nz + 3) / 4 * 4;
    R.n_inline = h->n_src_cells ? h->n_src_entries : 0;
    for (int e = 0; e < R.n_inline; e++) {
        R.inl_i[e] = h->inl_i[e]; R.inl_j[e] = h->inl_j[e]; R.inl_k[e] = h->inl_k[e];
        R.inl_src[e] = h->inl_src[e]; R.inl_weight[e] = h->inl_weight[e];
    }
    R.n_sources = h->n_sources;
    R.n_probes = h->n_probes; R.n_rec = h->n_probes; R.probe_ijk = h->d_probe_ijk.p;
    R.err_flag = h->d_err.p; R.split = h->opt_res_split;
    R.n_ops = (int)h->plane_ops.size();
    for (int o = 0; o < R.n_ops; o++) R.ops[o] = h->plane_ops[o]->op;
    return 1;
}

struct ResidentLauncher {
    sb_solver *h; ResParams &R; size_t smem;
    template <bool GEOM, bool UNI, int NS> int run()
    {
        auto kern = R.n_ops ? k5_resident<GEOM, UNI, NS, true> : k5_resident<GEOM, UNI, NS, false>;
        static thread_local const void *last_kern = nullptr; static thread_local size_t last_smem = 0; static thread_local int last_per_sm = 0;
        static thread_local int last_dev = -1;
        int per_sm = last_per_sm;
        if (last_kern != (const void *)kern || last_smem != smem || last_dev != h->device) {   // (per launch these two calls cost ~20 us)
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, K5_NT, smem));
            last_kern = (const void *)kern; last_smem = smem; last_per_sm = per_sm; last_dev = h->device;
        }
        if ((long long)per_sm * h->n_sm < (long long)R.nbi * R.nbj) {
            fail("resident kernel: %d boxes cannot be co-resident", R.nbi * R.nbj);
            return 2;
        }
        void *args[] = {&R};
        const cudaError_t e = cudaLaunchCooperativeKernel((const void *)kern, dim3(R.nbi * R.nbj), dim3(K5_NT), args, smem, h->stream);
        if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {   // e.g. SMs partitioned away (MPS / MIG)
            cudaGetLastError();
            fail("resident kernel: cooperative launch refused (%s)", cudaGetErrorString(e));
            return 2;
        }
        CU(e);
        return 0;
    }
};

// enqueue one resident launch of n_steps; touches no solver state except the exchange area and its step tags
static int resident_enqueue(sb_solver *h, ResParams &R, int n_steps, const double *src_dev, float *rec_dev)
{
    const int nb = R.nbi * R.nbj;
    R.n_steps = n_steps; R.src_vals = src_dev; R.rec = rec_dev;
    R.xch_face = std::max(R.LI, R.LJ) * R.kp / 2;
    const size_t need = (size_t)2 * nb * 4 * R.xch_face;
    if (h->d_res_xch.n < need || !h->d_res_xch.p || h->res_epoch > 0xF0000000u) {   // (re)start the tags at zero
        if (h->d_res_xch.alloc(need)) return 1;
        CU(cudaMemsetAsync(h->d_res_xch.p, 0, h->d_res_xch.n * sizeof(uint4), h->stream));
        h->res_epoch = 0;
    }
    R.xch = h->d_res_xch.p; R.tag_base = h->res_epoch;
    h->res_epoch += (unsigned)n_steps;
    const size_t smem = (size_t)res_smem_bytes(R.LI, R.LJ, R.kp, R.n_probes, R.mask != nullptr);
    ResidentLauncher launcher{h, R, smem};
    return res_dispatch(R.mask != nullptr, R.icx == nullptr, R.n_sponge, launcher);   // 2 = not all boxes fit at once
}

// Box grids worth measuring: the heuristic choice (fewest items per thread) and the neighbours that trade an item
// per thread for a shorter box perimeter, i.e. less face exchange -- which of the two decides depends on the grid
// (64^3: 13 x 11 boxes 2.6 us/step, 22 x 6 boxes 3.0; 100^3: 20 x 7 and 7 x 21 equal).
static std::vector<std::pair<int, int>> resident_candidates(const sb_solver *h)
{
    struct Cand { int nbi, nbj; long long iters, bytes; int perim; };
    const sb_grid_desc &d = h->d;
    const int kp = (d.nz + 3) / 4 * 4, K4 = kp / 4;
    std::vector<Cand> all;
    for (int nbi = 1; nbi <= d.nx && nbi <= h->n_sm; nbi++) {
        const int nbj = std::min(d.ny, h->n_sm / nbi);
        if (nbj < 1) break;
        const int LI = (d.nx + nbi - 1) / nbi, LJ = (d.ny + nbj - 1) / nbj;
        if ((long long)LJ * K4 > K5_NT) continue;
        const long long bytes = res_smem_bytes(LI, LJ, kp, h->n_probes, h->have_mask);
        if (bytes > h->smem_optin) continue;
        const int ncol = LJ * K4, G = K5_NT / ncol;
        const long long items = (long long)LI * ncol + ncol + (long long)LI * K4;
        all.push_back({nbi, nbj, std::max((items + K5_NT - 1) / K5_NT, (long long)(LI + G - 1) / G), bytes, LI + LJ});
    }
    std::vector<std::pair<int, int>> out;
    if (all.empty()) return out;
    long long min_iters = all[0].iters;
    for (auto &c : all) min_iters = std::min(min_iters, c.iters);
    auto push = [&](const Cand &c) {
        for (auto &o : out) if (o.first == c.nbi && o.second == c.nbj) return;
        if (out.size() < 4) out.emplace_back(c.nbi, c.nbj);
    };
    std::sort(all.begin(), all.end(), [](const Cand &a, const Cand &b) { return std::tie(a.iters, a.bytes) < std::tie(b.iters, b.bytes); });
    for (size_t q = 0; q < all.size() && q < 3; q++) push(all[q]);                 // [0] = res_choose_partition's choice
    std::vector<Cand> near;
    for (auto &c : all) if (c.iters <= min_iters + 1) near.push_back(c);
    std::sort(near.begin(), near.end(), [](const Cand &a, const Cand &b) { return std::tie(a.perim, a.iters, a.bytes) < std::tie(b.perim, b.iters, b.bytes); });
    for (size_t q = 0; q < near.size() && q < 3; q++) push(near[q]);
    return out;
}

// Times the candidate box grids on the live state: a resident launch reads set[cur] and writes set[(cur + n) & 1], so
// an odd number of trial steps leaves the current state untouched (records go to a scratch buffer).
static int resident_autotune(sb_solver *h, const double *src_dev, int n_steps)
{
    const int key = (h->have_mask ? 1 : 0) | (h->nonuniform ? 2 : 0) | ((int)h->sponges.size() << 2) | (h->n_probes << 6);
    if (h->res_tuned_key == key) return 0;
    if (!h->plane_ops.empty()) return 0;                      // trial steps would advance the planes' `prev` state
    TuneVal cached;
    if (tune_lookup(h, key, 1, cached)) {
        h->res_tuned_nbi = cached.v[0]; h->res_tuned_nbj = cached.v[1]; h->res_tuned_key = key;
        return 0;
    }
    int n_trial = std::min(n_steps, 33);
    if (!(n_trial & 1)) n_trial--;
    if (n_trial < 33) return 0;                               // too short to tell box grids apart: keep the heuristic
    const auto cands = resident_candidates(h);
    if (cands.size() < 2) return 0;
    if (h->d_res_scratch.alloc((size_t)n_trial * std::max(1, h->n_probes))) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f; int best_c = 0;
    h->res_tuned_key = -1;
    for (int c = 0; c < (int)cands.size(); c++) {
        float t_min = 1e30f;
        for (int rep = 0; rep < 2; rep++) {
            ResParams R; const char *why_not = nullptr;
            if (!resident_plan(h, R, &why_not, cands[c].first, cands[c].second)) break;
            cudaEventRecord(e0, h->stream);
            const int rc = resident_enqueue(h, R, n_trial, src_dev, h->d_res_scratch.p);
            if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
            cudaEventRecord(e1, h->stream);
            cudaEventSynchronize(e1);
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) t_min = std::min(t_min, ms);
        }
        if (t_min < best) { best = t_min; best_c = c; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    CU(cudaGetLastError());
    h->res_tuned_nbi = cands[best_c].first; h->res_tuned_nbj = cands[best_c].second; h->res_tuned_key = key;
    tune_store(h, key, 1, TuneVal{{h->res_tuned_nbi, h->res_tuned_nbj, 0, 0}, best});
    return 0;
}

static int launch_resident(sb_solver *h, ResParams &R, int n_steps, const double *src_dev, float *rec_dev)
{
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->opt_profile) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, h->stream); }
    if (const int rc = resident_enqueue(h, R, n_steps, src_dev, rec_dev)) {
        if (h->opt_profile) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
        return rc;                           // 2 = the device cannot hold all boxes at once: the caller falls back
    }
    if (h->opt_profile) { cudaEventRecord(ev1, h->stream); h->prof.push_back({ev0, ev1, n_steps}); }
    h->kernels_launched++;
    h->steps_done += n_steps;
    if (n_steps & 1) h->cur = 1 - h->cur;
    h->last_variant = SB_KERNEL_RESIDENT;
    h->resident_used = true;
    h->res_nbi = R.nbi; h->res_nbj = R.nbj;
    return 0;
}

// ---- K6: step-pipelined persistent launch of the marching kernel (sb_pipeline.cuh) -------------------------
static const char *pipeline_why_not(const sb_solver *h)
{
    const sb_grid_desc &d = h->d;
    if (d.has_lower || d.has_upper || h->have_peers) return "decomposed slab";
    if (h->have_ade) return "ADE materials";
    if ((int)h->plane_ops.size() > SB_MAX_PLANE_OPS) return "more than 8 Mur / radiation planes";
    if (h->n_mics) return "microphones";
    if (h->n_src_cells && !h->inline_ok) return "more than 32 source cells or velocity sources";
    if (h->n_sm <= 0) return "device attributes unavailable";
    if (!h->coop_ok) return "cooperative launch unavailable on this device / in this process";
    return nullptr;
}

// A Mur / radiation plane is updated inside the tile that owns its cells (sb_pipeline.cuh): the face cell and its
// interior neighbour must fall into the same tile.  On the low side that takes a tile two cells thick, on the high side
// an extent n with (n - 1) % thickness != 0.
static bool pipe_tile_holds_planes(const sb_solver *h, int chunk, int rows_tile, int cols_tile)
{
    const int n[3] = {h->d.nx, h->d.ny, h->d.nz}, thick[3] = {chunk, rows_tile, cols_tile};
    for (auto *po : h->plane_ops) {
        const int a = po->op.axis, t = std::min(thick[a], n[a]);
        if (n[a] < 2 || t < 2) return false;
        if (po->op.side == 1 && (n[a] - 1) % t == 0) return false;
    }
    return true;
}

template <int RJ, bool GEOM, bool UNI, bool FLAT>
static int launch_pipeline_w(sb_solver *h, PipeParams &Q, dim3 blk, long long total)
{
    auto kern = k6_pipeline<RJ, GEOM, UNI, FLAT>;
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (int)(blk.x * blk.y), 0));
    if (per_sm < 1) return fail("pipelined kernel does not fit on an SM");
    const long long ctas = std::min<long long>(total, (long long)per_sm * h->n_sm);
    void *args[] = {&Q};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)ctas), blk, args, 0, h->stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        fail("pipelined kernel: cooperative launch refused (%s)", cudaGetErrorString(e));
        return 2;
    }
    CU(e);
    return 0;
}

template <int RJ, bool GEOM, bool UNI>
static int launch_pipeline_t(sb_solver *h, PipeParams &Q, dim3 blk, long long total, bool flat)
{
    return flat ? launch_pipeline_w<RJ, GEOM, UNI, true>(h, Q, blk, total) : launch_pipeline_w<RJ, GEOM, UNI, false>(h, Q, blk, total);
}

static int launch_pipeline(sb_solver *h, int n_steps, const double *src_dev, float *rec_dev)
{
    const sb_grid_desc &d = h->d;
    // Launch shape: unless the caller fixed one, the shape that pipelines best, not the one autotune() found for
    // step-by-step launches -- short tiles (1 row per thread, 4 warps, 64 registers: 8 CTAs per SM) of nx/16 planes
    // (measured, us/step: 200^3 48.2 / 44.8 / 43.8 / 50.4 with 4 / 8 / 12 / 16 planes; 300^3 156.3 / 151.0 / 149.8 with
    // 8 / 16 / 25; tools/k6_chunk_sweep.py).
    const int save[4] = {h->opt_rj, h->opt_wj, h->opt_wk, h->opt_chunk_i};
    if (h->opt_rj == 0) { h->opt_rj = 1; h->opt_wj = 4; h->opt_wk = 1; if (!h->opt_chunk_i) h->opt_chunk_i = std::max(8, std::min(32, d.nx / 16)); }
    int rj, wj, wk, chunk, gx, gy; bool w;
    const bool ops = !h->plane_ops.empty();                 // plane updates live in strip-mode tiles
    int shape_rc = march_shape(h, !ops, rj, wj, wk, chunk, gx, gy, w);
    if (!shape_rc && ops && !pipe_tile_holds_planes(h, chunk, rj * wj, 128 * wk)) {
        // another tile shape that keeps every (face, neighbour) pair together -- only where the caller left the shape open
        bool found = false;
        if (save[0] == 0) {
            static const int shapes[][2] = {{4, 1}, {3, 1}, {5, 1}, {2, 1}, {4, 2}, {3, 2}, {2, 2}, {2, 3}, {1, 3}};
            for (auto &sh : shapes) {
                for (int c = h->opt_chunk_i; c <= std::min(d.nx, h->opt_chunk_i + 3) && !found; c++)
                    if (pipe_tile_holds_planes(h, c, sh[0], 128 * sh[1])) {
                        h->opt_wj = sh[0]; h->opt_wk = sh[1]; h->opt_chunk_i = c;
                        found = true;
                    }
                if (found) break;
            }
        }
        if (found) shape_rc = march_shape(h, false, rj, wj, wk, chunk, gx, gy, w);
        else {
            h->opt_rj = save[0]; h->opt_wj = save[1]; h->opt_wk = save[2]; h->opt_chunk_i = save[3];
            if (h->opt_kernel == SB_KERNEL_AUTO) return 3;
            return fail("pipelined kernel not applicable: no tile shape keeps every Mur / radiation face next to its interior neighbour");
        }
    }
    h->opt_rj = save[0]; h->opt_wj = save[1]; h->opt_wk = save[2]; h->opt_chunk_i = save[3];
    if (shape_rc) return 1;
    PipeParams Q{};
    fill_params(h, Q.S);
    Q.S.i_begin = 0; Q.S.i_end = d.nx; Q.S.chunk_i = chunk;
    Q.S.n_inline = h->n_src_cells ? h->n_src_entries : 0;
    for (int e = 0; e < Q.S.n_inline; e++) {
        Q.S.inl_i[e] = h->inl_i[e]; Q.S.inl_j[e] = h->inl_j[e]; Q.S.inl_k[e] = h->inl_k[e];
        Q.S.inl_src[e] = h->inl_src[e]; Q.S.inl_weight[e] = h->inl_weight[e];
    }
    Q.n_steps = n_steps; Q.gx = gx; Q.gy = gy; Q.nchunks = (d.nx + chunk - 1) / chunk;
    const long long total = (long long)gx * gy * Q.nchunks * n_steps;
    if (total >= (1LL << 31) - 65536) return fail("pipelined kernel: too many tiles in one chunk of steps");
    // Left to itself (AUTO) the pipeline is only used where it was measured to pay: a step must offer at least as many
    // tiles as the GPU holds CTAs and at least four chunks of planes (long thin grids have neither: their few chunks
    // serialise on each other); 3 = "not worthwhile here", the caller goes on to the step-by-step path.
    // With Mur / radiation planes the alternative is one more launch per plane and step, and the pipeline wins from the
    // smallest grids on (us/step with six planes, step-by-step vs pipelined: 64^3 35.1 / 21.7, 100^3 33.0 / 24.2,
    // 160^3 48.8 / 39.7, 200^3 74.9 / 56.1, 256^3 119.3 / 95.4, 300^3 192.4 / 174.7; tools/plane_ops_timing.py).
    if (h->opt_kernel == SB_KERNEL_AUTO && (((long long)gx * gy * Q.nchunks < (long long)h->n_sm * 8 && !ops) || Q.nchunks < 4)) return 3;
    if (h->d_pipe_ctr.alloc((size_t)Q.nchunks + 1)) return 1;
    CU(cudaMemsetAsync(h->d_pipe_ctr.p, 0, ((size_t)Q.nchunks + 1) * sizeof(int), h->stream));
    Q.ticket = h->d_pipe_ctr.p; Q.done = h->d_pipe_ctr.p + 1; Q.err_flag = h->d_err.p;
    Q.src_vals = src_dev; Q.n_sources = h->n_sources;
    Q.rec = rec_dev; Q.n_rec = h->n_probes; Q.n_probes = h->n_probes; Q.probe_ijk = h->d_probe_ijk.p;
    Q.n_ops = (int)h->plane_ops.size();
    for (int o = 0; o < Q.n_ops; o++) Q.ops[o] = h->plane_ops[o]->op;
    const dim3 blk(32 * wk, wj);
    const bool geom = Q.S.mask != nullptr, uni = Q.S.icx == nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (h->opt_profile) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, h->stream); }
    int rc;
    if (rj == 1) rc = geom ? (uni ? launch_pipeline_t<1, true, true>(h, Q, blk, total, w) : launch_pipeline_t<1, true, false>(h, Q, blk, total, w))
                           : (uni ? launch_pipeline_t<1, false, true>(h, Q, blk, total, w) : launch_pipeline_t<1, false, false>(h, Q, blk, total, w));
    else         rc = geom ? (uni ? launch_pipeline_t<2, true, true>(h, Q, blk, total, w) : launch_pipeline_t<2, true, false>(h, Q, blk, total, w))
                           : (uni ? launch_pipeline_t<2, false, true>(h, Q, blk, total, w) : launch_pipeline_t<2, false, false>(h, Q, blk, total, w));
    if (rc) { if (h->opt_profile) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); } return rc; }
    if (h->opt_profile) { cudaEventRecord(ev1, h->stream); h->prof.push_back({ev0, ev1, n_steps}); }
    h->kernels_launched++;
    h->steps_done += n_steps;
    if (n_steps & 1) h->cur = 1 - h->cur;
    h->last_variant = SB_KERNEL_PIPELINE;
    h->resident_used = true;                 // sb_synchronize looks at the error flag
    return 0;
}

extern "C" int sb_step_n_async(sb_solver *h, int n_steps, const double *src_dev, float *rec_dev)
{
    CHECK_H(h);
    if (n_steps < 0) return fail("negative step count");
    if (!h->set[0][0]) return fail("fields not bound");
    if (!h->have_coeffs) return fail("coefficients not set");
    if (h->n_src_cells && !src_dev) return fail("source values required");
    if ((h->n_probes + h->n_mics) && !rec_dev) return fail("record buffer required");
    if (n_steps > 0 && (h->opt_kernel == SB_KERNEL_RESIDENT ||
                        (h->opt_kernel == SB_KERNEL_AUTO && n_steps >= h->opt_res_min_steps))) {
        ResParams R; const char *why_not = nullptr;
        if (resident_plan(h, R, &why_not)) {
            if (src_dev || !h->n_src_cells) {                    // measure the candidate box grids once per configuration
                const int trc = resident_autotune(h, src_dev, n_steps);
                if (trc == 1) return 1;
                if (trc == 0 && !resident_plan(h, R, &why_not)) return fail("resident kernel: %s", why_not);
            }
            const int rc = launch_resident(h, R, n_steps, src_dev, rec_dev);
            if (rc != 2 || h->opt_kernel == SB_KERNEL_RESIDENT) return rc ? 1 : 0;
            h->coop_ok = false;              // not all boxes fit at once here: use the step-by-step path from now on
        } else if (h->opt_kernel == SB_KERNEL_RESIDENT) return fail("resident kernel not applicable: %s", why_not);
    }
    if (n_steps > 0 && (h->opt_kernel == SB_KERNEL_PIPELINE ||
                        (h->opt_kernel == SB_KERNEL_AUTO && n_steps >= h->opt_res_min_steps &&
                         ((long long)h->d.nx * h->d.ny * h->d.nz >= h->opt_pipe_min_cells || !h->plane_ops.empty()) &&
                         (long long)h->d.nx * h->d.ny * h->d.nz <= h->opt_pipe_max_cells))) {
        const char *why_not = pipeline_why_not(h);
        if (!why_not) {
            const int rc = launch_pipeline(h, n_steps, src_dev, rec_dev);
            if ((rc != 2 && rc != 3) || h->opt_kernel == SB_KERNEL_PIPELINE) return rc ? 1 : 0;
            if (rc == 2) h->coop_ok = false;
        } else if (h->opt_kernel == SB_KERNEL_PIPELINE) return fail("pipelined kernel not applicable: %s", why_not);
    }
    if (h->opt_rj == 0 && n_steps > 0) {
        const int variant = h->opt_kernel == SB_KERNEL_AUTO ? SB_KERNEL_MARCH : h->opt_kernel;
        if (variant != SB_KERNEL_NAIVE && autotune(h)) return 1;
    }
    CU(cudaMemsetAsync(h->d_step_ctr.p, 0, sizeof(int), h->stream));
    const bool want_graph = h->opt_graph == 1 ||
        (h->opt_graph < 0 && n_steps >= 4 && (long long)h->d.nx * h->d.ny * h->d.nz <= (32LL << 20));
    if (want_graph && n_steps > 1 && !h->opt_profile) {
        {
            sb_solver::GraphKey key{n_steps, h->cur, src_dev, rec_dev, (h->have_ade && h->ade_fused) ? (int)(h->ade_phase % 6) : 0};
            if (h->graphs.size() > 64) drop_graphs(h);
            auto it = h->graphs.find(key);
            if (it == h->graphs.end()) {
                cudaGraph_t g;
                const long long k0 = h->kernels_launched; const int cur0 = h->cur; const long long s0 = h->steps_done;
                const long long ph0 = h->ade_phase;
                CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                int rc = 0;
                for (int s = 0; s < n_steps && !rc; s++) rc = enqueue_one_step(h, src_dev, rec_dev, s, s == n_steps - 1);
                cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
                h->cur = cur0; h->steps_done = s0; h->ade_phase = ph0;
                const long long per_launch = h->kernels_launched - k0; h->kernels_launched = k0;
                if (rc) return 1;
                if (ce != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(ce));
                cudaGraphExec_t ge;
                CU(cudaGraphInstantiate(&ge, g, 0));
                cudaGraphDestroy(g);
                it = h->graphs.emplace(key, sb_solver::GraphVal{ge, per_launch}).first;
            }
            CU(cudaGraphLaunch(it->second.exec, h->stream));
            h->kernels_launched += it->second.launches;    // bookkeeping equivalent to n_steps enqueues
            h->steps_done += n_steps;
            if (h->have_ade && h->ade_fused) h->ade_phase += n_steps;   // the replayed steps rotated the density-pole buffers
            if (n_steps & 1) h->cur = 1 - h->cur;
            return 0;
        }
    }
    for (int s = 0; s < n_steps; s++)
        if (enqueue_one_step(h, src_dev, rec_dev, s, s == n_steps - 1)) return 1;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sb_step_n(sb_solver *h, int n_steps, const double *src_host, float *rec_host)
{
    CHECK_H(h);
    if (n_steps <= 0) return n_steps == 0 ? 0 : fail("negative step count");
    const int n_rec = h->n_probes + h->n_mics;
    if (h->n_sources && h->n_src_cells && !src_host) return fail("source values required");
    const void *old_src = h->d_src_vals.p, *old_rec = h->d_record.p;
    if (h->d_src_vals.alloc((size_t)std::max(1, n_steps * std::max(1, h->n_sources)))) return 1;
    if (h->d_record.alloc((size_t)std::max(1, n_steps * std::max(1, n_rec)))) return 1;
    if (old_src != h->d_src_vals.p || old_rec != h->d_record.p) drop_graphs(h);   // cached graphs hold these pointers
    if (h->n_sources && src_host)
        CU(cudaMemcpyAsync(h->d_src_vals.p, src_host, (size_t)n_steps * h->n_sources * sizeof(double),
                           cudaMemcpyHostToDevice, h->stream));
    if (sb_step_n_async(h, n_steps, h->d_src_vals.p, h->d_record.p)) return 1;
    if (n_rec && rec_host)
        CU(cudaMemcpyAsync(rec_host, h->d_record.p, (size_t)n_steps * n_rec * sizeof(float),
                           cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (h->resident_used) return sb_synchronize(h);
    return 0;
}

extern "C" int sb_step_n_submit(sb_solver *h, int slot, int n_steps, const double *src_host, float *rec_host)
{
    CHECK_H(h);
    if (slot < 0 || slot > 1) return fail("slot must be 0 or 1");
    if (n_steps <= 0) return fail("step count must be positive");
    const int n_rec = h->n_probes + h->n_mics;
    if (h->n_sources && h->n_src_cells && !src_host) return fail("source values required");
    DBuf<double> &src = h->stage_src[slot];
    DBuf<float> &rec = h->stage_rec[slot];
    const void *old_src = src.p, *old_rec = rec.p;
    if (src.alloc((size_t)std::max(1, n_steps * std::max(1, h->n_sources)))) return 1;
    if (rec.alloc((size_t)std::max(1, n_steps * std::max(1, n_rec)))) return 1;
    if (old_src != src.p || old_rec != rec.p) drop_graphs(h);          // cached graphs hold these pointers
    if (!h->stage_done[slot]) CU(cudaEventCreateWithFlags(&h->stage_done[slot], cudaEventDisableTiming));
    if (h->n_sources && src_host)
        CU(cudaMemcpyAsync(src.p, src_host, (size_t)n_steps * h->n_sources * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (sb_step_n_async(h, n_steps, src.p, rec.p)) return 1;
    if (n_rec && rec_host)
        CU(cudaMemcpyAsync(rec_host, rec.p, (size_t)n_steps * n_rec * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->stage_done[slot], h->stream));
    return 0;
}

extern "C" int sb_step_n_wait(sb_solver *h, int slot)
{
    CHECK_H(h);
    if (slot < 0 || slot > 1 || !h->stage_done[slot]) return fail("nothing was submitted in this slot");
    CU(cudaEventSynchronize(h->stage_done[slot]));
    return 0;
}

extern "C" int sb_halo_planes(sb_solver *h, float **send_lo, float **send_hi, float **recv_lo, float **recv_hi,
                              int64_t *plane_elems)
{
    CHECK_H(h);
    if (!h->set[0][0]) return fail("fields not bound");
    float *p = h->set[h->cur][0];
    if (recv_lo) *recv_lo = p;
    if (send_lo) *send_lo = p + h->plane;
    if (send_hi) *send_hi = p + (long long)h->d.nx * h->plane;
    if (recv_hi) *recv_hi = p + (long long)(h->d.nx + 1) * h->plane;
    if (plane_elems) *plane_elems = h->plane;
    return 0;
}

extern "C" int sb_set_peers(sb_solver *h, float *const lo_p_sets[2], float *const hi_p_sets[2], int lo_nx,
                            int *my_flags, int *lo_flag, int *hi_flag)
{
    CHECK_H(h);
    drop_graphs(h);
    if (!my_flags) { h->have_peers = false; return 0; }
    if (h->d.has_lower && (!lo_p_sets || !lo_p_sets[0] || !lo_p_sets[1] || !lo_flag)) return fail("lower neighbour pointers missing");
    if (h->d.has_upper && (!hi_p_sets || !hi_p_sets[0] || !hi_p_sets[1] || !hi_flag)) return fail("upper neighbour pointers missing");
    for (int q = 0; q < 2; q++) {
        h->peer_lo_set[q] = h->d.has_lower ? lo_p_sets[q] : nullptr;
        h->peer_hi_set[q] = h->d.has_upper ? hi_p_sets[q] : nullptr;
    }
    h->peer_lo_nx = lo_nx; h->my_flags = my_flags; h->sig_lo = lo_flag; h->sig_hi = hi_flag;
    CU(cudaMemsetAsync(h->d_step_global.p, 0, sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->d_err.p, 0, sizeof(int), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->have_peers = true;
    return 0;
}

extern "C" int sb_energy(sb_solver *h, double rho, double c, double dV, double *out)
{
    CHECK_H(h);
    if (!out) return fail("null argument");
    const sb_grid_desc &d = h->d;
    CU(cudaMemsetAsync(h->d_energy.p, 0, 2 * sizeof(double), h->stream));
    k_energy<<<148 * 4, 256, 0, h->stream>>>(plane0(h, h->cur, 0), plane0(h, h->cur, 1), plane0(h, h->cur, 2),
                                             plane0(h, h->cur, 3), h->have_mask ? h->mask.p + h->plane : nullptr,
                                             d.nx, d.ny, d.nz, d.pitch, h->plane, h->d_energy.p);
    h->kernels_launched++;
    double s[2];
    CU(cudaMemcpyAsync(s, h->d_energy.p, sizeof s, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out = 0.5 * s[0] / (rho * c * c) * dV + 0.5 * rho * s[1] * dV;
    return 0;
}

extern "C" int sb_reset(sb_solver *h)
{
    CHECK_H(h);
    for (int s = 0; s < 2; s++)
        for (int f = 0; f < 4; f++)
            if (h->set[s][f]) CU(cudaMemsetAsync(h->set[s][f], 0, (size_t)h->elems * 4, h->stream));
    if (h->have_ade && !h->ade_fused) {
        CU(cudaMemsetAsync(h->ade_J.p, 0, h->ade_J.n * 4, h->stream));
        CU(cudaMemsetAsync(h->ade_Jp.p, 0, h->ade_Jp.n * 4, h->stream));
    }
    if (h->have_ade && h->ade_fused) {
        for (int q = 0; q < h->ade.n_poles; q++)
            for (int b = 0; b < h->fpole[q].nbuf; b++)
                CU(cudaMemsetAsync(h->fpole[q].b[b].p, 0, h->fpole[q].b[b].n * 4, h->stream));
        h->ade_phase = 0;
    }
    for (auto *po : h->plane_ops) CU(cudaMemsetAsync(po->prev.p, 0, po->prev.n * 4, h->stream));
    CU(cudaMemsetAsync(h->d_step_global.p, 0, sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->d_err.p, 0, sizeof(int), h->stream));          // a timed-out wait is not carried into the next run
    // the neighbours' "steps done" words restart with the step counter (else the first waits of the new run pass vacuously)
    if (h->have_peers && h->my_flags) CU(cudaMemsetAsync(h->my_flags, 0, 2 * sizeof(int), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->cur = 0; h->steps_done = 0;
    return 0;
}

extern "C" int sb_set_option(sb_solver *h, int option, int value)
{
    if (!h) return fail("null handle");
    switch (option) {
        case SB_OPT_KERNEL: if (value < 0 || value > 5 || value == 3) return fail("bad kernel variant %d", value); h->opt_kernel = value; break;
        case SB_OPT_ROWS_PER_THREAD: if (value < 0 || value > 2) return fail("rows_per_thread must be 0 (auto), 1 or 2"); h->opt_rj = value; break;
        case SB_OPT_WARPS_J: if (value < 0 || value > 8) return fail("warps_j out of range"); h->opt_wj = value; break;
        case SB_OPT_WARPS_K: if (value < 1 || value > 8) return fail("warps_k out of range"); h->opt_wk = value; break;
        case SB_OPT_CHUNK_I: if (value < 0) return fail("chunk_i must be >= 0"); h->opt_chunk_i = value; break;
        case SB_OPT_USE_GRAPH: h->opt_graph = value < 0 ? -1 : (value ? 1 : 0); break;
        case SB_OPT_PROFILE: h->opt_profile = value ? 1 : 0; break;
        case SB_OPT_FUSE_K3: h->opt_fuse_k3 = value < 0 ? 0 : std::min(value, 2); break;
        case SB_OPT_PLANE_MAP: if (value < 0 || value > 2) return fail("plane_map must be 0 (auto), 1 (strips) or 2 (flat)");
                               h->opt_plane_map = value; break;
        case SB_OPT_ADE_LAYOUT: if (value < 0 || value > 3) return fail("ade_layout must be 0 (auto), 1 (compact), 2 (dense) or 3 (fused)");
                                h->opt_ade_layout = value; break;
        case SB_OPT_ADE_CHUNK_I: if (value < 0) return fail("ade_chunk_i must be >= 0"); h->opt_ade_chunk = value; break;
        case SB_OPT_ADE_WARPS: if (value < 0 || value > 8) return fail("ade_warps must be 0..8"); h->opt_ade_warps = value; break;
        case SB_OPT_ADE_OCCUPANCY: if (value != 2 && value != 3) return fail("ade_occupancy must be 2 or 3"); h->opt_ade_occ = value; break;
        case SB_OPT_RESIDENT_SPLIT: h->opt_res_split = value ? 1 : 0; break;
        case SB_OPT_RESIDENT_MIN_STEPS: h->opt_res_min_steps = std::max(1, value); break;
        default: return fail("unknown option %d", option);
    }
    drop_graphs(h);
    return 0;
}

extern "C" int sb_query(sb_solver *h, sb_stats *out)
{
    if (!h || !out) return fail("null argument");
    out->cells = (int64_t)h->d.nx * h->d.ny * h->d.nz;
    out->steps_done = h->steps_done;
    out->kernels_launched = h->kernels_launched;
    double b = 32.0;
    if (h->have_mask) b += 1.0;
    if (h->have_ade) {
        double per_cell = 0.0;
        for (int q = 0; q < h->ade.n_poles; q++) per_cell += h->ade.poles[q].is_lorentz ? 16.0 : 8.0;
        if (!h->have_mask && (h->ade_fused || h->ade_concurrent)) per_cell += 1.0;    // the mask byte carrying the ADE bits
        b += per_cell * (double)h->ade_material_cells / (double)out->cells;
    }
    out->algorithmic_bytes_per_cell = b;
    out->kernel_variant = h->last_variant;
    out->pitch = h->d.pitch;
    return 0;
}

extern "C" int sb_profile_read(sb_solver *h, double *mean_ms, double *min_ms, int *n_launches)
{
    CHECK_H(h);
    CU(cudaStreamSynchronize(h->stream));
    double sum = 0.0, mn = 1e300; int n = 0;
    for (auto &pr : h->prof) {                  // per step: a resident launch is divided by the steps it covers
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.e0, pr.e1) == cudaSuccess) { sum += ms; mn = std::min(mn, (double)ms / pr.steps); n += pr.steps; }
        cudaEventDestroy(pr.e0); cudaEventDestroy(pr.e1);
    }
    h->prof.clear();
    if (mean_ms) *mean_ms = n ? sum / n : 0.0;
    if (min_ms) *min_ms = n ? mn : 0.0;
    if (n_launches) *n_launches = n;
    return 0;
}

extern "C" int sb_tuned(sb_solver *h, int32_t shape_out[4], float *ms_out)
{
    if (!h || !shape_out) return fail("null argument");
    for (int q = 0; q < 4; q++) shape_out[q] = h->tuned[q];
    if (ms_out) *ms_out = h->tuned_key >= 0 ? h->tuned_ms : 0.f;
    return 0;
}

extern "C" int sb_synchronize(sb_solver *h)
{
    CHECK_H(h);
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    if (h->have_peers || h->resident_used) {
        int err = 0;
        CU(cudaMemcpy(&err, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (err == 2) return fail("resident kernel: timed out waiting for a neighbour box's step flag");
        if (err == 3) return fail("pipelined kernel: timed out waiting for a chunk of the previous step");
        if (err) return fail("timed out waiting for a neighbour slab's step flag (peer-to-peer halo)");
    }
    return 0;
}
