// k5_emu.cu -- TEST INFRASTRUCTURE: lockstep CPU execution of the resident kernel's phase functions
// (strata_fdtd_b200/csrc/sb_resident.cuh) so that tests/test_resident_emulation.py can check the box
// decomposition, halo indexing and deferred sponge against the oracle without a GPU.  Every "thread"
// of every "CTA" runs each phase to completion before the next phase starts, which is what the
// barriers and step flags of k5_resident guarantee on the device.  Never part of the product library.
#define SB_RESIDENT_NO_KERNEL            // host emulation only: skip the __global__ kernel (and its 12 device variants)
#include "../../strata_fdtd_b200/csrc/sb_resident.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace sb;

namespace {
template <typename T> struct Aligned {
    T *p = nullptr; size_t n = 0;
    explicit Aligned(size_t count) : n(count)
    {
        void *q = nullptr;
        if (posix_memalign(&q, 64, (count ? count : 1) * sizeof(T))) std::abort();
        p = static_cast<T *>(q);
        std::memset(p, 0, (count ? count : 1) * sizeof(T));
    }
    Aligned(const T *src, size_t count) : Aligned(count) { if (src) std::memcpy(p, src, count * sizeof(T)); }
    ~Aligned() { std::free(p); }
    Aligned(const Aligned &) = delete;
};


// host stand-in for RecvPoll: in lockstep the data is always there; a wrong tag is an indexing bug
static int recv_host(const ResParams &R, const ResBlock &B, float *sm, int s)
{
    const ResMap M(R);
    const ResHalo H(R, B);
    const unsigned tag = R.tag_base + (unsigned)s;
    int bad = 0;
    for (int idx = 0; idx < H.total; idx++) {
        const uint4 *src; int dst;
        H.item(R, B, M, s & 1, idx, src, dst);
        const uint4 a = src[0], b = src[1];
        if (a.y != tag || a.w != tag || b.y != tag || b.w != tag) bad++;
        const unsigned u[4] = {a.x, a.z, b.x, b.z};
        memcpy(sm + dst, u, 16);
    }
    return bad;
}

struct Runner {
    ResParams &R; std::vector<float *> &smem;
    template <bool GEOM, bool UNI, int NS> int run();
};

template <bool GEOM, bool UNI, int NS> int Runner::run()
{
    int bad = 0;
    const int nb = R.nbi * R.nbj;
    const ResMap M(R);
    std::vector<std::vector<int>> own(nb);                       // (slot, offset) pairs of the probes a box owns
    std::vector<std::vector<ResThread>> thr(nb);                 // the per-thread constants the kernel keeps in registers
    for (int b = 0; b < nb; b++) {
        const ResBlock B = res_block(R, b);
        for (int t = 0; t < K5_NT; t++) thr[b].push_back(res_thread(R, B, t));
        for (int t = 0; t < K5_NT; t++) res_load(R, B, smem[b], t);
        for (int t = 0; t < R.n_probes; t++) {
            const int i = R.probe_ijk[3 * t] - B.i0, j = R.probe_ijk[3 * t + 1] - B.j0, k = R.probe_ijk[3 * t + 2];
            if (i >= 0 && i < B.li_n && j >= 0 && j < B.lj_n) { own[b].push_back(t); own[b].push_back(M.p(i, j) + k); }
        }
    }
    auto phase_v = [&](int s, int pass) {
        for (int b = 0; b < nb; b++) {
            const ResBlock B = res_block(R, b);
            for (int t = 0; t < K5_NT; t++) res_phase_v<GEOM, UNI, NS>(R, B, thr[b][t], smem[b], s, pass);
        }
    };
    for (int s = 0; s < R.n_steps; s++) {
        if (s > 0) {
            if (R.split) phase_v(s, 0);
            for (int b = 0; b < nb; b++) {
                const ResBlock B = res_block(R, b);
                bad += recv_host(R, B, smem[b], s);
            }
            phase_v(s, R.split ? 1 : 2);
        } else {
            phase_v(0, 2);
        }
        for (int b = 0; b < nb; b++) {
            const ResBlock B = res_block(R, b);
            for (int t = 0; t < K5_NT; t++) res_phase_p<GEOM, UNI, NS>(R, B, thr[b][t], smem[b], s);
        }
        for (int b = 0; b < nb; b++)
            for (size_t q = 0; q + 1 < own[b].size(); q += 2)
                R.rec[(long long)s * R.n_rec + own[b][q]] = smem[b][own[b][q + 1]];
    }
    for (int b = 0; b < nb; b++) {
        const ResBlock B = res_block(R, b);
        for (int t = 0; t < K5_NT; t++) res_store(R, B, smem[b], t);
    }
    return bad;
}
}  // namespace

// fields: 8 padded buffers [(nx+2)][ny][pitch] (set 0 then set 1; p, vx, vy, vz), updated in place.
// x tables arrive with nx+2 entries (index -1 first), y tables with ny+4, z tables with pitch+4 -- the layouts
// sb_api.cu uploads.  Returns 0, 1 when the grid does not fit the given SM count / shared-memory limit, 2 when a
// received face carried the wrong step tag.
extern "C" int k5emu_run(int nx, int ny, int nz, int pitch, float **fields, int cur, int n_steps, const uint8_t *mask,
                         const float *cvx, const float *cvy, const float *cvz,
                         const float *icx, const float *icy, const float *icz,
                         int n_sponge, const float **decx, const float **decy, const float **decz, float cp,
                         int n_inline, const int *inl_ijks, const double *inl_weight,
                         const double *src_vals, int n_sources, int n_probes, const int *probe_ijk, float *rec, int n_rec,
                         int nbi, int nbj, int n_sm, long long smem_limit, int split, int *chosen)
{
    const long long plane = (long long)ny * pitch, elems = (long long)(nx + 2) * plane;
    if (nbi <= 0 && !res_choose_partition(nx, ny, nz, n_sm, smem_limit, n_probes, mask != nullptr, &nbi, &nbj)) return 1;
    if (chosen) { chosen[0] = nbi; chosen[1] = nbj; }
    std::vector<Aligned<float> *> F;
    for (int q = 0; q < 8; q++) F.push_back(new Aligned<float>(fields[q], (size_t)elems));
    Aligned<uint8_t> mk(mask, mask ? (size_t)elems : 0);
    Aligned<float> a_cvx(cvx, nx + 2), a_cvy(cvy, ny + 4), a_cvz(cvz, pitch + 4);
    Aligned<float> a_icx(icx, icx ? nx + 2 : 0), a_icy(icy, icy ? ny + 4 : 0), a_icz(icz, icz ? pitch + 4 : 0);
    std::vector<Aligned<float> *> D;
    ResParams R{};
    for (int q = 0; q < 8; q++) R.set[q / 4][q % 4] = F[q]->p + plane;
    R.cur = cur; R.n_steps = n_steps;
    R.mask = mask ? mk.p + plane : nullptr;
    R.cvx = a_cvx.p + 1; R.cvy = a_cvy.p; R.cvz = a_cvz.p;
    R.icx = icx ? a_icx.p + 1 : nullptr; R.icy = icx ? a_icy.p : nullptr; R.icz = icx ? a_icz.p : nullptr;
    R.n_sponge = n_sponge;
    for (int q = 0; q < n_sponge; q++) {
        D.push_back(new Aligned<float>(decx[q], nx + 2)); R.decx[q] = D.back()->p + 1;
        D.push_back(new Aligned<float>(decy[q], ny + 4)); R.decy[q] = D.back()->p;
        D.push_back(new Aligned<float>(decz[q], pitch + 4)); R.decz[q] = D.back()->p;
    }
    R.cp = cp; R.nx = nx; R.ny = ny; R.nz = nz; R.pitch = pitch; R.plane = plane;
    R.nbi = nbi; R.nbj = nbj; R.LI = (nx + nbi - 1) / nbi; R.LJ = (ny + nbj - 1) / nbj; R.kp = (nz + 3) / 4 * 4;
    if (R.LJ * (R.kp / 4) > K5_NT) return 1;
    R.n_inline = n_inline;
    for (int q = 0; q < n_inline; q++) {
        R.inl_i[q] = inl_ijks[4 * q]; R.inl_j[q] = inl_ijks[4 * q + 1]; R.inl_k[q] = inl_ijks[4 * q + 2];
        R.inl_src[q] = inl_ijks[4 * q + 3]; R.inl_weight[q] = inl_weight[q];
    }
    R.src_vals = src_vals; R.n_sources = n_sources;
    R.n_probes = n_probes; R.n_rec = n_rec; R.probe_ijk = probe_ijk; R.rec = rec;
    R.err_flag = nullptr; R.split = split;
    R.xch_face = (R.LI > R.LJ ? R.LI : R.LJ) * R.kp / 2;
    static unsigned epoch = 0;                                  // tags keep growing across calls, as in the library
    static Aligned<uint4> *xch = nullptr; static size_t xch_n = 0;
    const size_t need = (size_t)2 * nbi * nbj * 4 * R.xch_face;
    if (!xch || xch_n < need) { delete xch; xch = new Aligned<uint4>(need); xch_n = need; }
    R.xch = xch->p; R.tag_base = epoch; epoch += (unsigned)n_steps;
    const size_t sm_floats = (size_t)res_smem_bytes(R.LI, R.LJ, R.kp, n_probes, mask != nullptr) / 4;
    std::vector<Aligned<float> *> S;
    std::vector<float *> smem;
    for (int b = 0; b < nbi * nbj; b++) { S.push_back(new Aligned<float>(sm_floats)); smem.push_back(S.back()->p); }
    // poison the shared memory so that a read of something never loaded shows up
    for (float *p : smem) for (size_t q = 0; q < sm_floats; q++) p[q] = 1.0e30f;
    Runner runner{R, smem};
    const int bad = res_dispatch(mask != nullptr, icx == nullptr, n_sponge, runner);
    for (int q = 0; q < 8; q++) std::memcpy(fields[q], F[q]->p, (size_t)elems * sizeof(float));
    for (auto *a : F) delete a;
    for (auto *a : D) delete a;
    for (auto *a : S) delete a;
    return bad ? 2 : 0;
}
