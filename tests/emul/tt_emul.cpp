// TEST INFRASTRUCTURE ONLY -- CTA emulator for the iso3dfd temporal-tile kernel (yask_b200/csrc/yb_iso3dfd_tt.cuh).
//
// The product header is compiled here by g++ with YB_TT_HOST_EMUL: the very same tt_sweep / tt_step1 / tt_step2 /
// tt_loads code the GPU runs, on a back end that executes the threads of a CTA in a loop, copies TMA boxes on the host
// (zero fill outside the array, as the hardware does) and models the full barriers (phase parity, transaction
// completion).  It verifies index arithmetic, ring-slot reuse and barrier phases bit for bit against the oracle on a
// machine without a GPU; it is NOT a product path (nothing under yask_b200/ links it) and it is not timed.
//
// Two completion models per run, chosen by `lazy`:
//   eager: a box lands in shared memory the moment it is issued  (the earliest the hardware may write: a load that is
//          issued while its slot is still being read corrupts the result)
//   lazy : a box lands when its barrier is waited on              (the latest: a read that is not covered by the right
//          wait sees poisoned shared memory)
// Shared memory starts poisoned with NaNs.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define YB_TT_HOST_EMUL 1
#define YB_TT_SMEM_HOOKS 1
#define YB_DEVFN static inline
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }

#include "../../yask_b200/csrc/yb_iso3dfd_tt.cuh"

using namespace yb;

// ---- race detector --------------------------------------------------------------------------------------------------------
// Every generic-proxy access to shared memory reports here.  Between two CTA barriers ("epoch") a 16-byte vector that one thread
// writes must not be read or written by another thread: the emulator runs the threads one after the other, so such a race
// would go unnoticed in the data.  (TMA writes are ordered by the full barriers and are covered by the eager / lazy models.)
namespace {
struct RaceState {
    const uint8_t* base = nullptr;
    size_t nvec = 0;
    std::vector<int> rd_epoch, rd_tid, wr_epoch, wr_tid;     // per 16-byte vector: last reader / writer and when
    int epoch = 0, tid = -1, races = 0;
    // bank-conflict model (optional): the addresses of every thread's shared-memory vector accesses of the current phase, in
    // program order; a phase = one threads() call of the back end
    bool model_banks = false;
    std::vector<std::vector<long>> seq;      // [tid] -> 16-byte vector indices
    unsigned long long n_inst = 0, n_wave = 0;   // warp-level vector accesses and their wavefronts (4 per conflict-free access)
    void reset(const uint8_t* b, size_t bytes) {
        base = b; nvec = bytes / 16; epoch = 0; tid = -1; races = 0;
        rd_epoch.assign(nvec, -1); rd_tid.assign(nvec, -1); wr_epoch.assign(nvec, -1); wr_tid.assign(nvec, -1);
    }
    long idx(const float* p) const {
        const long off = reinterpret_cast<const uint8_t*>(p) - base;
        return (off >= 0 && size_t(off) < nvec * 16) ? off / 16 : -1;
    }
} g_race;
}  // namespace
namespace yb {
void tt_hook_smem_read(const float* p) {
    const long i = g_race.idx(p);
    if (i < 0) { g_race.races++; return; }                                   // read outside the CTA's shared memory
    if (g_race.wr_epoch[i] == g_race.epoch && g_race.wr_tid[i] != g_race.tid) g_race.races++;    // read-after-write without a barrier
    if (g_race.model_banks && g_race.tid >= 0) g_race.seq[g_race.tid].push_back(i);
    if (g_race.rd_epoch[i] != g_race.epoch) { g_race.rd_epoch[i] = g_race.epoch; g_race.rd_tid[i] = g_race.tid; }
    else if (g_race.rd_tid[i] != g_race.tid) g_race.rd_tid[i] = -2;          // several readers in this epoch
}
void tt_hook_smem_write(const float* p) {
    const long i = g_race.idx(p);
    if (i < 0) { g_race.races++; return; }
    if (g_race.rd_epoch[i] == g_race.epoch && g_race.rd_tid[i] != g_race.tid) g_race.races++;    // write-after-read without a barrier
    if (g_race.wr_epoch[i] == g_race.epoch && g_race.wr_tid[i] != g_race.tid) g_race.races++;    // write-after-write
    g_race.wr_epoch[i] = g_race.epoch; g_race.wr_tid[i] = g_race.tid;
    if (g_race.model_banks && g_race.tid >= 0) g_race.seq[g_race.tid].push_back(i);
}
}  // namespace yb

namespace {

// a padded 3-D array as a tensor map sees it: dims (z, y, x), z unit stride
struct Tensor {
    const float* base;
    long long nz, ny, nx, sy, sx;
};

struct Pending { uint32_t off; const Tensor* t; int bz, by, z, y, x; };

template <class T>
struct Emul {
    std::vector<uint8_t> smem;
    Tensor pin, prev, v;
    bool lazy = false;
    int error = 0;
    struct Bar { unsigned phases = 0; bool armed = false; std::vector<Pending> pend; } bars[T::NS];
    std::vector<TTVec4> vregs;
    std::vector<TTThread<T>> ths;
    TTThread<T>& thread(int tid) { return ths[tid]; }

    Emul() : smem(T::SMEM_BYTES), vregs(size_t(T::THREADS) * T::S2_ROUNDS), ths(T::THREADS) {
        const uint32_t nan = 0x7fc00000u;
        for (size_t i = 0; i + 4 <= smem.size(); i += 4) memcpy(&smem[i], &nan, 4);
    }
    void copy_box(const Pending& p) {
        float* dst = reinterpret_cast<float*>(smem.data() + p.off);
        for (int y = 0; y < p.by; y++)
            for (int z = 0; z < p.bz; z++) {
                const long long gz = p.z + z, gy = p.y + y, gx = p.x;
                const bool in = gz >= 0 && gz < p.t->nz && gy >= 0 && gy < p.t->ny && gx >= 0 && gx < p.t->nx;
                dst[y * p.bz + z] = in ? p.t->base[gx * p.t->sx + gy * p.t->sy + gz] : 0.f;
            }
    }
    template <class F> void threads(F f) {
        if (g_race.model_banks) g_race.seq.assign(T::THREADS, std::vector<long>());
        for (int t = 0; t < T::THREADS; t++) { g_race.tid = t; f(t); }
        g_race.tid = -1;
        if (!g_race.model_banks) return;
        // 128-bit accesses are served per quarter-warp: 8 lanes x 16 B cover the 32 banks once; two different vectors in the same
        // 16-byte bank group (vector index mod 8) cost one more wavefront, the same vector twice is a broadcast
        for (int w = 0; w < T::THREADS; w += 32) {
            size_t ni = 0;
            for (int l = 0; l < 32 && w + l < T::THREADS; l++) ni = std::max(ni, g_race.seq[w + l].size());
            for (size_t i = 0; i < ni; i++) {
                bool any = false;
                for (int qw = 0; qw < 4; qw++) {
                    long seen[8][8]; int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int l = qw * 8; l < qw * 8 + 8 && w + l < T::THREADS; l++) {
                        if (i >= g_race.seq[w + l].size()) continue;
                        const long v = g_race.seq[w + l][i];
                        const int g = int(v & 7);
                        bool dup = false;
                        for (int c = 0; c < cnt[g]; c++) dup = dup || seen[g][c] == v;
                        if (!dup) seen[g][cnt[g]++] = v;
                    }
                    int mx = 0;
                    for (int g = 0; g < 8; g++) mx = std::max(mx, cnt[g]);
                    g_race.n_wave += mx;
                    any = any || mx > 0;
                }
                g_race.n_inst += any;
            }
        }
    }
    template <class F> void once(F f) { f(); }
    void barrier() { g_race.epoch++; }
    TTVec4* vreg(int tid) { return &vregs[size_t(tid) * T::S2_ROUNDS]; }
    void issue(const TTLoads& L, int b) {
        Bar& br = bars[b];
        if (br.armed) { error = 1; return; }          // a phase is re-armed before its previous use was waited on
        uint32_t bytes = T::P_BYTES;
        std::vector<Pending> ps;
        ps.push_back(Pending{L.p_off, &pin, T::IZ, T::IY, L.pz, L.py, L.px});
        if (L.step1) {
            ps.push_back(Pending{L.pv_off, &prev, T::S1Z, T::S1Y, L.sz, L.sy, L.sx});
            ps.push_back(Pending{L.v_off, &v, T::S1Z, T::S1Y, L.vz, L.vy, L.vx});
            bytes += 2 * T::S_BYTES;
        }
        if (bytes != L.bytes) { error = 2; return; }   // expect_tx would never be met (hang) or be exceeded
        for (auto& p : ps) {
            if (p.off % 128 != 0 || (p.z * 4) % 16 != 0) { error = 3; return; }   // TMA destination / box start alignment
            if (!lazy) copy_box(p);
        }
        br.armed = true;
        if (lazy) br.pend = ps;
    }
    void wait_full(int b, uint32_t parity) {
        Bar& br = bars[b];
        if ((br.phases & 1u) == parity) {              // the phase waited for has not completed yet: it must be in flight
            if (!br.armed) { error = 4; return; }      // nothing in flight: the GPU would hang here
            for (auto& p : br.pend) copy_box(p);
            br.pend.clear();
            br.armed = false;
            br.phases++;
        }
        // else: that phase completed earlier -- legal only if nothing newer is in flight on this barrier that we
        // should have waited for; the data check catches a stale read
    }
};

template <class T, int MODE>
int run(const float* pprev, const float* pcur, const float* vel, float* out1, float* out2, const int* n, const int* ppad, const int* vpad,
        const float* coef, int grid, int nchunks, int lazy) {
    const long long pz = n[2] + 2 * ppad[2], py = n[1] + 2 * ppad[1], px = n[0] + 2 * ppad[0];
    const long long vz = n[2] + 2 * vpad[2], vy = n[1] + 2 * vpad[1], vx = n[0] + 2 * vpad[0];
    TTParams P{};
    P.p_sy = pz; P.p_sx = pz * py; P.v_sy = vz; P.v_sx = vz * vy;
    const long long porg = ppad[0] * P.p_sx + ppad[1] * P.p_sy + ppad[2];
    P.out1 = out1 + porg; P.out2 = out2 + porg;
    P.vel = vel + vpad[0] * P.v_sx + vpad[1] * P.v_sy + vpad[2];
    P.nx = n[0]; P.ny = n[1]; P.nz = n[2];
    P.pad_x = ppad[0]; P.pad_y = ppad[1]; P.pad_z = ppad[2];
    P.vpad_x = vpad[0]; P.vpad_y = vpad[1]; P.vpad_z = vpad[2];
    for (int r = 0; r <= T::R; r++) P.c[r] = coef[r];
    P.nty = (P.ny + T::TY - 1) / T::TY;
    P.ntz = (P.nz + T::TZ - 1) / T::TZ;
    if (nchunks < 1 || nchunks > TT_MAX_CHUNKS || nchunks > P.nx) return -1;
    P.nchunks = nchunks;
    for (int k = 0; k < nchunks; k++) {
        const long long x0 = (long long)P.nx * k / nchunks, x1 = (long long)P.nx * (k + 1) / nchunks;
        P.cx0[k] = int(x0); P.clen[k] = int(x1 - x0);
    }
    for (int blk = 0; blk < grid; blk++) {
        Emul<T> be;
        be.lazy = lazy != 0;
        be.pin = Tensor{pcur, pz, py, px, P.p_sy, P.p_sx};
        be.prev = Tensor{pprev, pz, py, px, P.p_sy, P.p_sx};
        be.v = Tensor{vel, vz, vy, vx, P.v_sy, P.v_sx};
        g_race.reset(be.smem.data(), be.smem.size());
        tt_sweep<T, MODE>(be, be.smem.data(), P, blk, grid);
        if (be.error) return be.error;
        if (g_race.races) return 6;                        // data race between threads inside one barrier interval
        for (auto& b : be.bars) if (b.armed) return 5;     // loads still in flight when the CTA exits
    }
    return 0;
}

template <class T>
int run_mode(int mode, const float* a, const float* b, const float* c, float* d, float* e, const int* n, const int* pp, const int* vp, const float* coef,
             int grid, int nchunks, int lazy) {
    switch (mode) {
        case 0: return run<T, 0>(a, b, c, d, e, n, pp, vp, coef, grid, nchunks, lazy);
        case 1: return run<T, 1>(a, b, c, d, e, n, pp, vp, coef, grid, nchunks, lazy);
        default: return run<T, 2>(a, b, c, d, e, n, pp, vp, coef, grid, nchunks, lazy);
    }
}

}  // namespace

// Bank-conflict model: switch recording on / off and read (warp-level vector accesses, wavefronts) since the last reset.
extern "C" void tt_emul_bank_model(int on) { g_race.model_banks = on != 0; g_race.n_inst = g_race.n_wave = 0; }
extern "C" void tt_emul_bank_stats(unsigned long long* inst, unsigned long long* waves) { *inst = g_race.n_inst; *waves = g_race.n_wave; }

// p arrays: (nx+2*ppad[0], ny+2*ppad[1], nz+2*ppad[2]) floats, v: same with vpad.  out1 / out2 must arrive holding copies of
// pprev / pcur (their halo cells are what the engine's begin_run() replicates); the domain parts are overwritten with
// p(t+1) / p(t+2).  `variant`: 0 = the shipped tile of that radius, 1 = a small tile (more tiles and rounds per test); 2, 3 = the same two
// with the x neighbours in register queues (TTile XQ = 1); 4, 5 = the shipped tile with 512 threads, XQ = 0 / 1.
// Returns 0, or a protocol error code (see Emul).
extern "C" int tt_emul_run(int radius, int variant, int mode, const float* pprev, const float* pcur, const float* vel, float* out1, float* out2,
                           const int* n, const int* ppad, const int* vpad, const float* coef, int grid, int nchunks, int lazy) {
    if (radius == 1 && variant == 0) return run_mode<TTile<1, 16, 128, 3, 256>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 2 && variant == 0) return run_mode<TTile<2, 16, 128, 2, 256>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 1 && variant == 1) return run_mode<TTile<1, 4, 16, 2, 32>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 2 && variant == 1) return run_mode<TTile<2, 4, 16, 1, 32>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 1 && variant == 2) return run_mode<TTile<1, 16, 128, 3, 256, 1>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 2 && variant == 2) return run_mode<TTile<2, 16, 128, 3, 256, 1>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 1 && variant == 3) return run_mode<TTile<1, 4, 16, 2, 32, 1>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 2 && variant == 3) return run_mode<TTile<2, 4, 16, 1, 32, 1>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 1 && variant == 4) return run_mode<TTile<1, 16, 128, 3, 512, 0>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 2 && variant == 4) return run_mode<TTile<2, 16, 128, 2, 512, 0>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 1 && variant == 5) return run_mode<TTile<1, 16, 128, 3, 512, 1>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    if (radius == 2 && variant == 5) return run_mode<TTile<2, 16, 128, 3, 512, 1>>(mode, pprev, pcur, vel, out1, out2, n, ppad, vpad, coef, grid, nchunks, lazy);
    return -2;
}
