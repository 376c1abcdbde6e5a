// iso3dfd point-update kernels for sm_100a.
//
// Equation and association order: /root/reference/src/stencils/Iso3dfdStencil.cpp:63-137
// (get_next_p) as canonicalised by the reference compiler ("-target pseudo", SURVEY.md
// Appendix A) -- this replaces the generated stencil_iso3dfd_part_1::calc_vectors()
// (emitter: /root/reference/src/compiler/lib/Cpp.cpp:377-1175) and the loop hierarchy that
// calls it (/root/reference/src/kernel/lib/context.cpp:631-1174, stencil_calc.hpp:444-860).
//
//   acc = p*c0;  r=1..R: acc += (((((p[x-r]+p[x+r])+p[y-r])+p[y+r])+p[z-r])+p[z+r]) * c_r
//   p(t+1) = ((2*p) - p(t-1)) + acc*v
//
// FP modes (see oracle/yask_oracle.c): 0 strict IEEE mul/add, 1 canonical FMA,
// 2 = FMA pattern GCC 13 emits for the reference's default build (bit-exact vs it).
#pragma once
#include <type_traits>

#include "yb_ptx.cuh"
#include "yb_iso3dfd_math.cuh"

namespace yb {

constexpr int ISO_MAX_R = 8;
constexpr int ISO_MAX_CHUNKS = 40;   // x chunks of one launch (boundary chunks + interior chunks)

struct IsoParams {
    float* out;             // &p_next[domain origin]; written in place over p(t-1)
    long long out_sx, out_sy;  // element strides of the p arrays (z stride is 1)
    const float* cur;       // &p_cur[domain origin]  (naive kernel only)
    const float* prev;      // &p_prev[domain origin] (naive kernel only): == out unless the var has extra storage slots
    const float* vel;       // &v[domain origin]      (naive kernel only)
    long long v_sx, v_sy;
    int nx, ny, nz;         // rank-domain sizes
    int x_begin, x_end;     // sub-range of x planes to compute (boundary/interior split)
    int y_begin, y_end;
    int z_begin, z_end;
    int pad_x, pad_y, pad_z;    // p arrays: alloc index of domain origin (TMA coordinates)
    int vpad_x, vpad_y, vpad_z; // v array: same
    int nty, ntz, nchunks;      // tiling of [begin,end): tiles in y,z; chunks along x
    // Every chunk: first plane computed, number of planes, sweep direction (+1 / -1).  Work units are numbered chunk-major,
    // so the chunks listed first are swept first by every CTA.  A multi-rank launch lists the chunk that starts at plane 0
    // (swept upwards) and the chunk that ends at plane nx-1 (swept DOWNWARDS) first: the planes the x neighbours need are
    // then the first R planes of those sweeps.  (A downward sweep gives the same bits: the x neighbours enter the sum as
    // p[x-r] + p[x+r], and IEEE addition commutes.)
    int cx0[ISO_MAX_CHUNKS], clen[ISO_MAX_CHUNKS], cdir[ISO_MAX_CHUNKS];
    int pol_c, pol_h, pol_pv;   // L2 eviction policy of the TMA streams: 0 normal, 1 evict_first, 2 evict_last
    int st_cs;                  // 1: results leave with streaming (evict-first) stores
    int peer_probe;             // DEBUG (option peer_probe): 1 = skip the peer stores (timing probe only: halos are NOT exchanged),
                                // 2 = CTAs start a few microseconds apart so that their boundary phases do not coincide
    // Fused halo exchange: when non-null, the first / last R computed x planes are ALSO stored into the lower /
    // upper x neighbour's halo cells (peer HBM over NVLink), indexed exactly like `out`.
    float* peer_lo;
    float* peer_hi;
    // In-kernel completion signal of the boundary planes (replaces the push + signal kernels that used to follow the
    // sweep): units [0, sig_units) are the units whose sweeps START with boundary planes.  Every consumer warp of the
    // grid arrives once on *sig_counter, right after the R-th plane of its last such unit (system-scope fence first);
    // the arrival that completes the count (sig_total) publishes sig_epoch into the neighbours' flag words with release
    // semantics and resets the counter.  The neighbours' next step, which needs exactly these planes, therefore has them
    // while this rank is still sweeping the rest (the overlap of /root/reference/src/kernel/lib/context.cpp:378-475:
    // exterior first, exchange during the interior) -- and no plane is swept twice or warmed up twice for it.
    int sig_units;
    unsigned int sig_total;
    unsigned int* sig_counter;
    unsigned long long* sig_flag_lo;
    unsigned long long* sig_flag_hi;
    unsigned long long sig_epoch;
    float c[ISO_MAX_R + 1];
};

// ---------------------------------------------------------------------------------------------
// Reference-order direct kernel: one thread per point, straight global loads (through L1/L2).
// Used for radii/shapes the tiled kernel does not cover and as the on-device cross-check of the
// tiled kernel at sizes the CPU oracle cannot reach.  Any radius 1..8.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) iso3dfd_direct_kernel(IsoParams P, int R) {
    const int z = P.z_begin + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = P.y_begin + blockIdx.y;
    const int x = P.x_begin + blockIdx.z;
    if (z >= P.z_end || y >= P.y_end || x >= P.x_end) return;
    const long long o = (long long)x * P.out_sx + (long long)y * P.out_sy + z;
    const float* pc = P.cur + o;
    const float center = pc[0];
    float acc = 0.f;
    for (int r = 1; r <= R; r++) {
        acc = iso_group<MODE>(acc, center, P.c[0], P.c[r], pc[-r * P.out_sx], pc[r * P.out_sx], pc[-r * P.out_sy],
                              pc[r * P.out_sy], pc[-r], pc[r], r == 1);
    }
    const float v = P.vel[(long long)x * P.v_sx + (long long)y * P.v_sy + z];
    P.out[o] = iso_final<MODE>(acc, center, P.prev[o], v);
}

// ---------------------------------------------------------------------------------------------
// Tiled TMA kernel ("2.5-D sweep").
//
// A CTA owns a (TY x TZ) column of the (y,z) plane and sweeps it along x (the outermost,
// largest-stride axis) over a chunk of planes.  A thread owns two vertically adjacent rows x four
// z-consecutive points for every plane of the sweep.
//
//   * x neighbours: a 2R+1 deep register queue of the thread's own quads, rotated by register moves.
//   * y,z neighbours: the current plane with halo, (TY+2R) x (TZ+2HZ) floats, staged in shared
//     memory by ONE TMA box load (cp.async.bulk.tensor.3d, box depth 1).
//   * the queue is fed by a second, halo-less TMA box R planes ahead of the current one;
//     p(t-1) and v tiles arrive the same way (evict-first: they are streamed exactly once).
//   * a ring of STAGES such stage buffers is filled by a producer warpgroup (one elected lane) running
//     STAGES-1 sweep steps ahead; full/empty mbarriers, no __syncthreads in the sweep loop.
//   * results leave as 128-bit coalesced stores straight over p(t-1) (the reference's
//     2-slot write-back, /root/reference/src/compiler/lib/Var.cpp:435-464).
//
// Persistent grid: one CTA per SM; work units (y-tile, z-tile, x-chunk) are dealt round-robin so
// that the units in flight at any time are neighbours in (y,z) and share their halos through L2.
// ---------------------------------------------------------------------------------------------
struct IsoMaps {
    CUtensorMap h;  // p_cur, box (TZ+2HZ, TY+2R, 1)
    CUtensorMap c;  // p_cur, box (TZ, TY, 1)
    CUtensorMap p;  // p_prev, box (TZ, TY, 1)
    CUtensorMap v;  // v,      box (TZ, TY, 1)
};

// Sweep position of a CTA in its flattened (unit, iteration) sequence.
struct IsoCursor {
    int unit;      // current work unit
    int it;        // iteration within unit: 0 .. n_it-1
    int n_it;      // lx_unit + 2R
    int x0, y0, z0;  // unit origin (domain coordinates); x0 = first plane computed
    int dir;         // sweep direction along x
    int stage;
    uint32_t phase;
};

template <class T>
__device__ __forceinline__ void iso_unit_setup(IsoCursor& cu, const IsoParams& P) {
    int u = cu.unit;
    const int tz = u % P.ntz; u /= P.ntz;
    const int ty = u % P.nty; u /= P.nty;
    cu.z0 = P.z_begin + tz * T::TZ;
    cu.y0 = P.y_begin + ty * T::TY;
    cu.x0 = P.cx0[u];
    cu.dir = P.cdir[u];
    cu.n_it = P.clen[u] + 2 * T::R;
    cu.it = 0;
}

// ---------------------------------------------------------------------------------------------
// Thread layout ("row pairs").
//
// ncu on the first version of this kernel (one row x 4 z per thread, sweep loop unrolled 2R+1 times;
// profiles/r1_iso3dfd.md) showed the sweep limited by shared-memory wavefronts (23 LDS.128 per 4 points,
// every y neighbour loaded once per point) and by instruction-cache misses of the unrolled body.  Hence:
//   * a thread owns TWO vertically adjacent rows x 4 z (8 points).  The y window of the pair is
//     18 rows, each loaded ONCE and used by both rows: 16+8 neighbour loads per 8 points
//     instead of 2 x 20 -- 40 % fewer shared-memory wavefronts per point;
//   * the x queue is rotated by register moves (loop NOT unrolled 2R+1 times), so the whole
//     sweep body fits the instruction cache.
// ---------------------------------------------------------------------------------------------
template <int R_, int TYP_, int TZQ_, int STAGES_>
struct IsoTile2 {
    static constexpr int R = R_, TYP = TYP_, TZQ = TZQ_, STAGES = STAGES_;
    static constexpr int TY = 2 * TYP;
    static constexpr int TZ = 4 * TZQ;
    static constexpr int HZ = (R + 3) / 4 * 4;
    static constexpr int ZQ = HZ / 4;
    static constexpr int HP = TZ + 2 * HZ;
    static constexpr int HROWS = TY + 2 * R;
    static constexpr int THREADS = TYP * TZQ;
    static constexpr int NWARPS = THREADS / 32;
    static constexpr int QN = 2 * R + 1;
    static constexpr uint32_t H_BYTES = HROWS * HP * 4;
    static constexpr uint32_t C_BYTES = TY * TZ * 4;
    static constexpr uint32_t H_OFF = 0;
    static constexpr uint32_t C_OFF = (H_BYTES + 127) / 128 * 128;
    static constexpr uint32_t P_OFF = C_OFF + C_BYTES;
    static constexpr uint32_t V_OFF = P_OFF + C_BYTES;
    static constexpr uint32_t STAGE_BYTES = V_OFF + C_BYTES;
    static constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr uint32_t SMEM_BYTES = BAR_OFF + 2 * STAGES * 8 + 128;
    static_assert(THREADS % 32 == 0, "whole warps");
    static_assert(C_BYTES % 128 == 0, "TMA destination alignment");
};

// Boundary-phase completion (see IsoParams::sig_*): called once per consumer warp, by all of its lanes.
static __device__ __noinline__ void iso_boundary_arrive(const IsoParams& P, int lane) {
    __threadfence_system();          // this lane's peer stores are ordered before what follows, at system scope
    __syncwarp();
    if (lane == 0) {
        const unsigned int prev = atomicAdd(P.sig_counter, 1u);
        if (prev + 1u == P.sig_total) {            // every warp of the grid has finished its boundary units
            __threadfence_system();
            atomicExch(P.sig_counter, 0u);         // ready for the next launch (stream-ordered after this one)
            if (P.sig_flag_lo) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.sig_flag_lo), "l"(P.sig_epoch) : "memory");
            if (P.sig_flag_hi) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(P.sig_flag_hi), "l"(P.sig_epoch) : "memory");
        }
    }
}

__device__ __forceinline__ void f4_to_arr(const float4& v, float* a) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }

// PW = 1: a dedicated producer warp (warp NWARPS) issues the TMA loads, so no compute warp carries the
//         issue code (with 2 warps per scheduler a straggling warp 0 delays every stage hand-over).
// U  = 2: the sweep loop handles two planes per trip and rotates the x queues by two entries every second
//         plane (half the register moves of U = 1) at the price of one extra queue entry per row.
template <class T, int MODE, int PW = 0, int U = 1>
__global__ void __launch_bounds__(T::THREADS + 128 * PW, 1)
iso3dfd_tma2_kernel(const __grid_constant__ IsoMaps M, const __grid_constant__ IsoParams P) {
    constexpr int R = T::R, QN = T::QN, ZQ = T::ZQ;
    constexpr int QL = QN + (U - 1);          // queue entries kept per row
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbase + T::BAR_OFF);
    uint64_t* empty_bar = full_bar + T::STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int nunits = P.nty * P.ntz * P.nchunks;
    const int prod_tid = PW ? T::THREADS : 0;

    if (tid == 0) {
        tma_prefetch_desc(&M.h); tma_prefetch_desc(&M.c); tma_prefetch_desc(&M.p); tma_prefetch_desc(&M.v);
        for (int s = 0; s < T::STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], T::NWARPS); }
        fence_barrier_init();
    }
    __syncthreads();

    // ---- producer -----------------------------------------------------------------------------------
    IsoCursor pr;
    pr.unit = blockIdx.x; pr.stage = 0; pr.phase = 0; pr.it = 0; pr.n_it = 0; pr.x0 = pr.y0 = pr.z0 = 0; pr.dir = 1;
    const uint64_t pol_pv = l2_policy(P.pol_pv), pol_c = l2_policy(P.pol_c), pol_h = l2_policy(P.pol_h);
    bool pr_live = (tid == prod_tid) && (pr.unit < nunits);
    if (pr_live) iso_unit_setup<T>(pr, P);

    auto produce_one = [&]() {
        mbar_wait(&empty_bar[pr.stage], pr.phase ^ 1u);
        uint8_t* st = sbase + pr.stage * T::STAGE_BYTES;
        uint64_t* fb = &full_bar[pr.stage];
        const bool compute = pr.it >= 2 * R;
        mbar_arrive_expect_tx(fb, compute ? (T::H_BYTES + 3 * T::C_BYTES) : T::C_BYTES);
        const int cz = P.pad_z + pr.z0, cy = P.pad_y + pr.y0;
        tma_load_3d_hint(st + T::C_OFF, &M.c, fb, cz, cy, P.pad_x + pr.x0 + pr.dir * (pr.it - R), pol_c);
        if (compute) {
            const int xo = pr.x0 + pr.dir * (pr.it - 2 * R);
            tma_load_3d_hint(st + T::H_OFF, &M.h, fb, cz - T::HZ, cy - R, P.pad_x + xo, pol_h);
            tma_load_3d_hint(st + T::P_OFF, &M.p, fb, cz, cy, P.pad_x + xo, pol_pv);
            tma_load_3d_hint(st + T::V_OFF, &M.v, fb, P.vpad_z + pr.z0, P.vpad_y + pr.y0, P.vpad_x + xo, pol_pv);
        }
        if (++pr.stage == T::STAGES) { pr.stage = 0; pr.phase ^= 1u; }
        if (++pr.it == pr.n_it) {
            pr.unit += gridDim.x;
            if (pr.unit < nunits) iso_unit_setup<T>(pr, P); else pr_live = false;
        }
    };
    if (PW) {
        // Warp-specialised register budget (setmaxnreg works on whole warpgroups): the producer warpgroup
        // hands its registers back, the two consumer warpgroups take them.
        if (tid >= T::THREADS) {          // producer warpgroup: one lane streams every load of this CTA, then exits
            asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
            while (pr_live) produce_one();
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
    } else if (tid == 0) {
        for (int k = 0; k < T::STAGES - 1 && pr_live; k++) produce_one();
    }

    if (P.peer_probe == 2) __nanosleep((blockIdx.x & 7u) * 2000u);
    // ---- consumer ------------------------------------------------------------------------------------
    const int rp = tid / T::TZQ;           // row pair 0..TYP-1  (rows 2rp, 2rp+1)
    const int quad = tid % T::TZQ;
    const uint32_t h_own = ((2 * rp + R) * T::HP + T::HZ + 4 * quad) * 4;  // row a centre in H
    const uint32_t c_own = (2 * rp * T::TZ + 4 * quad) * 4;               // row a in C/P/V tiles

    IsoCursor cu;
    cu.stage = 0; cu.phase = 0;
    float4 qa[QL], qb[QL];
#pragma unroll
    for (int k = 0; k < QL; k++) { qa[k] = make_float4(0.f, 0.f, 0.f, 0.f); qb[k] = qa[k]; }

    bool sig_pending = P.sig_counter != nullptr;
    for (cu.unit = blockIdx.x; cu.unit < nunits; cu.unit += gridDim.x) {
        if (sig_pending && cu.unit >= P.sig_units) { iso_boundary_arrive(P, lane); sig_pending = false; }   // (a CTA without boundary-first units)
        iso_unit_setup<T>(cu, P);
        const int ya = cu.y0 + 2 * rp;
        const int zq = cu.z0 + 4 * quad;
        const int nz_ok = max(0, min(4, P.z_end - zq));
        const int nva = (ya < P.y_end) ? nz_ok : 0;
        const int nvb = (ya + 1 < P.y_end) ? nz_ok : 0;
        float* out_a = P.out + (long long)ya * P.out_sy + zq + (long long)cu.x0 * P.out_sx;       // plane x0; step s -> + dir*s planes
        const long long out_step = (long long)cu.dir * P.out_sx;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_a) & 15) == 0);
        // this warp's boundary planes are complete after plane R-1 of its last boundary-first unit
        const bool sig_here = sig_pending && cu.unit < P.sig_units && cu.unit + int(gridDim.x) >= P.sig_units;

        // One sweep step.  B = index of the oldest live queue entry after the push (window = q[B .. B+2R]).
        auto step = [&](auto Bc, const int it) {
            constexpr int B = decltype(Bc)::value;
            if (!PW && tid == 0 && pr_live) produce_one();

            const uint8_t* st = sbase + cu.stage * T::STAGE_BYTES;
            mbar_wait(&full_bar[cu.stage], cu.phase);

            // push plane x0 - R + it into the x queues
            if (B == 0) {
#pragma unroll
                for (int k = 0; k + U < QL; k++) { qa[k] = qa[k + U]; qb[k] = qb[k + U]; }
            }
            qa[B + QN - 1] = *reinterpret_cast<const float4*>(st + T::C_OFF + c_own);
            qb[B + QN - 1] = *reinterpret_cast<const float4*>(st + T::C_OFF + c_own + T::TZ * 4);

            if (MODE == 3) {
                // DEBUG (mem_probe=1): memory-system ceiling probe -- same TMA traffic and stores, no stencil math.
                if (it >= 2 * R) {
                    const float4 pva = *reinterpret_cast<const float4*>(st + T::P_OFF + c_own);
                    const float4 vva = *reinterpret_cast<const float4*>(st + T::V_OFF + c_own);
                    const float4 pvb = *reinterpret_cast<const float4*>(st + T::P_OFF + c_own + T::TZ * 4);
                    const float4 vvb = *reinterpret_cast<const float4*>(st + T::V_OFF + c_own + T::TZ * 4);
                    const float4 ha = *reinterpret_cast<const float4*>(st + T::H_OFF + h_own);
                    float4 ra = make_float4(pva.x + vva.x * qa[B + R].x, pva.y + vva.y * ha.y, pva.z + vva.z, pva.w + vva.w);
                    float4 rb = make_float4(pvb.x + vvb.x * qb[B + R].x, pvb.y + vvb.y, pvb.z + vvb.z, pvb.w + vvb.w);
                    float* oa = out_a + (long long)(it - 2 * R) * out_step;
                    if (vec_ok && nva == 4) stg128(oa, ra);
                    if (vec_ok && nvb == 4) stg128(oa + P.out_sy, rb);
                }
            } else if (it >= 2 * R) {
                const float* hp = reinterpret_cast<const float*>(st + T::H_OFF + h_own);
                float pa[4], pb[4];
                f4_to_arr(qa[B + R], pa);
                f4_to_arr(qb[B + R], pb);
                // z windows of both rows (centre quads come from the queue)
                float za[4 * (2 * ZQ + 1)], zb[4 * (2 * ZQ + 1)];
#pragma unroll
                for (int k = -ZQ; k <= ZQ; k++) {
                    if (k == 0) { f4_to_arr(qa[B + R], &za[4 * ZQ]); f4_to_arr(qb[B + R], &zb[4 * ZQ]); continue; }
                    f4_to_arr(*reinterpret_cast<const float4*>(hp + 4 * k), &za[4 * (k + ZQ)]);
                    f4_to_arr(*reinterpret_cast<const float4*>(hp + T::HP + 4 * k), &zb[4 * (k + ZQ)]);
                }
                float acca[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
                // y window: w(k) = hp + (k - R) * HP ; row a centre = w(R), row b centre = w(R+1)
                float wlo_prev[4], whi_prev[4];   // w(R - (r-1)) and w(R+1 + (r-1)) from the previous radius
                f4_to_arr(qa[B + R], wlo_prev);   // r=1: row b's y-1 neighbour is row a's centre
                f4_to_arr(qb[B + R], whi_prev);   //      row a's y+1 neighbour is row b's centre
#pragma unroll
                for (int r = 1; r <= R; r++) {
                    float wlo[4], whi[4], xm[4], xp[4];
                    f4_to_arr(*reinterpret_cast<const float4*>(hp - r * T::HP), wlo);        // row a - r
                    f4_to_arr(*reinterpret_cast<const float4*>(hp + (r + 1) * T::HP), whi);  // row b + r
                    f4_to_arr(qa[B + R - r], xm); f4_to_arr(qa[B + R + r], xp);
#pragma unroll
                    for (int i = 0; i < 4; i++)   // row a: y-r = wlo, y+r = w(R+r) = previous whi
                        acca[i] = iso_group<MODE>(acca[i], pa[i], P.c[0], P.c[r], xm[i], xp[i], wlo[i], whi_prev[i],
                                                  za[4 * ZQ + i - r], za[4 * ZQ + i + r], r == 1);
                    f4_to_arr(qb[B + R - r], xm); f4_to_arr(qb[B + R + r], xp);
#pragma unroll
                    for (int i = 0; i < 4; i++)   // row b: y-r = w(R+1-r) = previous wlo, y+r = whi
                        accb[i] = iso_group<MODE>(accb[i], pb[i], P.c[0], P.c[r], xm[i], xp[i], wlo_prev[i], whi[i],
                                                  zb[4 * ZQ + i - r], zb[4 * ZQ + i + r], r == 1);
#pragma unroll
                    for (int i = 0; i < 4; i++) { wlo_prev[i] = wlo[i]; whi_prev[i] = whi[i]; }
                }
                const float4 pva = *reinterpret_cast<const float4*>(st + T::P_OFF + c_own);
                const float4 vva = *reinterpret_cast<const float4*>(st + T::V_OFF + c_own);
                const float4 pvb = *reinterpret_cast<const float4*>(st + T::P_OFF + c_own + T::TZ * 4);
                const float4 vvb = *reinterpret_cast<const float4*>(st + T::V_OFF + c_own + T::TZ * 4);
                float4 ra, rb;
                ra.x = iso_final<MODE>(acca[0], pa[0], pva.x, vva.x); ra.y = iso_final<MODE>(acca[1], pa[1], pva.y, vva.y);
                ra.z = iso_final<MODE>(acca[2], pa[2], pva.z, vva.z); ra.w = iso_final<MODE>(acca[3], pa[3], pva.w, vva.w);
                rb.x = iso_final<MODE>(accb[0], pb[0], pvb.x, vvb.x); rb.y = iso_final<MODE>(accb[1], pb[1], pvb.y, vvb.y);
                rb.z = iso_final<MODE>(accb[2], pb[2], pvb.z, vvb.z); rb.w = iso_final<MODE>(accb[3], pb[3], pvb.w, vvb.w);
                float* oa = out_a + (long long)(it - 2 * R) * out_step;
                float* ob = oa + P.out_sy;
                if (vec_ok && nva == 4) { if (P.st_cs) stg128_cs(oa, ra); else stg128(oa, ra); }
                else if (nva > 0) { oa[0] = ra.x; if (nva > 1) oa[1] = ra.y; if (nva > 2) oa[2] = ra.z; if (nva > 3) oa[3] = ra.w; }
                if (vec_ok && nvb == 4) { if (P.st_cs) stg128_cs(ob, rb); else stg128(ob, rb); }
                else if (nvb > 0) { ob[0] = rb.x; if (nvb > 1) ob[1] = rb.y; if (nvb > 2) ob[2] = rb.z; if (nvb > 3) ob[3] = rb.w; }
                // fused halo exchange: boundary planes also go straight into the x neighbours' halo cells
                const int xo = cu.x0 + cu.dir * (it - 2 * R);
                float* peer = (P.peer_lo != nullptr && xo < R) ? P.peer_lo : ((P.peer_hi != nullptr && xo >= P.nx - R) ? P.peer_hi : nullptr);
                if (peer != nullptr && P.peer_probe != 1) {
                    float* qa_ = peer + (oa - P.out);
                    float* qb_ = qa_ + P.out_sy;
                    if (vec_ok && nva == 4) stg128(qa_, ra);
                    else if (nva > 0) { qa_[0] = ra.x; if (nva > 1) qa_[1] = ra.y; if (nva > 2) qa_[2] = ra.z; if (nva > 3) qa_[3] = ra.w; }
                    if (vec_ok && nvb == 4) stg128(qb_, rb);
                    else if (nvb > 0) { qb_[0] = rb.x; if (nvb > 1) qb_[1] = rb.y; if (nvb > 2) qb_[2] = rb.z; if (nvb > 3) qb_[3] = rb.w; }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[cu.stage]);
            if (++cu.stage == T::STAGES) { cu.stage = 0; cu.phase ^= 1u; }
            if (sig_here && it == 3 * R - 1) { iso_boundary_arrive(P, lane); sig_pending = false; }
        };

        if (U == 1) {
#pragma unroll 1
            for (int it = 0; it < cu.n_it; it++) step(std::integral_constant<int, 0>{}, it);
        } else {
            // trip = [rotate by 2, push] then [push]: the second plane lands one entry further along
#pragma unroll 1
            for (int it = 0; it < cu.n_it; it += 2) {
                step(std::integral_constant<int, 0>{}, it);
                if (it + 1 < cu.n_it) step(std::integral_constant<int, U - 1>{}, it + 1);
            }
        }
    }
    if (sig_pending) iso_boundary_arrive(P, lane);
}

}  // namespace yb
