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

namespace yb {

constexpr int ISO_MAX_R = 8;

struct IsoParams {
    float* out;             // &p_next[domain origin]; written in place over p(t-1)
    long long out_sx, out_sy;  // element strides of the p arrays (z stride is 1)
    const float* cur;       // &p_cur[domain origin]  (naive kernel only)
    const float* vel;       // &v[domain origin]      (naive kernel only)
    long long v_sx, v_sy;
    int nx, ny, nz;         // rank-domain sizes
    int x_begin, x_end;     // sub-range of x planes to compute (boundary/interior split)
    int y_begin, y_end;
    int z_begin, z_end;
    int pad_x, pad_y, pad_z;    // p arrays: alloc index of domain origin (TMA coordinates)
    int vpad_x, vpad_y, vpad_z; // v array: same
    int nty, ntz, nchunks, lx;  // tiling of [begin,end): tiles in y,z; chunks of lx planes in x
    int pol_c, pol_h, pol_pv;   // L2 eviction policy of the TMA streams: 0 normal, 1 evict_first, 2 evict_last
    int st_cs;                  // 1: results leave with streaming (evict-first) stores
    // Fused halo exchange: when non-null, the first / last R computed x planes are ALSO stored into the lower /
    // upper x neighbour's halo cells (peer HBM over NVLink), indexed exactly like `out`.
    float* peer_lo;
    float* peer_hi;
    float c[ISO_MAX_R + 1];
};

template <int MODE>
__device__ __forceinline__ float iso_group(float acc, float pc, float c0, float cr, float xm, float xp, float ym, float yp,
                                           float zm, float zp, bool first) {
    float s = __fadd_rn(xm, xp);
    s = __fadd_rn(s, ym);
    s = __fadd_rn(s, yp);
    s = __fadd_rn(s, zm);
    s = __fadd_rn(s, zp);
    if (MODE == 0) {
        if (first) acc = __fmul_rn(pc, c0);
        return __fadd_rn(acc, __fmul_rn(s, cr));
    } else if (MODE == 1) {
        if (first) acc = __fmul_rn(pc, c0);
        return __fmaf_rn(s, cr, acc);
    } else {
        if (first) return __fmaf_rn(pc, c0, __fmul_rn(s, cr));
        return __fmaf_rn(s, cr, acc);
    }
}

template <int MODE>
__device__ __forceinline__ float iso_final(float acc, float pc, float prev, float v) {
    // 2*p is exact, so fma(2,p,-prev) == round((2*p) - prev): one instruction, same bits.
    float lhs = __fmaf_rn(2.0f, pc, -prev);
    if (MODE == 0) return __fadd_rn(lhs, __fmul_rn(acc, v));
    return __fmaf_rn(acc, v, lhs);
}

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
    P.out[o] = iso_final<MODE>(acc, center, P.out[o], v);
}

// ---------------------------------------------------------------------------------------------
// Tiled TMA kernel ("2.5-D sweep").
//
// A CTA owns a (TY x TZ) column of the (y,z) plane and sweeps it along x (the outermost,
// largest-stride axis) over a chunk of `lx` planes.  Thread (row j, quad l) computes the four
// z-consecutive points z0+4l..+3 of row j for every plane of the sweep.
//
//   * x neighbours: a 2R+1 deep register queue of the thread's own quad, rotated by *static*
//     renaming (the sweep loop is unrolled 2R+1 times) -- no data movement.
//   * y,z neighbours: the current plane with halo, (TY+2R) x (TZ+2HZ) floats, staged in shared
//     memory by ONE TMA box load (cp.async.bulk.tensor.3d, box depth 1).
//   * the queue is fed by a second, halo-less TMA box R planes ahead of the current one;
//     p(t-1) and v tiles arrive the same way (evict-first: they are streamed exactly once).
//   * a ring of STAGES such stage buffers is filled by one elected producer thread running
//     STAGES-1 sweep steps ahead; full/empty mbarriers, no __syncthreads in the sweep loop.
//   * results leave as 128-bit coalesced stores straight over p(t-1) (the reference's
//     2-slot write-back, /root/reference/src/compiler/lib/Var.cpp:435-464).
//
// Persistent grid: one CTA per SM; work units (y-tile, z-tile, x-chunk) are dealt round-robin so
// that the units in flight at any time are neighbours in (y,z) and share their halos through L2.
// ---------------------------------------------------------------------------------------------
template <int R_, int TY_, int TZQ_, int STAGES_>
struct IsoTile {
    static constexpr int R = R_, TY = TY_, TZQ = TZQ_, STAGES = STAGES_;
    static constexpr int TZ = 4 * TZQ;
    static constexpr int HZ = (R + 3) / 4 * 4;      // z halo kept in smem (multiple of 4 for LDS.128 alignment)
    static constexpr int ZQ = HZ / 4;               // halo quads each side
    static constexpr int HP = TZ + 2 * HZ;          // pitch of the haloed plane
    static constexpr int HROWS = TY + 2 * R;
    static constexpr int THREADS = TY * TZQ;
    static constexpr int NWARPS = THREADS / 32;
    static constexpr int QN = 2 * R + 1;            // register queue depth
    static constexpr uint32_t H_BYTES = HROWS * HP * 4;
    static constexpr uint32_t C_BYTES = TY * TZ * 4;
    static constexpr uint32_t H_OFF = 0;
    static constexpr uint32_t C_OFF = (H_BYTES + 127) / 128 * 128;
    static constexpr uint32_t P_OFF = C_OFF + C_BYTES;
    static constexpr uint32_t V_OFF = P_OFF + C_BYTES;
    static constexpr uint32_t STAGE_BYTES = V_OFF + C_BYTES;
    static constexpr uint32_t BAR_OFF = STAGES * STAGE_BYTES;
    static constexpr uint32_t SMEM_BYTES = BAR_OFF + 2 * STAGES * 8 + 128;  // +128: manual alignment slack
    static_assert(THREADS % 32 == 0, "whole warps");
    static_assert(C_BYTES % 128 == 0, "TMA destination alignment");
};

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
    int x0, y0, z0;  // unit origin (domain coordinates)
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
    cu.x0 = P.x_begin + u * P.lx;
    const int lxu = min(P.lx, P.x_end - cu.x0);
    cu.n_it = lxu + 2 * T::R;
    cu.it = 0;
}

template <class T, int MODE>
__global__ void __launch_bounds__(T::THREADS, 1)
iso3dfd_tma_kernel(const __grid_constant__ IsoMaps M, const __grid_constant__ IsoParams P) {
    constexpr int R = T::R, QN = T::QN, ZQ = T::ZQ;
    extern __shared__ uint8_t smem_raw[];
    // 128-B align the dynamic smem base (TMA destinations need it).
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbase + T::BAR_OFF);
    uint64_t* empty_bar = full_bar + T::STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int nunits = P.nty * P.ntz * P.nchunks;

    if (tid == 0) {
        tma_prefetch_desc(&M.h); tma_prefetch_desc(&M.c); tma_prefetch_desc(&M.p); tma_prefetch_desc(&M.v);
        for (int s = 0; s < T::STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], T::NWARPS); }
        fence_barrier_init();
    }
    __syncthreads();

    // ---- producer state (thread 0 only) -------------------------------------------------
    IsoCursor pr;
    pr.unit = blockIdx.x; pr.stage = 0; pr.phase = 0;
    const uint64_t pol_stream = l2_policy_evict_first();
    bool pr_live = (tid == 0) && (pr.unit < nunits);
    if (pr_live) iso_unit_setup<T>(pr, P);

    auto produce_one = [&]() {
        // wait until every consumer warp released this stage (first pass: passes immediately)
        mbar_wait(&empty_bar[pr.stage], pr.phase ^ 1u);
        uint8_t* st = sbase + pr.stage * T::STAGE_BYTES;
        uint64_t* fb = &full_bar[pr.stage];
        const bool compute = pr.it >= 2 * R;
        mbar_arrive_expect_tx(fb, compute ? (T::H_BYTES + 3 * T::C_BYTES) : T::C_BYTES);
        const int xc = pr.x0 - R + pr.it;  // plane entering the x queue
        tma_load_3d(st + T::C_OFF, &M.c, fb, P.pad_z + pr.z0, P.pad_y + pr.y0, P.pad_x + xc);
        if (compute) {
            const int xo = pr.x0 + pr.it - 2 * R;  // plane being computed
            tma_load_3d(st + T::H_OFF, &M.h, fb, P.pad_z + pr.z0 - T::HZ, P.pad_y + pr.y0 - R, P.pad_x + xo);
            tma_load_3d_hint(st + T::P_OFF, &M.p, fb, P.pad_z + pr.z0, P.pad_y + pr.y0, P.pad_x + xo, pol_stream);
            tma_load_3d_hint(st + T::V_OFF, &M.v, fb, P.vpad_z + pr.z0, P.vpad_y + pr.y0, P.vpad_x + xo, pol_stream);
        }
        if (++pr.stage == T::STAGES) { pr.stage = 0; pr.phase ^= 1u; }
        if (++pr.it == pr.n_it) {
            pr.unit += gridDim.x;
            if (pr.unit < nunits) iso_unit_setup<T>(pr, P); else pr_live = false;
        }
    };
    if (tid == 0) {
        for (int k = 0; k < T::STAGES - 1 && pr_live; k++) produce_one();
    }

    // ---- consumer -----------------------------------------------------------------------
    const int row = tid / T::TZQ;          // 0..TY-1
    const int quad = tid % T::TZQ;         // 0..TZQ-1
    const uint32_t h_own = ((row + R) * T::HP + T::HZ + 4 * quad) * 4;  // byte offset of own quad in H
    const uint32_t c_own = (row * T::TZ + 4 * quad) * 4;

    IsoCursor cu;
    cu.stage = 0; cu.phase = 0;
    float4 q[QN];
#pragma unroll
    for (int k = 0; k < QN; k++) q[k] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (cu.unit = blockIdx.x; cu.unit < nunits; cu.unit += gridDim.x) {
        iso_unit_setup<T>(cu, P);
        const int y = cu.y0 + row;
        const int zq = cu.z0 + 4 * quad;
        const bool row_ok = y < P.y_end;
        const int nvalid = row_ok ? max(0, min(4, P.z_end - zq)) : 0;
        float* out_col = P.out + (long long)y * P.out_sy + zq;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_col) & 15) == 0);

        for (int itb = 0; itb < cu.n_it; itb += QN) {
#pragma unroll
            for (int u = 0; u < QN; u++) {
                const int it = itb + u;
                if (it >= cu.n_it) break;
                // keep the ring full: issue the load STAGES-1 steps ahead
                if (tid == 0 && pr_live) produce_one();

                const uint8_t* st = sbase + cu.stage * T::STAGE_BYTES;
                mbar_wait(&full_bar[cu.stage], cu.phase);

                // newest plane (x0 - R + it) enters the queue slot u
                q[u] = *reinterpret_cast<const float4*>(st + T::C_OFF + c_own);

                if (it >= 2 * R) {
                    const float* hp = reinterpret_cast<const float*>(st + T::H_OFF + h_own);
                    // centre plane index in queue: it - R  -> slot (u - R) mod QN
                    const float4 cen = q[(u + QN - R) % QN];
                    // z window: quads -ZQ..+ZQ around own quad (centre taken from the queue)
                    float zw[4 * (2 * ZQ + 1)];
#pragma unroll
                    for (int k = -ZQ; k <= ZQ; k++) {
                        float4 t = (k == 0) ? cen : *reinterpret_cast<const float4*>(hp + 4 * k);
                        zw[4 * (k + ZQ) + 0] = t.x; zw[4 * (k + ZQ) + 1] = t.y;
                        zw[4 * (k + ZQ) + 2] = t.z; zw[4 * (k + ZQ) + 3] = t.w;
                    }
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
                    const float pc[4] = {cen.x, cen.y, cen.z, cen.w};
#pragma unroll
                    for (int r = 1; r <= R; r++) {
                        const float4 xm = q[(u + QN - R - r) % QN];
                        const float4 xp = q[(u + QN - R + r) % QN];
                        const float4 ym = *reinterpret_cast<const float4*>(hp - r * T::HP);
                        const float4 yp = *reinterpret_cast<const float4*>(hp + r * T::HP);
                        const float xm_[4] = {xm.x, xm.y, xm.z, xm.w}, xp_[4] = {xp.x, xp.y, xp.z, xp.w};
                        const float ym_[4] = {ym.x, ym.y, ym.z, ym.w}, yp_[4] = {yp.x, yp.y, yp.z, yp.w};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            acc[i] = iso_group<MODE>(acc[i], pc[i], P.c[0], P.c[r], xm_[i], xp_[i], ym_[i], yp_[i],
                                                     zw[4 * ZQ + i - r], zw[4 * ZQ + i + r], r == 1);
                        }
                    }
                    const float4 pv = *reinterpret_cast<const float4*>(st + T::P_OFF + c_own);
                    const float4 vv = *reinterpret_cast<const float4*>(st + T::V_OFF + c_own);
                    float4 res;
                    res.x = iso_final<MODE>(acc[0], pc[0], pv.x, vv.x);
                    res.y = iso_final<MODE>(acc[1], pc[1], pv.y, vv.y);
                    res.z = iso_final<MODE>(acc[2], pc[2], pv.z, vv.z);
                    res.w = iso_final<MODE>(acc[3], pc[3], pv.w, vv.w);
                    float* o = out_col + (long long)(cu.x0 + it - 2 * R) * P.out_sx;
                    if (nvalid == 4 && vec_ok) {
                        stg128(o, res);
                    } else if (nvalid > 0) {
                        o[0] = res.x;
                        if (nvalid > 1) o[1] = res.y;
                        if (nvalid > 2) o[2] = res.z;
                        if (nvalid > 3) o[3] = res.w;
                    }
                }
                // release the stage: one arrive per warp once all its lanes are done reading
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[cu.stage]);
                if (++cu.stage == T::STAGES) { cu.stage = 0; cu.phase ^= 1u; }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Tiled TMA kernel, generation 2 ("row-pair" threads).
//
// ncu on generation 1 (profiles/iso3dfd_r1_gen1.md) showed the sweep limited by shared-memory
// wavefronts (23 LDS.128 per 4 points, every y neighbour loaded once per point) and by
// instruction-cache misses of the 17x-unrolled body.  Generation 2 changes two things:
//   * a thread owns TWO vertically adjacent rows x 4 z (8 points).  The y window of the pair is
//     18 rows, each loaded ONCE and used by both rows: 16+8 neighbour loads per 8 points
//     instead of 2 x 20 -- 40 % fewer shared-memory wavefronts per point;
//   * the x queue is rotated by register moves (loop NOT unrolled 2R+1 times), so the whole
//     sweep body fits the instruction cache.
// Everything else (TMA stage ring, mbarriers, persistent unit scheduling, arithmetic order) is as
// in generation 1.
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
    pr.unit = blockIdx.x; pr.stage = 0; pr.phase = 0; pr.it = 0; pr.n_it = 0; pr.x0 = pr.y0 = pr.z0 = 0;
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
        tma_load_3d_hint(st + T::C_OFF, &M.c, fb, cz, cy, P.pad_x + pr.x0 - R + pr.it, pol_c);
        if (compute) {
            const int xo = pr.x0 + pr.it - 2 * R;
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

    for (cu.unit = blockIdx.x; cu.unit < nunits; cu.unit += gridDim.x) {
        iso_unit_setup<T>(cu, P);
        const int ya = cu.y0 + 2 * rp;
        const int zq = cu.z0 + 4 * quad;
        const int nz_ok = max(0, min(4, P.z_end - zq));
        const int nva = (ya < P.y_end) ? nz_ok : 0;
        const int nvb = (ya + 1 < P.y_end) ? nz_ok : 0;
        float* out_a = P.out + (long long)ya * P.out_sy + zq + (long long)(cu.x0 - 2 * R) * P.out_sx;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_a) & 15) == 0);

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
                    float* oa = out_a + (long long)it * P.out_sx;
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
                float* oa = out_a + (long long)it * P.out_sx;
                float* ob = oa + P.out_sy;
                if (vec_ok && nva == 4) { if (P.st_cs) stg128_cs(oa, ra); else stg128(oa, ra); }
                else if (nva > 0) { oa[0] = ra.x; if (nva > 1) oa[1] = ra.y; if (nva > 2) oa[2] = ra.z; if (nva > 3) oa[3] = ra.w; }
                if (vec_ok && nvb == 4) { if (P.st_cs) stg128_cs(ob, rb); else stg128(ob, rb); }
                else if (nvb > 0) { ob[0] = rb.x; if (nvb > 1) ob[1] = rb.y; if (nvb > 2) ob[2] = rb.z; if (nvb > 3) ob[3] = rb.w; }
                // fused halo exchange: boundary planes also go straight into the x neighbours' halo cells
                const int xo = cu.x0 + it - 2 * R;
                float* peer = (P.peer_lo != nullptr && xo < R) ? P.peer_lo : ((P.peer_hi != nullptr && xo >= P.nx - R) ? P.peer_hi : nullptr);
                if (peer != nullptr) {
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
}

// ---------------------------------------------------------------------------------------------
// Tiled TMA kernel, generation 3 ("resident planes").
//
// Generation 2 fetches every plane twice from L2 (halo-less, R planes early, to feed the x queue, and
// again with its halo when it becomes the current plane); ncu showed 21.5 GB of TMA reads and 19.7 GB
// of DRAM traffic per 1024^3 launch against 17.2 GB algorithmic, and a memory-only probe of that traffic
// pattern capped at ~293-315 GPts/s.  Generation 3 loads each plane ONCE, with its halo, into a ring of
// NS > R shared-memory slots where it stays resident from the step it feeds the queue (x+R) until it has
// been the current plane (x): one haloed box per plane instead of two boxes, no R-plane reuse distance in
// L2.  p(t-1) and v travel through their own small ring.  Threads, arithmetic and stores are those of
// generation 2.
//
//   iteration I (per CTA, global over its work units):
//     feed    : own quads of plane I           <- H slot  I      mod NS   (wait full_h)
//     compute : plane I - R (if inside the unit's sweep) from H slot (I-R) mod NS, P/V slot c mod NPV
//     release : every warp arrives on done[I mod 4]; the producer thread, at the top of iteration I+1,
//               waits for it, then refills the H slot of plane I-R (with plane I-R+NS) and the next P/V slot.
// ---------------------------------------------------------------------------------------------
template <int R_, int TYP_, int TZQ_, int NS_, int NPV_>
struct IsoTile3 {
    static constexpr int R = R_, TYP = TYP_, TZQ = TZQ_, NS = NS_, NPV = NPV_;
    static constexpr int TY = 2 * TYP;
    static constexpr int TZ = 4 * TZQ;
    static constexpr int HZ = (R + 3) / 4 * 4;
    static constexpr int ZQ = HZ / 4;
    static constexpr int HP = TZ + 2 * HZ;
    static constexpr int HROWS = TY + 2 * R;
    static constexpr int THREADS = TYP * TZQ;
    static constexpr int NWARPS = THREADS / 32;
    static constexpr int QN = 2 * R + 1;
    static constexpr int NDONE = 4;
    static constexpr uint32_t H_BYTES = HROWS * HP * 4;
    static constexpr uint32_t H_STRIDE = (H_BYTES + 127) / 128 * 128;
    static constexpr uint32_t C_BYTES = TY * TZ * 4;
    static constexpr uint32_t PV_STRIDE = 2 * C_BYTES;            // P tile then V tile
    static constexpr uint32_t PV_OFF = NS * H_STRIDE;
    static constexpr uint32_t BAR_OFF = PV_OFF + NPV * PV_STRIDE;
    static constexpr uint32_t SMEM_BYTES = BAR_OFF + (NS + NPV + NDONE) * 8 + 128;
    static_assert(NS > R + 1, "ring must hold planes x .. x+R plus at least one in flight");
    static_assert(THREADS % 32 == 0, "whole warps");
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can use");
};

template <class T, int MODE>
__global__ void __launch_bounds__(T::THREADS, 1)
iso3dfd_tma3_kernel(const __grid_constant__ IsoMaps M, const __grid_constant__ IsoParams P) {
    constexpr int R = T::R, QN = T::QN, ZQ = T::ZQ, NS = T::NS, NPV = T::NPV, NDONE = T::NDONE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* sbase = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* full_h = reinterpret_cast<uint64_t*>(sbase + T::BAR_OFF);
    uint64_t* full_pv = full_h + NS;
    uint64_t* done_bar = full_pv + NPV;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int nunits = P.nty * P.ntz * P.nchunks;

    if (tid == 0) {
        tma_prefetch_desc(&M.h); tma_prefetch_desc(&M.p); tma_prefetch_desc(&M.v);
        for (int s = 0; s < NS; s++) mbar_init(&full_h[s], 1);
        for (int s = 0; s < NPV; s++) mbar_init(&full_pv[s], 1);
        for (int s = 0; s < NDONE; s++) mbar_init(&done_bar[s], T::NWARPS);
        fence_barrier_init();
    }
    __syncthreads();

    // ---- producer state (thread 0) ---------------------------------------------------------------------
    const uint64_t pol_pv = l2_policy(P.pol_pv), pol_h = l2_policy(P.pol_h);
    IsoCursor ph;   // next haloed plane to request: it = plane index within unit, 0 .. lx_u + 2R - 1
    IsoCursor pp;   // next P/V tile to request:     it = compute index within unit, 0 .. lx_u - 1
    ph.unit = pp.unit = blockIdx.x;
    ph.stage = pp.stage = 0; ph.phase = pp.phase = 0;
    ph.it = pp.it = 0; ph.n_it = pp.n_it = 0; ph.x0 = ph.y0 = ph.z0 = pp.x0 = pp.y0 = pp.z0 = 0;
    int pg = 0;          // global index of the next plane to request
    int pc = 0;          // global index of the next compute step whose P/V to request
    bool ph_live = (tid == 0) && (blockIdx.x < nunits), pp_live = ph_live;
    if (ph_live) { iso_unit_setup<T>(ph, P); iso_unit_setup<T>(pp, P); pp.n_it -= 2 * R; }

    auto request_plane = [&]() {
        uint64_t* fb = &full_h[ph.stage];
        mbar_arrive_expect_tx(fb, T::H_BYTES);
        tma_load_3d_hint(sbase + ph.stage * T::H_STRIDE, &M.h, fb, P.pad_z + ph.z0 - T::HZ, P.pad_y + ph.y0 - R, P.pad_x + ph.x0 - R + ph.it, pol_h);
        if (++ph.stage == NS) ph.stage = 0;
        pg++;
        if (++ph.it == ph.n_it) {
            ph.unit += gridDim.x;
            if (ph.unit < nunits) iso_unit_setup<T>(ph, P); else ph_live = false;
        }
    };
    auto request_pv = [&]() {
        uint64_t* fb = &full_pv[pp.stage];
        uint8_t* dst = sbase + T::PV_OFF + pp.stage * T::PV_STRIDE;
        mbar_arrive_expect_tx(fb, 2 * T::C_BYTES);
        tma_load_3d_hint(dst, &M.p, fb, P.pad_z + pp.z0, P.pad_y + pp.y0, P.pad_x + pp.x0 + pp.it, pol_pv);
        tma_load_3d_hint(dst + T::C_BYTES, &M.v, fb, P.vpad_z + pp.z0, P.vpad_y + pp.y0, P.vpad_x + pp.x0 + pp.it, pol_pv);
        if (++pp.stage == NPV) pp.stage = 0;
        pc++;
        if (++pp.it == pp.n_it) {
            pp.unit += gridDim.x;
            if (pp.unit < nunits) { iso_unit_setup<T>(pp, P); pp.n_it -= 2 * R; } else pp_live = false;
        }
    };

    // ---- consumer state --------------------------------------------------------------------------------
    const int rp = tid / T::TZQ;
    const int quad = tid % T::TZQ;
    const uint32_t h_own = ((2 * rp + R) * T::HP + T::HZ + 4 * quad) * 4;   // row a centre inside an H slot
    const uint32_t c_own = (2 * rp * T::TZ + 4 * quad) * 4;                 // row a inside a P or V tile

    int I = 0;                         // global iteration
    int cdone = 0;                     // compute steps finished before iteration I (all warps)
    int feed_slot = 0, cur_slot = NS - R % NS;      // I mod NS and (I - R) mod NS
    if (cur_slot >= NS) cur_slot -= NS;
    uint32_t feed_par = 0;             // parity of the fill of feed_slot that holds plane I
    int pv_slot = 0; uint32_t pv_par = 0;
    int done_slot = 0; uint32_t done_par = 0;       // barrier of iteration I
    float4 qa[QN], qb[QN];
#pragma unroll
    for (int k = 0; k < QN; k++) { qa[k] = make_float4(0.f, 0.f, 0.f, 0.f); qb[k] = qa[k]; }

    IsoCursor cu;
    for (cu.unit = blockIdx.x; cu.unit < nunits; cu.unit += gridDim.x) {
        iso_unit_setup<T>(cu, P);
        const int ya = cu.y0 + 2 * rp;
        const int zq = cu.z0 + 4 * quad;
        const int nz_ok = max(0, min(4, P.z_end - zq));
        const int nva = (ya < P.y_end) ? nz_ok : 0;
        const int nvb = (ya + 1 < P.y_end) ? nz_ok : 0;
        float* out_a = P.out + (long long)ya * P.out_sy + zq + (long long)(cu.x0 - 2 * R) * P.out_sx;
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(out_a) & 15) == 0);

#pragma unroll 1
        for (int it = 0; it < cu.n_it; it++, I++) {
            const bool compute = it >= 2 * R;
            if (tid == 0) {
                // all warps have finished iteration I-1: its current plane's slot and its P/V slot are free
                if (I > 0) {
                    const int ds = done_slot == 0 ? NDONE - 1 : done_slot - 1;
                    const uint32_t dp = done_slot == 0 ? done_par ^ 1u : done_par;
                    mbar_wait(&done_bar[ds], dp);
                }
                while (ph_live && pg <= I + NS - R - 1) request_plane();
                while (pp_live && pc < cdone + NPV) request_pv();
            }

            // feed the x queues with plane I
            mbar_wait(&full_h[feed_slot], feed_par);
            {
                const uint8_t* hs = sbase + feed_slot * T::H_STRIDE + h_own;
#pragma unroll
                for (int k = 0; k < QN - 1; k++) { qa[k] = qa[k + 1]; qb[k] = qb[k + 1]; }
                qa[QN - 1] = *reinterpret_cast<const float4*>(hs);
                qb[QN - 1] = *reinterpret_cast<const float4*>(hs + T::HP * 4);
            }

            if (compute) {
                mbar_wait(&full_pv[pv_slot], pv_par);
                const float* hp = reinterpret_cast<const float*>(sbase + cur_slot * T::H_STRIDE + h_own);
                const uint8_t* pvs = sbase + T::PV_OFF + pv_slot * T::PV_STRIDE + c_own;
                float pa[4], pb[4];
                f4_to_arr(qa[R], pa);
                f4_to_arr(qb[R], pb);
                float za[4 * (2 * ZQ + 1)], zb[4 * (2 * ZQ + 1)];
#pragma unroll
                for (int k = -ZQ; k <= ZQ; k++) {
                    if (k == 0) { f4_to_arr(qa[R], &za[4 * ZQ]); f4_to_arr(qb[R], &zb[4 * ZQ]); continue; }
                    f4_to_arr(*reinterpret_cast<const float4*>(hp + 4 * k), &za[4 * (k + ZQ)]);
                    f4_to_arr(*reinterpret_cast<const float4*>(hp + T::HP + 4 * k), &zb[4 * (k + ZQ)]);
                }
                float acca[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};
                float wlo_prev[4], whi_prev[4];
                f4_to_arr(qa[R], wlo_prev);
                f4_to_arr(qb[R], whi_prev);
#pragma unroll
                for (int r = 1; r <= R; r++) {
                    float wlo[4], whi[4], xm[4], xp[4];
                    f4_to_arr(*reinterpret_cast<const float4*>(hp - r * T::HP), wlo);
                    f4_to_arr(*reinterpret_cast<const float4*>(hp + (r + 1) * T::HP), whi);
                    f4_to_arr(qa[R - r], xm); f4_to_arr(qa[R + r], xp);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        acca[i] = iso_group<MODE>(acca[i], pa[i], P.c[0], P.c[r], xm[i], xp[i], wlo[i], whi_prev[i],
                                                  za[4 * ZQ + i - r], za[4 * ZQ + i + r], r == 1);
                    f4_to_arr(qb[R - r], xm); f4_to_arr(qb[R + r], xp);
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        accb[i] = iso_group<MODE>(accb[i], pb[i], P.c[0], P.c[r], xm[i], xp[i], wlo_prev[i], whi[i],
                                                  zb[4 * ZQ + i - r], zb[4 * ZQ + i + r], r == 1);
#pragma unroll
                    for (int i = 0; i < 4; i++) { wlo_prev[i] = wlo[i]; whi_prev[i] = whi[i]; }
                }
                const float4 pva = *reinterpret_cast<const float4*>(pvs);
                const float4 vva = *reinterpret_cast<const float4*>(pvs + T::C_BYTES);
                const float4 pvb = *reinterpret_cast<const float4*>(pvs + T::TZ * 4);
                const float4 vvb = *reinterpret_cast<const float4*>(pvs + T::C_BYTES + T::TZ * 4);
                float4 ra, rb;
                ra.x = iso_final<MODE>(acca[0], pa[0], pva.x, vva.x); ra.y = iso_final<MODE>(acca[1], pa[1], pva.y, vva.y);
                ra.z = iso_final<MODE>(acca[2], pa[2], pva.z, vva.z); ra.w = iso_final<MODE>(acca[3], pa[3], pva.w, vva.w);
                rb.x = iso_final<MODE>(accb[0], pb[0], pvb.x, vvb.x); rb.y = iso_final<MODE>(accb[1], pb[1], pvb.y, vvb.y);
                rb.z = iso_final<MODE>(accb[2], pb[2], pvb.z, vvb.z); rb.w = iso_final<MODE>(accb[3], pb[3], pvb.w, vvb.w);
                float* oa = out_a + (long long)it * P.out_sx;
                float* ob = oa + P.out_sy;
                if (vec_ok && nva == 4) stg128(oa, ra);
                else if (nva > 0) { oa[0] = ra.x; if (nva > 1) oa[1] = ra.y; if (nva > 2) oa[2] = ra.z; if (nva > 3) oa[3] = ra.w; }
                if (vec_ok && nvb == 4) stg128(ob, rb);
                else if (nvb > 0) { ob[0] = rb.x; if (nvb > 1) ob[1] = rb.y; if (nvb > 2) ob[2] = rb.z; if (nvb > 3) ob[3] = rb.w; }
                if (++pv_slot == NPV) { pv_slot = 0; pv_par ^= 1u; }
                cdone++;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&done_bar[done_slot]);
            if (++done_slot == NDONE) { done_slot = 0; done_par ^= 1u; }
            if (++feed_slot == NS) { feed_slot = 0; feed_par ^= 1u; }
            if (++cur_slot == NS) cur_slot = 0;
        }
    }
}

}  // namespace yb
