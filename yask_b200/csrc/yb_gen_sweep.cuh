// "Sweep" variant of the emitter-generated kernels: TMA-staged shared-memory planes with the same 2.5-D march along x
// as the hand-written iso3dfd kernel, for multi-var stencils.  EXPERIMENTAL (option gen_sweep=1, off by default).
//
// A CTA of 256 threads (two row groups of 128) owns a (TY rows x 128 z) tile and marches over a chunk of x planes.  Every full-rank var the
// part reads is a *stream*: its planes arrive by TMA (cp.async.bulk.tensor.3d, box = tile + the stream's y/z reach)
// into a shared-memory ring of (x reach + PF) slots, PF planes ahead of their first use, and complete on one mbarrier
// per sweep iteration.  The statements are the same as in the direct kernel (same order, same rounding); only the
// full-rank reads come from shared memory.  Lower-rank vars (1-D sponge arrays, scalars) are read from global memory
// as before; outputs are stored straight to global memory.
#pragma once
#include "yb_gen.cuh"
#include "yb_ptx.cuh"

namespace yb { namespace gen {

constexpr int GEN_SW_MAX_STREAMS = 24;
constexpr int GEN_SW_TZ = 128;          // z extent of a tile
constexpr int GEN_SW_THREADS = 256;     // two row groups of GEN_SW_TZ threads: thread (tz, g) computes rows g*TY/2 .. g*TY/2 + TY/2 - 1

struct GenSweepParams {
    GenParams g;                               // box, pointers and strides as for the direct kernel
    const CUtensorMap* maps;                   // device array, one per stream: 3-D (z, y, x) view of the var's step slot,
                                               // box (pz, rows, 1); built once per slot combination and cached (yb_gen.cu)
    int px, py, pz;                            // left pads of the shared geometry (tensor coordinate of local index 0)
    int lx, nchunks;                           // planes per sweep chunk, chunks along x
    int nzb, nyb;                              // tiles along z and y
    int bar_off;                               // byte offset of the mbarriers in dynamic shared memory
};

// (GenSweepFn, GenSweepStream and GenSweep -- the host-side description filled in by the generated describe() -- are
// declared in yb_gen.cuh next to GenPart.)

#define GEN_SW_FN(k, T, M) reinterpret_cast<yb::gen::GenSweepFn>(static_cast<void (*)(const yb::gen::GenSweepParams)>(k<T, M>))

#ifdef __CUDACC__
// Prologue shared by all sweep kernels: tile coordinates, barrier set-up.
#define GEN_SWEEP_BEGIN(TY_, PF_)                                                                            \
    extern __shared__ __align__(128) unsigned char sw_smem[];                                                \
    const GenParams& P = SP.g;                                                                               \
    constexpr int SW_TY = (TY_), SW_PF = (PF_), SW_NB = (PF_) + 1;                                           \
    uint64_t* sw_bar = reinterpret_cast<uint64_t*>(sw_smem + SP.bar_off);                                    \
    const int sw_bz = int(blockIdx.x) % SP.nzb;                                                              \
    const int sw_by = (int(blockIdx.x) / SP.nzb) % SP.nyb;                                                   \
    const int sw_bc = int(blockIdx.x) / (SP.nzb * SP.nyb);                                                   \
    const int z0 = P.zb + sw_bz * GEN_SW_TZ;                                                                 \
    const int y0_ = P.yb + sw_by * SW_TY;                                                                    \
    const int xs = P.xb + sw_bc * SP.lx;                                                                     \
    const int sw_len = min(SP.lx, P.xe - xs);                                                                \
    const int tz = int(threadIdx.x) & (GEN_SW_TZ - 1);                                                       \
    const int sw_rb = (int(threadIdx.x) / GEN_SW_TZ) * (SW_TY / 2);   /* first row of this thread's group */  \
    const int z = z0 + tz;                                                                                   \
    if (threadIdx.x == 0) {                                                                                  \
        for (int b = 0; b < SW_NB; b++) mbar_init(&sw_bar[b], 1);                                            \
        fence_barrier_init();                                                                                \
    }                                                                                                        \
    __syncthreads();

// One stream's loads for sweep iteration j (j == 0: the whole initial x reach; j > 0: the newest plane).
#define SW_LOAD(k, OFF, SLOT, NS, XL, XR, YL, ZL)                                                            \
    if (j == 0) {                                                                                            \
        for (int dx = (XL); dx <= (XR); dx++)                                                                \
            tma_load_3d(sw_smem + (OFF) + ((dx - (XL)) % (NS)) * (SLOT), &SP.maps[k], bar, SP.pz + z0 + (ZL), \
                        SP.py + y0_ + (YL), SP.px + xs + dx);                                                \
    } else {                                                                                                 \
        tma_load_3d(sw_smem + (OFF) + ((j + (XR) - (XL)) % (NS)) * (SLOT), &SP.maps[k], bar, SP.pz + z0 + (ZL), \
                    SP.py + y0_ + (YL), SP.px + xs + j + (XR));                                              \
    }

// Base pointer of stream k's plane x+dx for this thread: box row (YL) of the thread's first row, its z column.
#define SW_PLANE(OFF, SLOT, NS, XL, ZL, dx, PZ)                                                              \
    (reinterpret_cast<const T*>(sw_smem + (OFF) + ((unsigned(it) + unsigned((dx) - (XL))) % unsigned(NS)) * (SLOT)) + (tz - (ZL)) + sw_rb * (PZ))
#endif

} }  // namespace yb::gen
